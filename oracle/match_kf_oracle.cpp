// match_kf_oracle.cpp -- CPU oracle of the key-frame flavoured Hamming searches of ORBmatcher:
//   SearchByProjectionOnCam(F, cam, KF, sFound, th, ORBdist)   src/ORBmatcher.cc:812-951   (relocalisation)
//   SearchByProjection(KF, vpMapPoints, sFound, th, ORBdist)   src/ORBmatcher.cc:693-799   (search part, see below)
//   SearchByProjection(KF, query, Scw, vpPoints, vpMatched,th) src/ORBmatcher.cc:416-536   (loop closing)
//   Fuse(KF, vpMapPoints, th) / Fuse(KF, Scw, ...)             src/ORBmatcher.cc:1431-1556, 1560-1712 (search part)
//   SearchByBoWCrossCam(KF1, c1, KF2, c2, vpMatches12)         src/ORBmatcher.cc:297-414
//   SearchForTriangulation(KF1, KF2, F12, pairs, camS)         src/ORBmatcher.cc:1253-1427
//
// TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  PARITY UNPINNED by the reference (no tests or fixtures upstream for these functions).
// The functions that mutate the map inside their loop (:693-799, Fuse) are restated up to the decision "best key point and its
// distance for (camera, map point)"; AddMapPoint / Replace / AddObservation stay with the caller.
// Same determinisations as match_oracle.cpp: 3x3 * 3x1 products in FP32 left to right, cv::norm / Mat::dot accumulate in double,
// PredictScale's log() in double.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orb_oracle.h"

namespace {

const int TH_LOW = 50, HISTO_LENGTH = 30;
const int GRID_COLS = 64, GRID_ROWS = 48;

int descriptor_distance(const uint8_t* a, const uint8_t* b) {   // src/ORBmatcher.cc:2015-2031
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        std::memcpy(&pa, a + 4 * i, 4);
        std::memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// the per-camera grid a KeyFrame copies from its Frame (src/KeyFrame.cc ctor; filled by src/Frame.cc:179-196,380-390)
struct KFView {
    const orc_frame_t* F;
    int kf_quirk;                     // 1: KeyFrame::GetFeaturesInArea as upstream -- mvTotalKeysUn[camera-LOCAL index] (src/KeyFrame.cc:757)
    std::vector<int> first;
    std::vector<float> invW, invH;
    std::vector<std::vector<std::vector<std::vector<int>>>> grid;
    KFView(const orc_frame_t* f, int quirk) : F(f), kf_quirk(quirk) {
        const int C = f->n_cams;
        first.assign(C + 1, 0);
        for (int c = 0; c < C; c++) first[c + 1] = first[c] + f->n_kp[c];
        invW.resize(C); invH.resize(C); grid.resize(C);
        for (int c = 0; c < C; c++) {
            const float* b = f->bounds + 4 * c;
            invW[c] = (float)GRID_COLS / (float)(b[1] - b[0]);
            invH[c] = (float)GRID_ROWS / (float)(b[3] - b[2]);
            grid[c].assign(GRID_COLS, std::vector<std::vector<int>>(GRID_ROWS));
            for (int i = 0; i < f->n_kp[c]; i++) {
                const orc_keypoint_t& kp = f->kps_un[first[c] + i];
                const int px = (int)lrintf((kp.x - b[0]) * invW[c]);
                const int py = (int)lrintf((kp.y - b[2]) * invH[c]);
                if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
                grid[c][px][py].push_back(i);
            }
        }
    }
    // Frame::GetFeaturesInArea with level filter  src/Frame.cc:316-376
    std::vector<int> frame_features_in_area(int c, float x, float y, float r, int minLevel, int maxLevel) const {
        std::vector<int> v;
        const float* b = F->bounds + 4 * c;
        const int nMinCellX = std::max(0, (int)std::floor((x - b[0] - r) * invW[c]));
        if (nMinCellX >= GRID_COLS) return v;
        const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - b[0] + r) * invW[c]));
        if (nMaxCellX < 0) return v;
        const int nMinCellY = std::max(0, (int)std::floor((y - b[2] - r) * invH[c]));
        if (nMinCellY >= GRID_ROWS) return v;
        const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - b[2] + r) * invH[c]));
        if (nMaxCellY < 0) return v;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (int local : grid[c][ix][iy]) {
                    const orc_keypoint_t& kp = F->kps_un[first[c] + local];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    if (std::fabs(kp.x - x) < r && std::fabs(kp.y - y) < r) v.push_back(local);
                }
        return v;
    }
    // KeyFrame::GetFeaturesInArea  src/KeyFrame.cc:729-768 (no level filter)
    std::vector<int> kf_features_in_area(int c, float x, float y, float r) const {
        std::vector<int> v;
        const float* b = F->bounds + 4 * c;
        const int nMinCellX = std::max(0, (int)std::floor((x - b[0] - r) * invW[c]));
        if (nMinCellX >= GRID_COLS) return v;
        const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - b[0] + r) * invW[c]));
        if (nMaxCellX < 0) return v;
        const int nMinCellY = std::max(0, (int)std::floor((y - b[2] - r) * invH[c]));
        if (nMinCellY >= GRID_ROWS) return v;
        const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - b[2] + r) * invH[c]));
        if (nMaxCellY < 0) return v;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (int local : grid[c][ix][iy]) {
                    const orc_keypoint_t& kpUn = F->kps_un[kf_quirk ? local : first[c] + local];     // :757
                    if (std::fabs(kpUn.x - x) < r && std::fabs(kpUn.y - y) < r) v.push_back(local);
                }
        return v;
    }
};

int predict_scale(float mfMaxDistance, float currentDist, float logScaleFactor, int nLevels) {   // src/MapPoint.cc:423-455
    const float ratio = mfMaxDistance / currentDist;
    int nScale = (int)std::ceil(std::log((double)ratio) / (double)logScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= nLevels) nScale = nLevels - 1;
    return nScale;
}

float norm3(const float* v) { return (float)std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]); }   // cv::norm
double dot3(const float* a, const float* b) { return (double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2]; }  // Mat::dot

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {   // src/ORBmatcher.cc:1969-2010
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

int rot_bin(float rot) {
    const float factor = 1.0f / HISTO_LENGTH;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

}  // namespace

extern "C" {

// SearchByProjectionOnCam(pF, query, pKF, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:812-951.
// P lists pKF->GetMapPointMatches() (valid = pMP && !isBad() && !sAlreadyFound.count(pMP); angle = pKF->mvTotalKeysUn[i].angle);
// blocked[g] = pF->mvpMapPoints[g] != NULL.  kp_to_point[g] receives i where pF->mvpMapPoints[g] = vpMPs[i], -1 where the
// rotation check resets it to NULL.
int orc_search_by_projection_reloc(const orc_frame_t* F, const orc_frustum_t* V, int query, const orc_points_t* P, float th, int ORBdist,
                                   int check_ori, const uint8_t* blocked_in, int32_t* kp_to_point) {
    KFView G(F, 0);
    const int totalN = G.first[F->n_cams];
    std::vector<uint8_t> blocked(blocked_in, blocked_in + totalN);
    int nmatches = 0;
    const float* R = V->Rsw + 9 * query;
    const float* t = V->tsw + 3 * query;
    const float* Osw = V->Ow + 3 * query;
    const float fx = V->K[4 * query], fy = V->K[4 * query + 1], cx = V->K[4 * query + 2], cy = V->K[4 * query + 3];
    const float* b = F->bounds + 4 * query;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < P->n; i++) {
        if (!P->valid[i]) continue;
        const float* X = P->pos + 3 * (size_t)i;
        const float xs = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
        const float ys = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
        const float zs = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
        const float invzs = (float)(1.0 / zs);                                 // no depth test here (:848)
        const float u = fx * xs * invzs + cx;
        const float v = fy * ys * invzs + cy;
        if (u < b[0] || u > b[1]) continue;
        if (v < b[2] || v > b[3]) continue;
        const float PO[3] = {X[0] - Osw[0], X[1] - Osw[1], X[2] - Osw[2]};
        const float dist3D = norm3(PO);
        const float maxDistance = 1.2f * P->max_dist[i], minDistance = 0.8f * P->min_dist[i];
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = predict_scale(P->max_dist[i], dist3D, V->log_scale_factor, F->n_levels);
        const float radius = th * F->scale_factors[nPredictedLevel];
        const std::vector<int> vIndicesFrame = G.frame_features_in_area(query, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1);
        if (vIndicesFrame.empty()) continue;
        int bestDist = 256, bestglobalIdx = -1;
        for (int local : vIndicesFrame) {
            const int g = G.first[query] + local;
            if (blocked[g]) continue;
            const int dist = descriptor_distance(P->desc + 32 * (size_t)i, F->desc + 32 * (size_t)g);
            if (dist < bestDist) { bestDist = dist; bestglobalIdx = g; }
        }
        if (bestDist <= ORBdist) {
            kp_to_point[bestglobalIdx] = i;
            blocked[bestglobalIdx] = 1;
            nmatches++;
            if (check_ori) rotHist[rot_bin(P->angle[i] - F->kps_un[bestglobalIdx].angle)].push_back(bestglobalIdx);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int g : rotHist[i]) { kp_to_point[g] = -1; nmatches--; }
    }
    return nmatches;
}

// SearchByProjection(pKF, query, Scq_w, vpPoints, vpMatched, th)  src/ORBmatcher.cc:416-536.  V slot `query` holds the decomposed
// similarity (Rcqw, tcqw, Ocqw :431-435).  valid = !isBad() && !spAlreadyFound.count(pMP).  matched_local / local_to_point are
// indexed like the reference indexes vpMatched: by the CAMERA-LOCAL key point index (:504, :523).
int orc_search_by_projection_sim3(const orc_frame_t* KF, const orc_frustum_t* V, int query, const orc_points_t* P, int th, int kf_quirk,
                                  const uint8_t* matched_local, int32_t* local_to_point) {
    KFView G(KF, kf_quirk);
    const int nLocal = KF->n_kp[query];
    std::vector<uint8_t> matched(matched_local, matched_local + nLocal);
    const float* R = V->Rsw + 9 * query;
    const float* t = V->tsw + 3 * query;
    const float* Ocqw = V->Ow + 3 * query;
    const float fx = V->K[4 * query], fy = V->K[4 * query + 1], cx = V->K[4 * query + 2], cy = V->K[4 * query + 3];
    const float* b = KF->bounds + 4 * query;
    int nmatches = 0;
    for (int iMP = 0; iMP < P->n; iMP++) {
        if (!P->valid[iMP]) continue;
        const float* X = P->pos + 3 * (size_t)iMP;
        const float pc0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
        const float pc1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
        const float pc2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
        if (pc2 < 0.0) continue;
        const float invz = 1 / pc2;
        const float x = pc0 * invz, y = pc1 * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!(u >= b[0] && u < b[1] && v >= b[2] && v < b[3])) continue;                 // KeyFrame::IsInImage  src/KeyFrame.cc:770-773
        const float maxDistance = 1.2f * P->max_dist[iMP], minDistance = 0.8f * P->min_dist[iMP];
        const float PO[3] = {X[0] - Ocqw[0], X[1] - Ocqw[1], X[2] - Ocqw[2]};
        const float dist = norm3(PO);
        if (dist < minDistance || dist > maxDistance) continue;
        if (dot3(PO, P->normal + 3 * (size_t)iMP) < 0.5 * dist) continue;
        const int nPredictedLevel = predict_scale(P->max_dist[iMP], dist, V->log_scale_factor, KF->n_levels);
        const float radius = th * KF->scale_factors[nPredictedLevel];
        const std::vector<int> vIndicesLocal = G.kf_features_in_area(query, u, v, radius);
        if (vIndicesLocal.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idxLocal : vIndicesLocal) {
            if (matched[idxLocal]) continue;
            const int kpLevel = KF->kps_un[G.first[query] + idxLocal].octave;              // mvvkeysUnTemp[query][idxLocal]
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int d = descriptor_distance(P->desc + 32 * (size_t)iMP, KF->desc + 32 * (size_t)(G.first[query] + idxLocal));
            if (d < bestDist) { bestDist = d; bestIdx = idxLocal; }
        }
        if (bestDist <= TH_LOW) { local_to_point[bestIdx] = iMP; matched[bestIdx] = 1; nmatches++; }
    }
    return nmatches;
}

// Search part of SearchByProjection(pKF, vpMapPoints, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:693-775: for camera s and map
// point i the key point (global index) with the smallest distance inside the window, and that distance (256 / -1 when none).
// The caller applies `bestDist <= ORBdist` and AddMapPoint / Replace (:777-793) in (s, i) order.
void orc_search_kf_points(const orc_frame_t* KF, const orc_frustum_t* V, const orc_points_t* P, float th, int kf_quirk, int32_t* best_kp,
                          int32_t* best_dist) {
    KFView G(KF, kf_quirk);
    for (int s = 0; s < KF->n_cams; s++) {
        const float* R = V->Rsw + 9 * s;
        const float* t = V->tsw + 3 * s;
        const float* Osw = V->Ow + 3 * s;
        const float* b = KF->bounds + 4 * s;
        for (int i = 0; i < P->n; i++) {
            int32_t& okp = best_kp[(size_t)s * P->n + i];
            int32_t& od = best_dist[(size_t)s * P->n + i];
            okp = -1; od = 256;
            if (!P->valid[i]) continue;
            const float* X = P->pos + 3 * (size_t)i;
            const float xs = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
            const float ys = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
            const float zs = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
            const float invzs = (float)(1.0 / zs);
            const float u = V->K[4 * s] * xs * invzs + V->K[4 * s + 2];
            const float v = V->K[4 * s + 1] * ys * invzs + V->K[4 * s + 3];
            if (u < b[0] || u > b[1]) continue;
            if (v < b[2] || v > b[3]) continue;
            const float PO[3] = {X[0] - Osw[0], X[1] - Osw[1], X[2] - Osw[2]};
            const float dist3D = norm3(PO);
            const float maxDistance = 1.2f * P->max_dist[i], minDistance = 0.8f * P->min_dist[i];
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            const int nPredictedLevel = predict_scale(P->max_dist[i], dist3D, V->log_scale_factor, KF->n_levels);
            const float radius = th * KF->scale_factors[nPredictedLevel];
            const std::vector<int> vIndicesInCam = G.kf_features_in_area(s, u, v, radius);
            if (vIndicesInCam.empty()) continue;
            int bestDist = 256, bestIdxglobal = -1;
            for (int localKpIdx : vIndicesInCam) {
                const int globalKpIdx = G.first[s] + localKpIdx;
                const int kpLevel = KF->kps_un[globalKpIdx].octave;
                if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel + 1) continue;
                const int dist = descriptor_distance(P->desc + 32 * (size_t)i, KF->desc + 32 * (size_t)globalKpIdx);
                if (dist < bestDist) { bestDist = dist; bestIdxglobal = globalKpIdx; }
            }
            okp = bestIdxglobal; od = bestDist;
        }
    }
}

// Search part of Fuse(pKF, vpMapPoints, th)  src/ORBmatcher.cc:1431-1527 (sim3 == 0) and of Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)
// src/ORBmatcher.cc:1560-1690 (sim3 == 1; V then holds the per-camera Rsw / tsw / Osw of :1573-1597).  valid = pMP && !isBad() &&
// !IsInKeyFrame(pKF) resp. !isBad() && !spAlreadyFound.count(pMP).  Output as orc_search_kf_points; the caller applies
// `bestDist <= TH_LOW` and the Replace / AddObservation branch.
void orc_fuse(const orc_frame_t* KF, const orc_frustum_t* V, const orc_points_t* P, float th, int sim3, int kf_quirk, int32_t* best_kp,
              int32_t* best_dist) {
    KFView G(KF, kf_quirk);
    for (int ic = 0; ic < KF->n_cams; ic++) {
        const float* R = V->Rsw + 9 * ic;
        const float* t = V->tsw + 3 * ic;
        const float* Os = V->Ow + 3 * ic;
        const float fx = V->K[4 * ic], fy = V->K[4 * ic + 1], cx = V->K[4 * ic + 2], cy = V->K[4 * ic + 3];
        const float* b = KF->bounds + 4 * ic;
        for (int i = 0; i < P->n; i++) {
            int32_t& okp = best_kp[(size_t)ic * P->n + i];
            int32_t& od = best_dist[(size_t)ic * P->n + i];
            okp = -1; od = 256;
            if (!P->valid[i]) continue;
            const float* X = P->pos + 3 * (size_t)i;
            const float p0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
            const float p1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
            const float p2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
            if (p2 < 0.0f) continue;
            const float invz = sim3 ? (float)(1.0 / p2) : 1 / p2;
            const float x = p0 * invz, y = p1 * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            if (!(u >= b[0] && u < b[1] && v >= b[2] && v < b[3])) continue;
            const float maxDistance = 1.2f * P->max_dist[i], minDistance = 0.8f * P->min_dist[i];
            const float PO[3] = {X[0] - Os[0], X[1] - Os[1], X[2] - Os[2]};
            const float dist3D = norm3(PO);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            if (dot3(PO, P->normal + 3 * (size_t)i) < 0.5 * dist3D) continue;
            const int nPredictedLevel = predict_scale(P->max_dist[i], dist3D, V->log_scale_factor, KF->n_levels);
            const float radius = th * KF->scale_factors[nPredictedLevel];
            const std::vector<int> vIndicesLocal = G.kf_features_in_area(ic, u, v, radius);
            if (vIndicesLocal.empty()) continue;
            int bestDist = 256, bestIdxglobal = -1;                     // (INT_MAX in the Sim3 variant :1652; same decisions)
            for (int idxLocal : vIndicesLocal) {
                const int idxglobal = G.first[ic] + idxLocal;
                const orc_keypoint_t& kp = KF->kps_un[idxglobal];
                const int kpLevel = kp.octave;
                if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
                if (!sim3) {
                    const float ex = u - kp.x, ey = v - kp.y;
                    const float e2 = ex * ex + ey * ey;
                    const float sigma2 = KF->scale_factors[kpLevel] * KF->scale_factors[kpLevel];        // mvLevelSigma2 / mvInvLevelSigma2
                    const float invSigma2 = 1.0f / sigma2;                                                // src/ORBextractor.cc:423-431
                    if (e2 * invSigma2 > 5.99) continue;
                }
                const int dist = descriptor_distance(P->desc + 32 * (size_t)i, KF->desc + 32 * (size_t)idxglobal);
                if (dist < bestDist) { bestDist = dist; bestIdxglobal = idxglobal; }
            }
            okp = bestIdxglobal; od = bestDist;
        }
    }
}

// SearchByBoWCrossCam(pKF1, c1, pKF2, c2, vpMatches12)  src/ORBmatcher.cc:297-414.  mp_valid* are indexed globally
// (pMP && !isBad()).  matches12[camera-local KF1 index] receives the GLOBAL KF2 key point index whose map point is
// vpMatches12[idx1local], -1 otherwise.
int orc_search_by_bow_kf(const orc_bowside_t* K1, int c1, const orc_bowside_t* K2, int c2, const uint8_t* mp_valid1, const uint8_t* mp_valid2,
                         float nnratio, int check_ori, int32_t* matches12) {
    int first1 = 0, first2 = 0;
    for (int c = 0; c < c1; c++) first1 += K1->n_kp[c];
    for (int c = 0; c < c2; c++) first2 += K2->n_kp[c];
    for (int i = 0; i < K1->n_kp[c1]; i++) matches12[i] = -1;
    std::vector<bool> vbMatched2(K2->n_kp[c2], false);
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    int it1 = K1->node_first[c1], end1 = K1->node_first[c1 + 1], it2 = K2->node_first[c2], end2 = K2->node_first[c2 + 1];
    while (it1 != end1 && it2 != end2) {
        if (K1->node_id[it1] == K2->node_id[it2]) {
            for (int a = K1->node_off[it1]; a < K1->node_off[it1 + 1]; a++) {
                const int idx1local = K1->idx[a], idx1global = first1 + idx1local;
                if (!mp_valid1[idx1global]) continue;
                const uint8_t* d1 = K1->desc + 32 * (size_t)idx1global;
                int bestDist1 = 256, bestIdx2local = -1, bestDist2 = 256;
                for (int e = K2->node_off[it2]; e < K2->node_off[it2 + 1]; e++) {
                    const int idx2local = K2->idx[e], idx2global = first2 + idx2local;
                    if (vbMatched2[idx2local] || !mp_valid2[idx2global]) continue;
                    const int dist = descriptor_distance(d1, K2->desc + 32 * (size_t)idx2global);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2local = idx2local; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        matches12[idx1local] = first2 + bestIdx2local;
                        vbMatched2[bestIdx2local] = true;
                        if (check_ori) rotHist[rot_bin(K1->angle[idx1global] - K2->angle[first2 + bestIdx2local])].push_back(idx1local);
                        nmatches++;
                    }
                }
            }
            it1++; it2++;
        } else if (K1->node_id[it1] < K2->node_id[it2]) {
            while (it1 != end1 && K1->node_id[it1] < K2->node_id[it2]) it1++;
        } else {
            while (it2 != end2 && K2->node_id[it2] < K1->node_id[it1]) it2++;
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int l : rotHist[i]) { matches12[l] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, camS)  src/ORBmatcher.cc:1253-1427 with CheckDistEpipolarLine :74-92.
// has_mp* (global) = GetMapPoint(g) != NULL; kps* are the global undistorted key point arrays; F12 row-major 3x3;
// C1sw / R2sw / t2sw / K2 give the epipole (:1261-1268).  matches12[camera-local KF1 index] = camera-local KF2 index or -1.
int orc_search_for_triangulation(const orc_bowside_t* K1, const orc_bowside_t* K2, int camS, const orc_keypoint_t* kps1, const orc_keypoint_t* kps2,
                                 const uint8_t* has_mp1, const uint8_t* has_mp2, const float* F12, const float* C1sw, const float* R2sw,
                                 const float* t2sw, const float* K2cam, const float* scale_factors, int check_ori, int32_t* matches12) {
    int first1 = 0, first2 = 0;
    for (int c = 0; c < camS; c++) { first1 += K1->n_kp[c]; first2 += K2->n_kp[c]; }
    const float C0 = R2sw[0] * C1sw[0] + R2sw[1] * C1sw[1] + R2sw[2] * C1sw[2] + t2sw[0];
    const float C1 = R2sw[3] * C1sw[0] + R2sw[4] * C1sw[1] + R2sw[5] * C1sw[2] + t2sw[1];
    const float C2 = R2sw[6] * C1sw[0] + R2sw[7] * C1sw[1] + R2sw[8] * C1sw[2] + t2sw[2];
    const float invz = 1.0f / C2;
    const float ex = K2cam[0] * C0 * invz + K2cam[2];
    const float ey = K2cam[1] * C1 * invz + K2cam[3];
    int nmatches = 0;
    std::vector<bool> vbMatched2(K2->n_kp[camS], false);
    for (int i = 0; i < K1->n_kp[camS]; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int it1 = K1->node_first[camS], end1 = K1->node_first[camS + 1], it2 = K2->node_first[camS], end2 = K2->node_first[camS + 1];
    while (it1 != end1 && it2 != end2) {
        if (K1->node_id[it1] == K2->node_id[it2]) {
            for (int a = K1->node_off[it1]; a < K1->node_off[it1 + 1]; a++) {
                const int local1 = K1->idx[a], global1 = first1 + local1;
                if (has_mp1[global1]) continue;
                const orc_keypoint_t& kp1 = kps1[global1];
                const uint8_t* d1 = K1->desc + 32 * (size_t)global1;
                int bestDist = TH_LOW, bestlocalIdx2 = -1;
                for (int e = K2->node_off[it2]; e < K2->node_off[it2 + 1]; e++) {
                    const int local2 = K2->idx[e], global2 = first2 + local2;
                    if (vbMatched2[local2] || has_mp2[global2]) continue;
                    const int dist = descriptor_distance(d1, K2->desc + 32 * (size_t)global2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const orc_keypoint_t& kp2 = kps2[global2];
                    const float distex = ex - kp2.x, distey = ey - kp2.y;
                    if (distex * distex + distey * distey < 100 * scale_factors[kp2.octave]) continue;
                    // CheckDistEpipolarLine :74-92
                    const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                    const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                    const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                    const float num = la * kp2.x + lb * kp2.y + lc;
                    const float den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    const float sigma2 = scale_factors[kp2.octave] * scale_factors[kp2.octave];     // mvLevelSigma2
                    if (dsqr < 3.84 * sigma2) { bestlocalIdx2 = local2; bestDist = dist; }
                }
                if (bestlocalIdx2 >= 0) {
                    matches12[local1] = bestlocalIdx2;
                    vbMatched2[bestlocalIdx2] = true;
                    nmatches++;
                    if (check_ori) rotHist[rot_bin(kp1.angle - kps2[first2 + bestlocalIdx2].angle)].push_back(local1);
                }
            }
            it1++; it2++;
        } else if (K1->node_id[it1] < K2->node_id[it2]) {
            while (it1 != end1 && K1->node_id[it1] < K2->node_id[it2]) it1++;
        } else {
            while (it2 != end2 && K2->node_id[it2] < K1->node_id[it1]) it2++;
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int l : rotHist[i]) { matches12[l] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // extern "C"
