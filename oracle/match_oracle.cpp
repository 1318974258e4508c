// match_oracle.cpp -- CPU oracle of the Hamming searches of ORBmatcher and of the Frame helpers under them.
//
// TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  Dependency-free restatement with the reference's own loop structure
// (std::vector grid cells, sequential claims); every function cites the lines it follows (paths relative to
// /root/reference).  PARITY UNPINNED by the reference: it has no tests or fixtures for these functions.
// Determinisations: the 3x3 * 3x1 float product of SearchByProjectionOnCam (cv::Mat operator*) is
// evaluated left-to-right in FP32 without FMA; PredictScale's log() is evaluated in double.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orb_oracle.h"

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;   // src/ORBmatcher.cc:57-59
const int GRID_COLS = 64, GRID_ROWS = 48;                  // include/Frame.h:39-40

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

int cv_round_f(float v) { return (int)lrintf(v); }

// Frame's per-camera keypoint grid (src/Frame.cc:156-159,172-196,380-390)
struct FrameView {
    const orc_frame_t* F;
    std::vector<int> first;                                        // global index of the first keypoint of camera c
    std::vector<float> invW, invH;
    std::vector<std::vector<std::vector<std::vector<int>>>> grid;  // [c][64][48] -> camera-local indices

    explicit FrameView(const orc_frame_t* f) : F(f) {
        const int C = f->n_cams;
        first.assign(C + 1, 0);
        for (int c = 0; c < C; c++) first[c + 1] = first[c] + f->n_kp[c];
        invW.resize(C); invH.resize(C); grid.resize(C);
        for (int c = 0; c < C; c++) {
            const float* b = f->bounds + 4 * c;   // minX maxX minY maxY
            invW[c] = (float)GRID_COLS / (float)(b[1] - b[0]);
            invH[c] = (float)GRID_ROWS / (float)(b[3] - b[2]);
            grid[c].assign(GRID_COLS, std::vector<std::vector<int>>(GRID_ROWS));
            for (int i = 0; i < f->n_kp[c]; i++) {
                const orc_keypoint_t& kp = f->kps_un[first[c] + i];
                const int px = cv_round_f((kp.x - b[0]) * invW[c]);           // PosInGrid :380-390
                const int py = cv_round_f((kp.y - b[2]) * invH[c]);
                if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
                grid[c][px][py].push_back(i);
            }
        }
    }

    // Frame::GetFeaturesInArea  src/Frame.cc:316-376 (camera-local indices)
    std::vector<int> features_in_area(int c, float x, float y, float r, int minLevel, int maxLevel) const {
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
                    const float distx = kp.x - x, disty = kp.y - y;
                    if (std::fabs(distx) < r && std::fabs(disty) < r) v.push_back(local);
                }
        return v;
    }
};

// ORBmatcher::ComputeThreeMaxima  src/ORBmatcher.cc:1969-2010
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
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

int rot_bin(float rot) {   // src/ORBmatcher.cc:243-248, 1070-1075
    const float factor = 1.0f / HISTO_LENGTH;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

}  // namespace

extern "C" {

// ORBmatcher::SearchByProjection(FramePtr, vector<MapPointPtr>, th)  src/ORBmatcher.cc:539-624.
// blocked[g] = pF->mvpMapPoints[g] && Observations()>0 on entry; kp_to_mp[g] receives the index of the map point written into
// mvpMapPoints[g] (:618), untouched otherwise.  Returns nmatches.
int orc_search_by_projection(const orc_frame_t* F, const orc_mp_t* mps, int n, float th, float nnratio, const uint8_t* blocked_in,
                             int32_t* kp_to_mp) {
    FrameView V(F);
    const int totalN = V.first[F->n_cams];
    std::vector<uint8_t> blocked(blocked_in, blocked_in + totalN);
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int i = 0; i < n; i++) {
        const orc_mp_t& mp = mps[i];
        if (!mp.valid) continue;                                    // !pMP || !mbTrackInView || isBad()  :550-552
        const int lvl = mp.level;
        float r = mp.view_cos > 0.998 ? 2.5 : 4.0;                  // RadiusByViewingCos :65-71
        if (bFactor) r *= th;
        const std::vector<int> vIndices = V.features_in_area(mp.cam, mp.u, mp.v, r * F->scale_factors[lvl], lvl - 1, lvl + 1);
        if (vIndices.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int local : vIndices) {
            const int g = V.first[mp.cam] + local;
            if (blocked[g]) continue;                               // :589-591
            const int dist = descriptor_distance(mp.desc, F->desc + 32 * (size_t)g);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F->kps_un[g].octave; bestIdx = g; }
            else if (dist < bestDist2) { bestLevel2 = F->kps_un[g].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            kp_to_mp[bestIdx] = i;                                  // pF->mvpMapPoints[bestglobalIdx] = pMP
            blocked[bestIdx] = mp.obs_positive ? 1 : 0;
            nmatches++;
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(cur, last, th, bMapScaled) -> SearchByProjectionOnCam  src/ORBmatcher.cc:634-690, 954-1113.
// kp_to_last[g] receives the last-frame global keypoint index whose map point is written into cur->mvpMapPoints[g], -1 when the
// rotation check removes it again; blocked as above.  per_cam (may be NULL) receives the per-camera match counts.
int orc_search_by_projection_last(const orc_frame_t* cur, const float* Rsw, const float* tsw, const float* K, const orc_lastframe_t* last,
                                  float th, int check_ori, int map_scaled, const uint8_t* blocked_in, int32_t* kp_to_last, int32_t* per_cam) {
    FrameView V(cur);
    const int totalN = V.first[cur->n_cams];
    std::vector<uint8_t> blocked(blocked_in, blocked_in + totalN);
    int nmatches = 0;
    for (int ic = 0; ic < cur->n_cams; ic++) {
        if (per_cam) per_cam[ic] = 0;
        if (ic != 0 && !map_scaled) continue;
        // ---- SearchByProjectionOnCam
        int nmatch = 0;
        const float* R = Rsw + 9 * ic;
        const float* t = tsw + 3 * ic;
        const float fx = K[4 * ic], fy = K[4 * ic + 1], cx = K[4 * ic + 2], cy = K[4 * ic + 3];
        const float* b = cur->bounds + 4 * ic;
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int i = 0; i < last->n; i++) {
            if (last->cam[i] != ic) continue;
            if (!last->valid[i]) continue;
            const float* X = last->pos + 3 * (size_t)i;
            const float xs = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
            const float ys = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
            const float zs = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
            if (zs < 0) continue;
            const float invzs = (float)(1.0 / zs);
            const float u = fx * xs * invzs + cx;
            const float v = fy * ys * invzs + cy;
            if (u < b[0] || u > b[1]) continue;
            if (v < b[2] || v > b[3]) continue;
            const int nLastOctave = last->octave[i];
            const float radius = th * cur->scale_factors[nLastOctave];
            const std::vector<int> vIndices = V.features_in_area(ic, u, v, radius, nLastOctave - 1, nLastOctave + 1);
            if (vIndices.empty()) continue;
            int bestDist = 256, bestIdx = -1;
            for (int local : vIndices) {
                const int g = V.first[ic] + local;
                if (blocked[g]) continue;
                const int dist = descriptor_distance(last->desc + 32 * (size_t)i, cur->desc + 32 * (size_t)g);
                if (dist < bestDist) { bestDist = dist; bestIdx = g; }
            }
            if (bestDist <= TH_HIGH) {
                kp_to_last[bestIdx] = i;
                blocked[bestIdx] = last->obs_positive[i] ? 1 : 0;
                nmatch++;
                if (check_ori) rotHist[rot_bin(last->angle[i] - cur->kps_un[bestIdx].angle)].push_back(bestIdx);
            }
        }
        if (check_ori) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++)
                if (i != ind1 && i != ind2 && i != ind3)
                    for (int g : rotHist[i]) { kp_to_last[g] = -1; blocked[g] = 0; nmatch--; }
        }
        if (per_cam) per_cam[ic] = nmatch;
        if (nmatch <= 20) { nmatches = nmatch; break; }      // :664-667 (upstream quirk: replaces the running total)
        nmatches += nmatch;
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(pF, pKF, vpMapPointMatches, bMapScaled) -> SearchByBoWCrossCam(pF, ic, pKF, ic, ...)
// src/ORBmatcher.cc:102-148, 162-294.  Feature vectors are CSR: node ids ascending (std::map order), camera-local indices.
// f_to_kf[global F keypoint] receives the global KF keypoint index whose map point is matched, -1 otherwise.
int orc_search_by_bow(const orc_bowside_t* F, const orc_bowside_t* KF, const uint8_t* kf_mp_valid, float nnratio, int check_ori, int map_scaled,
                      int32_t* f_to_kf) {
    const int C = KF->n_cams;
    std::vector<int> firstF(C + 1, 0), firstK(C + 1, 0);
    for (int c = 0; c < C; c++) { firstF[c + 1] = firstF[c] + F->n_kp[c]; firstK[c + 1] = firstK[c] + KF->n_kp[c]; }
    for (int g = 0; g < firstF[C]; g++) f_to_kf[g] = -1;
    int nmatches = 0;
    for (int ic = 0; ic < C; ic++) {
        if (ic != 0 && !map_scaled) continue;
        std::vector<int> inner(F->n_kp[ic], -1);            // vpInnerMPMatches (camera-local)
        int nm = 0;
        std::vector<int> rotHist[HISTO_LENGTH];
        int kf = KF->node_first[ic], kfEnd = KF->node_first[ic + 1], ff = F->node_first[ic], ffEnd = F->node_first[ic + 1];
        while (kf != kfEnd && ff != ffEnd) {
            if (KF->node_id[kf] == F->node_id[ff]) {
                for (int a = KF->node_off[kf]; a < KF->node_off[kf + 1]; a++) {
                    const int idxKF = KF->idx[a];
                    const int gKF = firstK[ic] + idxKF;
                    if (!kf_mp_valid[gKF]) continue;          // !pMP || isBad()
                    const uint8_t* dKF = KF->desc + 32 * (size_t)gKF;
                    int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                    for (int q = F->node_off[ff]; q < F->node_off[ff + 1]; q++) {
                        const int idxF = F->idx[q];
                        if (inner[idxF] >= 0) continue;       // :216
                        const int dist = descriptor_distance(dKF, F->desc + 32 * (size_t)(firstF[ic] + idxF));
                        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = idxF; }
                        else if (dist < bestDist2) bestDist2 = dist;
                    }
                    if (bestDist1 <= TH_LOW) {
                        if ((float)bestDist1 < nnratio * (float)bestDist2) {
                            inner[bestIdxF] = gKF;
                            if (check_ori) rotHist[rot_bin(KF->angle[gKF] - F->angle[firstF[ic] + bestIdxF])].push_back(bestIdxF);
                            nm++;
                        }
                    }
                }
                kf++; ff++;
            } else if (KF->node_id[kf] < F->node_id[ff]) {
                while (kf != kfEnd && KF->node_id[kf] < F->node_id[ff]) kf++;      // lower_bound
            } else {
                while (ff != ffEnd && F->node_id[ff] < KF->node_id[kf]) ff++;
            }
        }
        if (check_ori) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (int l : rotHist[i]) { inner[l] = -1; nm--; }
            }
        }
        nmatches += nm;
        for (int l = 0; l < F->n_kp[ic]; l++) if (inner[l] >= 0) f_to_kf[firstF[ic] + l] = inner[l];
    }
    return nmatches;
}

// Frame::isInFrustum(pMP, viewingCosLimit, bForAllCam) + MapPoint::PredictScale  src/Frame.cc:244-312, src/MapPoint.cc:440-455.
// out[i] = {in_view, cam, level} ; uvc[i] = {u, v, viewCos}
void orc_is_in_frustum(const orc_frustum_t* Q, const float* pos, const float* normal, const float* max_dist, const float* min_dist, int n,
                       float cos_limit, int for_all_cams, int32_t* out, float* uvc) {
    for (int i = 0; i < n; i++) {
        out[3 * i] = 0; out[3 * i + 1] = -1; out[3 * i + 2] = 0;
        uvc[3 * i] = uvc[3 * i + 1] = uvc[3 * i + 2] = 0;
        const float* P = pos + 3 * (size_t)i;
        for (int ic = 0; ic < Q->n_cams; ic++) {
            if (ic != 0 && !for_all_cams) continue;
            const float* R = Q->Rsw + 9 * ic;
            const float* t = Q->tsw + 3 * ic;
            const float X = R[0] * P[0] + R[1] * P[1] + R[2] * P[2] + t[0];
            const float Y = R[3] * P[0] + R[4] * P[1] + R[5] * P[2] + t[1];
            const float Z = R[6] * P[0] + R[7] * P[1] + R[8] * P[2] + t[2];
            if (Z < 0.0f) continue;
            const float invz = 1.0f / Z;
            const float u = Q->K[4 * ic] * X * invz + Q->K[4 * ic + 2];
            const float v = Q->K[4 * ic + 1] * Y * invz + Q->K[4 * ic + 3];
            const float* b = Q->bounds + 4 * ic;
            if (u < b[0] || u > b[1]) continue;
            if (v < b[2] || v > b[3]) continue;
            const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];   // MapPoint::Get{Max,Min}DistanceInvariance
            const float* O = Q->Ow + 3 * ic;
            const float PO[3] = {P[0] - O[0], P[1] - O[1], P[2] - O[2]};
            const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);   // cv::norm
            if (dist < minDistance || dist > maxDistance) continue;
            const float* Pn = normal + 3 * (size_t)i;
            const float viewCos = (float)(((double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2]) / dist);
            if (viewCos < cos_limit) continue;
            const float ratio = max_dist[i] / dist;                                             // PredictScale
            int nScale = (int)std::ceil(std::log((double)ratio) / (double)Q->log_scale_factor);
            if (nScale < 0) nScale = 0;
            else if (nScale >= Q->n_levels) nScale = Q->n_levels - 1;
            out[3 * i] = 1; out[3 * i + 1] = ic; out[3 * i + 2] = nScale;
            uvc[3 * i] = u; uvc[3 * i + 1] = v; uvc[3 * i + 2] = viewCos;
            break;
        }
    }
}


// Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:410-442, 454-490) -> cv::undistortPoints(mat, mat, K, distCoef, Mat(), K).
// OpenCV is not vendored in the reference; this restates cvUndistortPointsInternal (calib3d/undistort: normalise, 5 fixed-point
// iterations of the inverse Brown model in double, re-project with P = K, round to float) and is PINNED bit-exactly against cv2 4.13
// (tests/test_oracle_cv2.py, tests/golden/undistort.npz).
void orc_undistort_points(const float* pts, int n, const float* K4, const float* dist, int n_dist, float* out) {
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    double k[14];
    for (int i = 0; i < 14; i++) k[i] = i < n_dist ? (double)dist[i] : 0.0;
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; i++) {
        const double u = pts[2 * i], v = pts[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1.0);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
}
// bounds = {mvMinX, mvMaxX, mvMinY, mvMaxY}
void orc_image_bounds(int width, int height, const float* K4, const float* dist, int n_dist, float* bounds) {
    if (n_dist > 0 && dist[0] != 0.0f) {
        const float c[8] = {0.f, 0.f, (float)width, 0.f, 0.f, (float)height, (float)width, (float)height};
        float o[8];
        orc_undistort_points(c, 4, K4, dist, n_dist, o);
        bounds[0] = std::min(o[0], o[4]); bounds[1] = std::max(o[2], o[6]);
        bounds[2] = std::min(o[1], o[3]); bounds[3] = std::max(o[5], o[7]);
    } else {
        bounds[0] = 0.f; bounds[1] = (float)width; bounds[2] = 0.f; bounds[3] = (float)height;
    }
}

}  // extern "C"
