// orb_search.cu -- sm_100a guided Hamming searches behind orbm_search_by_projection / _last / _by_bow / orbm_is_in_frustum.
//
// Reference path (file:line under /root/reference): ORBmatcher::SearchByProjection(F, MPs, th) src/ORBmatcher.cc:539-624,
// SearchByProjection(F, lastF, th, scaled) -> SearchByProjectionOnCam :634-690, :954-1113, SearchByBoW -> SearchByBoWCrossCam
// :102-294, ComputeThreeMaxima :1969-2010, Frame::GetFeaturesInArea / PosInGrid src/Frame.cc:316-390, Frame::isInFrustum
// src/Frame.cc:244-312 + MapPoint::PredictScale src/MapPoint.cc:440-455.
//
// The reference's loops are order dependent: a keypoint that received a map point is skipped by the map points that
// follow.  The device formulation separates the two parts:
//   k_grid_build     CTA per camera: the 64x48 grid of Frame::Frame as a CSR (cells column-major like mvGrids[c][ix][iy],
//                    indices ascending inside a cell = insertion order)
//   k_window_cands   warp per query (map point): cells of the window in GetFeaturesInArea's order, level / window
//                    predicates, 256-bit Hamming distance of every survivor -> packed candidate list in reference order
//                    (two passes: count, then fill at the scanned offsets)
//   k_bow_cands      warp per key-frame feature of a shared vocabulary node: distances to the frame's features of that node
//   k_resolve        ONE warp per claim sequence: walks the queries in input order, lanes scan the candidate list against the
//                    claim bitmap in shared memory, warp-min of packed (distance, position) keys = the reference's
//                    first-wins best / second-best, acceptance gates, claim; then the rotation-histogram consistency pass
// There is no CPU fallback.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "orb_common.h"

#define GRID_COLS ORBM_GRID_COLS
#define GRID_ROWS ORBM_GRID_ROWS
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define MAX_CAMS 8
#define KP_BITS 17                       // global keypoint index < 131072
#define KP_MASK ((1u << KP_BITS) - 1u)
#define HISTO_LENGTH ORBM_HISTO_LENGTH

struct FrameDev {
    int n_cams, n_levels, totalN;
    int first[MAX_CAMS + 1];
    float minX[MAX_CAMS], maxX[MAX_CAMS], minY[MAX_CAMS], maxY[MAX_CAMS], invW[MAX_CAMS], invH[MAX_CAMS];
    float scale[16];
    const float4* kp;          // x, y, angle, octave (int bits)
    const uint4* desc;         // 2 x uint4 per keypoint
    int* cell_off;             // [n_cams][GRID_CELLS + 1]
    int* cell_idx;             // [totalN] camera-local indices, camera c at first[c]
    int kf_quirk;              // KeyFrame::GetFeaturesInArea as upstream: the window test reads mvTotalKeysUn[camera-LOCAL index] (src/KeyFrame.cc:757)
};

struct Query {                 // one window search
    int cam, valid;
    float u, v, r;
    int minLevel, maxLevel;
    int tag;                   // what the accepted keypoint is labelled with (map-point index / last-frame keypoint index)
    int obs_positive;
    float angle;               // for the rotation histogram (mode B)
    uint32_t desc[8];
};

__device__ __forceinline__ int hamming256(const uint32_t* q, uint4 a, uint4 b) {
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) + __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) +
           __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

// ------------------------------------------------------------------------------------------------ grid
// Frame::PosInGrid (src/Frame.cc:380-390) + the fill loop (:179-196).  One CTA per camera.
__global__ void __launch_bounds__(256) k_grid_build(FrameDev F) {
    extern __shared__ int sm[];
    int* cnt = sm;                                   // GRID_CELLS + 1
    short* cell = reinterpret_cast<short*>(sm + GRID_CELLS + 1);
    const int c = blockIdx.x, tid = threadIdx.x;
    const int n = F.first[c + 1] - F.first[c];
    for (int i = tid; i <= GRID_CELLS; i += 256) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const float4 k = F.kp[F.first[c] + i];
        const int px = __float2int_rn(__fmul_rn(__fsub_rn(k.x, F.minX[c]), F.invW[c]));
        const int py = __float2int_rn(__fmul_rn(__fsub_rn(k.y, F.minY[c]), F.invH[c]));
        int id = -1;
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) { id = px * GRID_ROWS + py; atomicAdd(&cnt[id], 1); }
        cell[i] = (short)id;
    }
    __syncthreads();
    if (tid == 0) {                                  // exclusive scan (3072 cells)
        int s = 0;
        for (int i = 0; i < GRID_CELLS; i++) { const int v = cnt[i]; cnt[i] = s; s += v; }
        cnt[GRID_CELLS] = s;
    }
    __syncthreads();
    int* off = F.cell_off + (size_t)c * (GRID_CELLS + 1);
    for (int i = tid; i <= GRID_CELLS; i += 256) off[i] = cnt[i];
    for (int i = tid; i < n; i += 256) {
        const int id = cell[i];
        if (id < 0) continue;
        int rank = 0;
        for (int j = 0; j < i; j++) rank += cell[j] == id;   // insertion order inside the cell
        F.cell_idx[F.first[c] + cnt[id] + rank] = i;
    }
}

// ------------------------------------------------------------------------------------------------ window candidates
// Frame::GetFeaturesInArea (src/Frame.cc:316-376) for one query per warp.  rec = dist << 22 | octave << 17 | global keypoint.
template <bool FILL>
__global__ void __launch_bounds__(128) k_window_cands(FrameDev F, const Query* __restrict__ qs, int nq, int* __restrict__ q_cnt,
                                                      const int* __restrict__ q_off, uint32_t* __restrict__ recs) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const Query& Q = qs[q];
    int run = 0;
    if (Q.valid) {
        const int c = Q.cam;
        const float x = Q.u, y = Q.v, r = Q.r;
        const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.minX[c]), r), F.invW[c])));
        const int x1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.minX[c]), r), F.invW[c])));
        const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.minY[c]), r), F.invH[c])));
        const int y1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.minY[c]), r), F.invH[c])));
        if (x0 < GRID_COLS && x1 >= 0 && y0 < GRID_ROWS && y1 >= 0) {
            uint32_t d[8];
#pragma unroll
            for (int i = 0; i < 8; i++) d[i] = Q.desc[i];
            const int* off = F.cell_off + (size_t)c * (GRID_CELLS + 1);
            const int* idx = F.cell_idx + F.first[c];
            const int base = FILL ? q_off[q] : 0;
            for (int ix = x0; ix <= x1; ix++) {
                const int a = off[ix * GRID_ROWS + y0], b = off[ix * GRID_ROWS + y1 + 1];   // cells (ix, y0..y1) are contiguous
                for (int e0 = a; e0 < b; e0 += 32) {
                    const int e = e0 + lane;
                    bool pred = false;
                    uint32_t rec = 0;
                    if (e < b) {
                        const int g = F.first[c] + idx[e];
                        const float4 k = F.kp[g];
                        const float4 kpos = F.kf_quirk ? F.kp[idx[e]] : k;
                        const int oct = __float_as_int(k.w);
                        pred = oct >= Q.minLevel && oct <= Q.maxLevel && fabsf(__fsub_rn(kpos.x, x)) < r && fabsf(__fsub_rn(kpos.y, y)) < r;
                        if (pred && FILL) {
                            const uint4 da = F.desc[2 * (size_t)g], db = F.desc[2 * (size_t)g + 1];
                            rec = ((uint32_t)hamming256(d, da, db) << 22) | ((uint32_t)oct << KP_BITS) | (uint32_t)g;
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, pred);
                    if (FILL && pred) recs[base + run + __popc(m & ((1u << lane) - 1))] = rec;
                    run += __popc(m);
                }
            }
        }
    }
    if (!FILL && lane == 0) q_cnt[q] = run;
}

// exclusive scan of the candidate counts (one CTA); total -> *total
__global__ void __launch_bounds__(1024) k_scan(const int* __restrict__ cnt, int n, int* __restrict__ off, int* __restrict__ total) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? cnt[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wsum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = carry + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0) + s - v;
        if (i < n) off[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { off[n] = carry; *total = carry; }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------ key-frame searches: projection
// One thread per (camera, map point): the gates that precede the window search in SearchByProjectionOnCam(F, cam, KF, ...)
// (src/ORBmatcher.cc:838-873), SearchByProjection(KF, MPs, ...) (:712-745), SearchByProjection(KF, query, Scw, ...) (:452-490),
// Fuse (:1448-1484) and Fuse(Scw) (:1606-1643).  FP32 left to right, cv::norm / Mat::dot in double, as the oracle.
enum { PS_DEPTH_POS = 1, PS_NORMALISE_FIRST = 2, PS_HALF_OPEN = 4, PS_VIEW_ANGLE = 8, PS_LEVEL_UP = 16, PS_CHI2 = 32 };
struct ProjDev {
    int n, n_levels, flags, cam0;            // query q: camera cam0 + q / n, map point q % n
    float th, log_scale;
    float R[MAX_CAMS][9], t[MAX_CAMS][3], Ow[MAX_CAMS][3], K[MAX_CAMS][4], b[MAX_CAMS][4];
    float scale[16];
    const uint8_t* valid; const float* pos; const float* normal; const float* max_dist; const float* min_dist; const uint4* desc; const float* angle;
};
__global__ void __launch_bounds__(128) k_project_queries(ProjDev P, int nq, Query* __restrict__ qs, int* __restrict__ q_seq, int* __restrict__ q_tag,
                                                         uint8_t* __restrict__ q_obs, float* __restrict__ q_angle) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int i = q % P.n, c = P.cam0 + q / P.n;
    Query Q;
    Q.cam = c; Q.valid = 0; Q.u = Q.v = Q.r = 0.f; Q.minLevel = Q.maxLevel = 0; Q.tag = i; Q.obs_positive = 1; Q.angle = P.angle ? P.angle[i] : 0.f;
    const uint4 da = P.desc[2 * (size_t)i], db = P.desc[2 * (size_t)i + 1];
    Q.desc[0] = da.x; Q.desc[1] = da.y; Q.desc[2] = da.z; Q.desc[3] = da.w; Q.desc[4] = db.x; Q.desc[5] = db.y; Q.desc[6] = db.z; Q.desc[7] = db.w;
    do {
        if (!P.valid[i]) break;
        const float* R = P.R[c];
        const float P0 = P.pos[3 * (size_t)i], P1 = P.pos[3 * (size_t)i + 1], P2 = P.pos[3 * (size_t)i + 2];
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], P0), __fmul_rn(R[1], P1)), __fmul_rn(R[2], P2)), P.t[c][0]);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], P0), __fmul_rn(R[4], P1)), __fmul_rn(R[5], P2)), P.t[c][1]);
        const float Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], P0), __fmul_rn(R[7], P1)), __fmul_rn(R[8], P2)), P.t[c][2]);
        if ((P.flags & PS_DEPTH_POS) && Z < 0.0f) break;
        const float invz = __fdiv_rn(1.0f, Z);
        float u, v;
        if (P.flags & PS_NORMALISE_FIRST) {
            u = __fadd_rn(__fmul_rn(P.K[c][0], __fmul_rn(X, invz)), P.K[c][2]);
            v = __fadd_rn(__fmul_rn(P.K[c][1], __fmul_rn(Y, invz)), P.K[c][3]);
        } else {
            u = __fadd_rn(__fmul_rn(__fmul_rn(P.K[c][0], X), invz), P.K[c][2]);
            v = __fadd_rn(__fmul_rn(__fmul_rn(P.K[c][1], Y), invz), P.K[c][3]);
        }
        if (P.flags & PS_HALF_OPEN) { if (!(u >= P.b[c][0] && u < P.b[c][1] && v >= P.b[c][2] && v < P.b[c][3])) break; }
        else { if (u < P.b[c][0] || u > P.b[c][1]) break; if (v < P.b[c][2] || v > P.b[c][3]) break; }
        const float PO0 = __fsub_rn(P0, P.Ow[c][0]), PO1 = __fsub_rn(P1, P.Ow[c][1]), PO2 = __fsub_rn(P2, P.Ow[c][2]);
        const float dist = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn((double)PO0, (double)PO0), __dmul_rn((double)PO1, (double)PO1)), __dmul_rn((double)PO2, (double)PO2)));
        const float maxDistance = __fmul_rn(1.2f, P.max_dist[i]), minDistance = __fmul_rn(0.8f, P.min_dist[i]);
        if (dist < minDistance || dist > maxDistance) break;
        if (P.flags & PS_VIEW_ANGLE) {
            const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)PO0, (double)P.normal[3 * (size_t)i]), __dmul_rn((double)PO1, (double)P.normal[3 * (size_t)i + 1])),
                                         __dmul_rn((double)PO2, (double)P.normal[3 * (size_t)i + 2]));
            if (dot < __dmul_rn(0.5, (double)dist)) break;
        }
        const float ratio = __fdiv_rn(P.max_dist[i], dist);                       // MapPoint::PredictScale  src/MapPoint.cc:423-455
        int nScale = (int)ceil(log((double)ratio) / (double)P.log_scale);
        if (nScale < 0) nScale = 0;
        else if (nScale >= P.n_levels) nScale = P.n_levels - 1;
        Q.valid = 1; Q.u = u; Q.v = v; Q.r = __fmul_rn(P.th, P.scale[nScale]);
        Q.minLevel = nScale - 1; Q.maxLevel = (P.flags & PS_LEVEL_UP) ? nScale + 1 : nScale;
    } while (false);
    qs[q] = Q;
    if (q_seq) { q_seq[q] = Q.valid ? 0 : -1; q_tag[q] = i; q_obs[q] = 1; q_angle[q] = Q.angle; }
}

// window search without claims: the key point with the smallest distance (first wins) for every query, one warp per query
// (the inner loops of src/ORBmatcher.cc:753-775, 1492-1527, 1654-1668).  chi2: Fuse's pixel gate e2 * invSigma2[level] > 5.99 (:1510-1517).
__global__ void __launch_bounds__(128) k_window_best(FrameDev F, const Query* __restrict__ qs, int nq, int chi2, int* __restrict__ best_kp,
                                                     int* __restrict__ best_dist) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const Query& Q = qs[q];
    unsigned long long best = ~0ull;
    if (Q.valid) {
        const int c = Q.cam;
        const float x = Q.u, y = Q.v, r = Q.r;
        const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.minX[c]), r), F.invW[c])));
        const int x1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.minX[c]), r), F.invW[c])));
        const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.minY[c]), r), F.invH[c])));
        const int y1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.minY[c]), r), F.invH[c])));
        if (x0 < GRID_COLS && x1 >= 0 && y0 < GRID_ROWS && y1 >= 0) {
            uint32_t d[8];
#pragma unroll
            for (int i = 0; i < 8; i++) d[i] = Q.desc[i];
            const int* off = F.cell_off + (size_t)c * (GRID_CELLS + 1);
            const int* idx = F.cell_idx + F.first[c];
            unsigned pos0 = 0;
            for (int ix = x0; ix <= x1; ix++) {
                const int a = off[ix * GRID_ROWS + y0], b = off[ix * GRID_ROWS + y1 + 1];
                for (int e = a + lane; e < b; e += 32) {
                    const int g = F.first[c] + idx[e];
                    const float4 k = F.kp[g];
                    const float4 kpos = F.kf_quirk ? F.kp[idx[e]] : k;
                    const int oct = __float_as_int(k.w);
                    if (!(fabsf(__fsub_rn(kpos.x, x)) < r && fabsf(__fsub_rn(kpos.y, y)) < r)) continue;
                    if (oct < Q.minLevel || oct > Q.maxLevel) continue;
                    if (chi2) {
                        const float ex = __fsub_rn(x, k.x), ey = __fsub_rn(y, k.y);
                        const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                        const float inv = __fdiv_rn(1.0f, __fmul_rn(F.scale[oct], F.scale[oct]));
                        if ((double)__fmul_rn(e2, inv) > 5.99) continue;
                    }
                    const unsigned long long key = ((unsigned long long)hamming256(d, F.desc[2 * (size_t)g], F.desc[2 * (size_t)g + 1]) << 48) |
                                                   ((unsigned long long)(pos0 + (unsigned)(e - a)) << 24) | (unsigned)g;
                    best = key < best ? key : best;
                }
                pos0 += (unsigned)(b - a);
            }
        }
    }
    best = warp_min_u64(best);
    if (lane == 0) { best_kp[q] = best == ~0ull ? -1 : (int)(best & KP_MASK); best_dist[q] = best == ~0ull ? 256 : (int)(best >> 48); }
}

// ------------------------------------------------------------------------------------------------ BoW candidates
struct BowQuery { int gKF, f_begin, f_end, cam, firstF; float angle; };   // f_begin..f_end: range in the target side's idx array
// gates of SearchForTriangulation (src/ORBmatcher.cc:1318-1335 + CheckDistEpipolarLine :74-92), evaluated per candidate
struct TriDev { int on; float ex, ey; float F12[9]; float scale[16]; const float4* kp1; const float4* kp2; };
#define REC_INVALID 0x3ffu
__global__ void __launch_bounds__(128) k_bow_cands(const BowQuery* __restrict__ qs, int nq, const int* __restrict__ q_off, const uint4* __restrict__ descKF,
                                                   const uint4* __restrict__ descF, const int* __restrict__ idxF, uint32_t* __restrict__ recs,
                                                   const uint8_t* __restrict__ t_skip, TriDev T) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const BowQuery Q = qs[q];
    const uint4 qa = descKF[2 * (size_t)Q.gKF], qb = descKF[2 * (size_t)Q.gKF + 1];
    const uint32_t d[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    const int base = q_off[q];
    float la = 0.f, lb = 0.f, lc = 0.f, den = 0.f;
    if (T.on) {
        const float4 k1 = T.kp1[Q.gKF];
        la = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, T.F12[0]), __fmul_rn(k1.y, T.F12[3])), T.F12[6]);
        lb = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, T.F12[1]), __fmul_rn(k1.y, T.F12[4])), T.F12[7]);
        lc = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, T.F12[2]), __fmul_rn(k1.y, T.F12[5])), T.F12[8]);
        den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    }
    for (int e = Q.f_begin + lane; e < Q.f_end; e += 32) {
        const int local = idxF[e];
        const size_t g = (size_t)Q.firstF + local;
        uint32_t dist = (uint32_t)hamming256(d, descF[2 * g], descF[2 * g + 1]);
        if (t_skip && t_skip[g]) dist = REC_INVALID;
        if (T.on && dist != REC_INVALID) {
            const float4 k2 = T.kp2[g];
            const int oct = __float_as_int(k2.w);
            const float dx = __fsub_rn(T.ex, k2.x), dy = __fsub_rn(T.ey, k2.y);
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, k2.x), __fmul_rn(lb, k2.y)), lc);
            bool ok = dist <= (uint32_t)ORBM_TH_LOW;
            ok = ok && !(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, T.scale[oct]));
            ok = ok && den != 0.0f;
            if (ok) {
                const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                ok = (double)dsqr < __dmul_rn(3.84, (double)__fmul_rn(T.scale[oct], T.scale[oct]));
            }
            if (!ok) dist = REC_INVALID;
        }
        recs[base + (e - Q.f_begin)] = (dist << 22) | (uint32_t)local;
    }
}

// ------------------------------------------------------------------------------------------------ resolve
enum { MODE_MP = 0, MODE_LAST = 1, MODE_BOW = 2, MODE_BOW_KF = 3, MODE_TRI = 4 };
struct ResolveArgs {
    int mode, nq, n_bits;            // n_bits: size of the claim bitmap (keypoints)
    const int* q_off;
    const uint32_t* recs;
    const int* q_seq;                // sequence (camera) of every query, -1 = not searched
    const int* q_tag;
    const uint8_t* q_obs;            // obs_positive per query (modes MP / LAST)
    const float* q_angle;            // query keypoint angle (modes LAST / BOW)
    const float* kp_angle;           // target keypoint angle, indexed like the claim bitmap (+ seq_base)
    const int* seq_base;             // mode BOW: first global keypoint of the camera (targets are camera-local)
    const uint8_t* blocked_in;       // may be NULL (mode BOW)
    int* out;                        // per target keypoint (global index)
    int* seq_matches;                // per sequence
    int* m_kp; int* m_bin;           // match lists for the rotation check, per sequence slices of nq entries
    float nnratio;
    int check_ori;
    int th_accept;                   // mode LAST: accept iff best <= th_accept (TH_HIGH, or the caller's ORBdist / TH_LOW in the key-frame variants)
};


__global__ void __launch_bounds__(32) k_resolve(ResolveArgs R) {
    extern __shared__ unsigned s_claim[];          // bitmap over target keypoints
    __shared__ int s_hist[HISTO_LENGTH];
    const int seq = blockIdx.x, lane = threadIdx.x;
    const int words = (R.n_bits + 31) >> 5;
    for (int w = lane; w < words; w += 32) {
        unsigned m = 0;
        if (R.blocked_in)
            for (int b = 0; b < 32 && w * 32 + b < R.n_bits; b++) m |= (R.blocked_in[w * 32 + b] ? 1u : 0u) << b;
        s_claim[w] = m;
    }
    if (lane < HISTO_LENGTH) s_hist[lane] = 0;
    __syncwarp();
    const unsigned long long NONE = ~0ull;
    const int tbase = R.seq_base ? R.seq_base[seq] : 0;
    int nmatch = 0, nlist = 0;
    int* m_kp = R.m_kp + (size_t)seq * R.nq;
    int* m_bin = R.m_bin + (size_t)seq * R.nq;
    for (int q = 0; q < R.nq; q++) {
        if (R.q_seq[q] != seq) continue;
        const int off = R.q_off[q], cnt = R.q_off[q + 1] - off;
        if (cnt == 0) continue;
        // key = dist << 48 | position << 24 | payload (octave << 17 | keypoint): min = first-wins best, second min = second best
        unsigned long long b1 = NONE, b2 = NONE;
        for (int pos = lane; pos < cnt; pos += 32) {
            const uint32_t rec = R.recs[off + pos];
            const uint32_t kp = rec & KP_MASK;
            if ((rec >> 22) > 256u) continue;                          // masked / gated out by the candidate kernel
            if ((s_claim[kp >> 5] >> (kp & 31)) & 1u) continue;
            // SearchForTriangulation replaces the best on ties (`dist > bestDist` continues, :1323): the LAST candidate wins
            const unsigned long long ord = R.mode == MODE_TRI ? (unsigned long long)(0xffffff - pos) : (unsigned long long)pos;
            const unsigned long long key = ((unsigned long long)(rec >> 22) << 48) | (ord << 24) | (rec & 0x3fffffu);
            if (key < b1) { b2 = b1; b1 = key; } else if (key < b2) b2 = key;
        }
        const unsigned long long best = warp_min_u64(b1);
        if (best == NONE) continue;
        const unsigned long long second = warp_min_u64(b1 == best ? b2 : b1);
        const int bd = (int)(best >> 48), kp = (int)(best & KP_MASK), lvl = (int)((best >> KP_BITS) & 31);
        const int bd2 = second == NONE ? 256 : (int)(second >> 48), lvl2 = second == NONE ? -1 : (int)((second >> KP_BITS) & 31);
        bool accept;
        if (R.mode == MODE_MP) accept = bd <= ORBM_TH_HIGH && !(lvl == lvl2 && (float)bd > R.nnratio * (float)bd2);
        else if (R.mode == MODE_LAST) accept = bd <= R.th_accept;
        else if (R.mode == MODE_BOW) accept = bd <= ORBM_TH_LOW && (float)bd < R.nnratio * (float)bd2;
        else if (R.mode == MODE_BOW_KF) accept = bd < ORBM_TH_LOW && (float)bd < R.nnratio * (float)bd2;      // strict, src/ORBmatcher.cc:363
        else accept = true;                                                                                  // MODE_TRI: every gate is in the records
        if (!accept) continue;
        const bool by_query = R.mode == MODE_BOW_KF || R.mode == MODE_TRI;      // vpMatches12 / vMatches12 are indexed by the query key point
        const int oidx = by_query ? R.q_tag[q] : tbase + kp;
        if (lane == 0) {
            R.out[oidx] = by_query ? kp : R.q_tag[q];
            const bool claim = R.mode >= MODE_BOW ? true : (R.q_obs[q] != 0);
            if (claim) s_claim[kp >> 5] |= 1u << (kp & 31); else s_claim[kp >> 5] &= ~(1u << (kp & 31));
            if (R.check_ori && R.mode != MODE_MP) {
                float rot = __fsub_rn(R.q_angle[q], R.kp_angle[tbase + kp]);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
                if (bin == HISTO_LENGTH) bin = 0;
                m_kp[nlist] = oidx; m_bin[nlist] = bin;
                s_hist[bin]++;
            }
        }
        nmatch++; nlist++;
        __syncwarp();
    }
    if (R.check_ori && R.mode != MODE_MP) {
        __syncwarp();
        int ind1 = -1, ind2 = -1, ind3 = -1;
        {   // ComputeThreeMaxima (every lane computes the same)
            int max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < HISTO_LENGTH; i++) {
                const int s = s_hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
        }
        int removed = 0;
        for (int i = lane; i < nlist; i += 32) {
            const int bin = m_bin[i];
            // MODE_LAST: -2 = "was matched, then removed by the rotation check": the host writes -1 (mvpMapPoints[g] = NULL, :1082-1101) over
            // whatever the caller's in/out array held; -1 = never touched
            if (bin != ind1 && bin != ind2 && bin != ind3) { R.out[m_kp[i]] = R.mode == MODE_LAST ? -2 : -1; removed++; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatch -= removed;
    }
    if (lane == 0) R.seq_matches[seq] = nmatch;
}

// ------------------------------------------------------------------------------------------------ isInFrustum
struct FrustumDev {
    int n_cams, n_levels, for_all;
    float R[MAX_CAMS][9], t[MAX_CAMS][3], Ow[MAX_CAMS][3], K[MAX_CAMS][4], b[MAX_CAMS][4];
    float log_scale, cos_limit;
};
__global__ void __launch_bounds__(128) k_frustum(FrustumDev Q, const float* __restrict__ pos, const float* __restrict__ normal,
                                                 const float* __restrict__ max_dist, const float* __restrict__ min_dist, int n,
                                                 int* __restrict__ out, float* __restrict__ uvc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int o0 = 0, o1 = -1, o2 = 0;
    float u_ = 0, v_ = 0, c_ = 0;
    const float P0 = pos[3 * i], P1 = pos[3 * i + 1], P2 = pos[3 * i + 2];
    for (int ic = 0; ic < Q.n_cams; ic++) {
        if (ic != 0 && !Q.for_all) continue;
        const float* R = Q.R[ic];
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], P0), __fmul_rn(R[1], P1)), __fmul_rn(R[2], P2)), Q.t[ic][0]);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], P0), __fmul_rn(R[4], P1)), __fmul_rn(R[5], P2)), Q.t[ic][1]);
        const float Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], P0), __fmul_rn(R[7], P1)), __fmul_rn(R[8], P2)), Q.t[ic][2]);
        if (Z < 0.0f) continue;
        const float invz = __fdiv_rn(1.0f, Z);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(Q.K[ic][0], X), invz), Q.K[ic][2]);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(Q.K[ic][1], Y), invz), Q.K[ic][3]);
        if (u < Q.b[ic][0] || u > Q.b[ic][1]) continue;
        if (v < Q.b[ic][2] || v > Q.b[ic][3]) continue;
        const float maxDistance = __fmul_rn(1.2f, max_dist[i]), minDistance = __fmul_rn(0.8f, min_dist[i]);
        const float PO0 = __fsub_rn(P0, Q.Ow[ic][0]), PO1 = __fsub_rn(P1, Q.Ow[ic][1]), PO2 = __fsub_rn(P2, Q.Ow[ic][2]);
        const float dist = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn((double)PO0, (double)PO0), __dmul_rn((double)PO1, (double)PO1)), __dmul_rn((double)PO2, (double)PO2)));
        if (dist < minDistance || dist > maxDistance) continue;
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)PO0, (double)normal[3 * i]), __dmul_rn((double)PO1, (double)normal[3 * i + 1])),
                                     __dmul_rn((double)PO2, (double)normal[3 * i + 2]));
        const float viewCos = (float)(dot / (double)dist);
        if (viewCos < Q.cos_limit) continue;
        const float ratio = __fdiv_rn(max_dist[i], dist);
        int nScale = (int)ceil(log((double)ratio) / (double)Q.log_scale);
        if (nScale < 0) nScale = 0;
        else if (nScale >= Q.n_levels) nScale = Q.n_levels - 1;
        o0 = 1; o1 = ic; o2 = nScale; u_ = u; v_ = v; c_ = viewCos;
        break;
    }
    out[3 * i] = o0; out[3 * i + 1] = o1; out[3 * i + 2] = o2;
    uvc[3 * i] = u_; uvc[3 * i + 1] = v_; uvc[3 * i + 2] = c_;
}

// ------------------------------------------------------------------------------------------------ undistort
// cv::undistortPoints(mat, mat, K, distCoef, Mat(), K) as Frame::UndistortKeyPoints / ComputeImageBounds call it (src/Frame.cc:410-490):
// cvUndistortPointsInternal restated -- FP64, no contraction (this file is built with -fmad=false), five fixed-point iterations.
struct UndistDev { double fx, fy, cx, cy, k[12]; };
__global__ void __launch_bounds__(128) k_undistort(UndistDev U, const float* __restrict__ in, int stride_floats, int n, float* __restrict__ out, int out_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double u = in[(size_t)i * stride_floats], v = in[(size_t)i * stride_floats + 1];
    const double ifx = 1. / U.fx, ify = 1. / U.fy;
    double x = (u - U.cx) * ifx, y = (v - U.cy) * ify;
    const double x0 = x, y0 = y;
    const double* k = U.k;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = (u - U.cx) * ifx; y = (v - U.cy) * ify; break; }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = U.fx * x + 0 * y + U.cx, yy = 0 * x + U.fy * y + U.cy, ww = 1. / (0 * x + 0 * y + 1.0);
    out[(size_t)i * out_stride] = (float)(xx * ww);
    out[(size_t)i * out_stride + 1] = (float)(yy * ww);
}

// ================================================================================================ host side
// The orbm handle (orb_match.cu) owns the stream; the scratch arena of the searches lives here, keyed by handle.
struct orbm;
cudaStream_t orbm_stream_of(orbm*);
int orbm_device_of(orbm*);
void orbm_count_launches(orbm*, int n);

namespace {

struct Arena {                 // grow-only device scratch + pinned staging, one per calling thread
    uint8_t* d = nullptr; size_t dcap = 0;
    uint8_t* h = nullptr; size_t hcap = 0;
    uint8_t* recs = nullptr; size_t rcap = 0;      // candidate records (sized after the count pass)
    int* h_total = nullptr; int* d_total = nullptr;
    int device = -1;
    void release() {           // frees on the arena's own device; errors (e.g. the runtime is already unloading at process exit) are ignored
        if (device >= 0 && (d || h || recs || h_total)) {
            int prev = -1;
            cudaGetDevice(&prev);
            if (cudaSetDevice(device) == cudaSuccess) {
                if (d) cudaFree(d);
                if (h) cudaFreeHost(h);
                if (recs) cudaFree(recs);
                if (h_total) cudaFreeHost(h_total);
            }
            if (prev >= 0) cudaSetDevice(prev);
            cudaGetLastError();
        }
        d = h = recs = nullptr; dcap = hcap = rcap = 0; h_total = d_total = nullptr; device = -1;
    }
    ~Arena() { release(); }    // thread exit: the calling thread's scratch goes with it
};
thread_local Arena g_arena;

int arena_reserve(Arena& A, int device, size_t dbytes, size_t hbytes) {
    if (A.device != device) {
        A.release();
        A.device = device;
    }
    if (!A.h_total) {
        ORB_CUDA(cudaHostAlloc((void**)&A.h_total, 64, cudaHostAllocMapped));
        ORB_CUDA(cudaHostGetDevicePointer((void**)&A.d_total, A.h_total, 0));
    }
    if (dbytes > A.dcap) {
        if (A.d) cudaFree(A.d);
        A.d = nullptr; A.dcap = 0;
        const size_t want = dbytes + dbytes / 2 + (1 << 20);
        ORB_CUDA(cudaMalloc((void**)&A.d, want));
        A.dcap = want;
    }
    if (hbytes > A.hcap) {
        if (A.h) cudaFreeHost(A.h);
        A.h = nullptr; A.hcap = 0;
        const size_t want = hbytes + hbytes / 2 + (1 << 20);
        ORB_CUDA(cudaHostAlloc((void**)&A.h, want, cudaHostAllocDefault));
        A.hcap = want;
    }
    return ORB_OK;
}

struct Bump {
    size_t cur = 0;
    size_t add(size_t n) { const size_t o = (cur + 255) & ~(size_t)255; cur = o + n; return o; }
};

int check_frame(const orbm_frame_t* f, const char* who) {
    if (!f || !f->n_kp || !f->bounds || !f->scale_factors) ORB_FAIL(ORB_E_INVALID, "%s: NULL frame field", who);
    if (f->n_cams < 1 || f->n_cams > MAX_CAMS || f->n_levels < 1 || f->n_levels > 16) ORB_FAIL(ORB_E_INVALID, "%s: n_cams / n_levels out of range", who);
    long long tot = 0;
    for (int c = 0; c < f->n_cams; c++) {
        if (f->n_kp[c] < 0 || f->n_kp[c] > 32767) ORB_FAIL(ORB_E_INVALID, "%s: camera %d has %d keypoints (max 32767)", who, c, f->n_kp[c]);
        tot += f->n_kp[c];
    }
    if (tot > (long long)KP_MASK) ORB_FAIL(ORB_E_INVALID, "%s: too many keypoints", who);
    if (tot && (!f->kps_un || !f->desc)) ORB_FAIL(ORB_E_INVALID, "%s: NULL keypoints / descriptors", who);
    return ORB_OK;
}

// frame -> staging: float4 keypoints and descriptors into hk / hd; fills the FrameDev header (device pointers are set by the caller)
void stage_frame(const orbm_frame_t* f, FrameDev& F, uint8_t* hk_bytes, uint8_t* hd) {
    memset(&F, 0, sizeof(F));
    F.n_cams = f->n_cams; F.n_levels = f->n_levels;
    for (int c = 0; c < f->n_cams; c++) {
        F.first[c + 1] = F.first[c] + f->n_kp[c];
        const float* b = f->bounds + 4 * c;
        F.minX[c] = b[0]; F.maxX[c] = b[1]; F.minY[c] = b[2]; F.maxY[c] = b[3];
        F.invW[c] = (float)GRID_COLS / (float)(b[1] - b[0]);      // mvfGridElementWidthInv  src/Frame.cc:156-159
        F.invH[c] = (float)GRID_ROWS / (float)(b[3] - b[2]);
    }
    for (int c = f->n_cams; c < MAX_CAMS; c++) F.first[c + 1] = F.first[f->n_cams];
    F.totalN = F.first[f->n_cams];
    for (int l = 0; l < f->n_levels; l++) F.scale[l] = f->scale_factors[l];
    float4* hk = (float4*)hk_bytes;
    for (int g = 0; g < F.totalN; g++) {
        const orb_keypoint_t& k = f->kps_un[g];
        float w;
        const int oct = k.octave;
        memcpy(&w, &oct, 4);
        hk[g] = make_float4(k.x, k.y, k.angle, w);
    }
    if (F.totalN) memcpy(hd, f->desc, 32 * (size_t)F.totalN);
}

int recs_reserve(Arena& A, size_t bytes) {
    if (bytes <= A.rcap) return ORB_OK;
    if (A.recs) cudaFree(A.recs);
    A.recs = nullptr; A.rcap = 0;
    const size_t want = bytes + bytes / 2 + (1 << 20);
    ORB_CUDA(cudaMalloc((void**)&A.recs, want));
    A.rcap = want;
    return ORB_OK;
}

}  // namespace

static int run_window_search(orbm_t* m, const orbm_frame_t* frame, std::vector<Query>& queries, const std::vector<int>& q_seq, int n_seq, int mode,
                             float nnratio, int check_ori, const uint8_t* blocked, std::vector<int>& out, std::vector<int>& seq_matches) {
    const int device = orbm_device_of(m);
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = orbm_stream_of(m);
    const int nq = (int)queries.size();
    int totalN = 0;
    for (int c = 0; c < frame->n_cams; c++) totalN += frame->n_kp[c];
    // ---- layout: staged prefix (host mirror == device prefix), then device-only buffers
    Bump B;
    const size_t o_kp = B.add(16 * (size_t)totalN), o_desc = B.add(32 * (size_t)totalN);
    const size_t o_q = B.add(sizeof(Query) * (size_t)nq), o_seq = B.add(4 * (size_t)nq), o_tag = B.add(4 * (size_t)nq), o_obs = B.add((size_t)nq),
                 o_qang = B.add(4 * (size_t)nq), o_kang = B.add(4 * (size_t)totalN), o_blk = B.add((size_t)totalN);
    const size_t staged = B.add(0);
    const size_t o_celloff = B.add(4 * (size_t)frame->n_cams * (GRID_CELLS + 1)), o_cellidx = B.add(4 * (size_t)std::max(totalN, 1));
    const size_t o_cnt = B.add(4 * (size_t)(nq + 1)), o_off = B.add(4 * (size_t)(nq + 1)), o_out = B.add(4 * (size_t)std::max(totalN, 1));
    const size_t o_sm = B.add(4 * (size_t)n_seq), o_mkp = B.add(4 * (size_t)std::max(nq, 1) * n_seq), o_mbin = B.add(4 * (size_t)std::max(nq, 1) * n_seq);
    const size_t total_bytes = B.add(0);
    Arena& A = g_arena;
    int rc = arena_reserve(A, device, total_bytes, staged);
    if (rc != ORB_OK) return rc;
    FrameDev F;
    stage_frame(frame, F, A.h + o_kp, A.h + o_desc);
    F.kp = (const float4*)(A.d + o_kp); F.desc = (const uint4*)(A.d + o_desc);
    F.cell_off = (int*)(A.d + o_celloff); F.cell_idx = (int*)(A.d + o_cellidx);
    if (nq) memcpy(A.h + o_q, queries.data(), sizeof(Query) * (size_t)nq);
    if (nq) memcpy(A.h + o_seq, q_seq.data(), 4 * (size_t)nq);
    for (int q = 0; q < nq; q++) {
        ((int*)(A.h + o_tag))[q] = queries[q].tag;
        (A.h + o_obs)[q] = (uint8_t)queries[q].obs_positive;
        ((float*)(A.h + o_qang))[q] = queries[q].angle;
    }
    for (int g = 0; g < totalN; g++) ((float*)(A.h + o_kang))[g] = frame->kps_un[g].angle;
    if (totalN) memcpy(A.h + o_blk, blocked, (size_t)totalN);
    ORB_CUDA(cudaMemcpyAsync(A.d, A.h, staged, cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemsetAsync(A.d + o_out, 0xff, 4 * (size_t)std::max(totalN, 1), st));
    ORB_CUDA(cudaMemsetAsync(A.d + o_sm, 0, 4 * (size_t)n_seq, st));
    const size_t gsm = (GRID_CELLS + 1) * 4 + 2 * 32768;
    ORB_CUDA(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    k_grid_build<<<frame->n_cams, 256, gsm, st>>>(F);
    int launches = 1;
    if (nq > 0) {
        const Query* dq = (const Query*)(A.d + o_q);
        k_window_cands<false><<<(nq + 3) / 4, 128, 0, st>>>(F, dq, nq, (int*)(A.d + o_cnt), nullptr, nullptr);
        k_scan<<<1, 1024, 0, st>>>((const int*)(A.d + o_cnt), nq, (int*)(A.d + o_off), A.d_total);
        launches += 2;
        ORB_CUDA(cudaStreamSynchronize(st));
        const int total = *A.h_total;
        rc = recs_reserve(A, 4 * (size_t)std::max(total, 1));
        if (rc != ORB_OK) return rc;
        uint32_t* recs = (uint32_t*)A.recs;
        if (total > 0) { k_window_cands<true><<<(nq + 3) / 4, 128, 0, st>>>(F, dq, nq, nullptr, (const int*)(A.d + o_off), recs); launches++; }
        ResolveArgs R;
        memset(&R, 0, sizeof(R));
        R.mode = mode; R.nq = nq; R.n_bits = totalN;
        R.q_off = (const int*)(A.d + o_off); R.recs = recs; R.q_seq = (const int*)(A.d + o_seq); R.q_tag = (const int*)(A.d + o_tag);
        R.q_obs = A.d + o_obs; R.q_angle = (const float*)(A.d + o_qang); R.kp_angle = (const float*)(A.d + o_kang);
        R.seq_base = nullptr; R.blocked_in = A.d + o_blk; R.out = (int*)(A.d + o_out); R.seq_matches = (int*)(A.d + o_sm);
        R.m_kp = (int*)(A.d + o_mkp); R.m_bin = (int*)(A.d + o_mbin); R.nnratio = nnratio; R.check_ori = check_ori; R.th_accept = ORBM_TH_HIGH;
        k_resolve<<<n_seq, 32, ((totalN + 31) / 32 + 1) * 4, st>>>(R);
        launches++;
    }
    ORB_CUDA(cudaGetLastError());
    out.assign(totalN, -1);
    seq_matches.assign(n_seq, 0);
    if (totalN) ORB_CUDA(cudaMemcpyAsync(out.data(), A.d + o_out, 4 * (size_t)totalN, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(seq_matches.data(), A.d + o_sm, 4 * (size_t)n_seq, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    orbm_count_launches(m, launches);
    return ORB_OK;
}

extern "C" {

int orbm_search_by_projection(orbm_t* m, const orbm_frame_t* frame, const orbm_mp_t* mps, int n, float th, float nnratio, const uint8_t* blocked,
                              int32_t* kp_to_mp, int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection: NULL handle");
    int rc = check_frame(frame, "orbm_search_by_projection");
    if (rc != ORB_OK) return rc;
    if (n < 0 || (n && !mps) || !blocked || !kp_to_mp) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection: bad argument");
    std::vector<Query> qs((size_t)n);
    std::vector<int> seq((size_t)n, 0);
    const bool bFactor = th != 1.0;
    for (int i = 0; i < n; i++) {
        const orbm_mp_t& p = mps[i];
        Query& q = qs[i];
        memset(&q, 0, sizeof(q));
        q.valid = p.valid != 0;
        if (!q.valid) continue;
        if (p.cam < 0 || p.cam >= frame->n_cams || p.level < 0 || p.level >= frame->n_levels) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection: map point %d has camera / level out of range", i);
        float r = p.view_cos > 0.998 ? 2.5 : 4.0;          // RadiusByViewingCos  src/ORBmatcher.cc:65-71
        if (bFactor) r *= th;
        q.cam = p.cam; q.u = p.u; q.v = p.v; q.r = r * frame->scale_factors[p.level];
        q.minLevel = p.level - 1; q.maxLevel = p.level + 1;
        q.tag = i; q.obs_positive = p.obs_positive != 0;
        memcpy(q.desc, p.desc, 32);
    }
    std::vector<int> out, sm;
    rc = run_window_search(m, frame, qs, seq, 1, MODE_MP, nnratio, 0, blocked, out, sm);
    if (rc != ORB_OK) return rc;
    for (size_t g = 0; g < out.size(); g++) if (out[g] >= 0) kp_to_mp[g] = out[g];
    if (nmatches) *nmatches = sm[0];
    return ORB_OK;
}

int orbm_search_by_projection_last(orbm_t* m, const orbm_frame_t* cur, const float* Rsw, const float* tsw, const float* K, const orbm_lastframe_t* last,
                                   float th, int check_orientation, int map_scaled, const uint8_t* blocked, int32_t* kp_to_last, int32_t* per_cam,
                                   int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_last: NULL handle");
    int rc = check_frame(cur, "orbm_search_by_projection_last");
    if (rc != ORB_OK) return rc;
    if (!Rsw || !tsw || !K || !last || !blocked || !kp_to_last || last->n < 0) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_last: bad argument");
    if (last->n && (!last->cam || !last->valid || !last->pos || !last->desc || !last->octave || !last->angle || !last->obs_positive))
        ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_last: NULL last-frame field");
    const int n = last->n, C = cur->n_cams;
    std::vector<Query> qs((size_t)n);
    std::vector<int> seq((size_t)n, -1);
    for (int i = 0; i < n; i++) {
        Query& q = qs[i];
        memset(&q, 0, sizeof(q));
        const int ic = last->cam[i];
        if (ic < 0 || ic >= C || (ic != 0 && !map_scaled) || !last->valid[i]) continue;
        // projection with the last frame's map point (src/ORBmatcher.cc:996-1011), FP32 left to right
        const float* R = Rsw + 9 * ic;
        const float* t = tsw + 3 * ic;
        const float* X = last->pos + 3 * (size_t)i;
        volatile float xs = R[0] * X[0]; xs = xs + R[1] * X[1]; xs = xs + R[2] * X[2]; xs = xs + t[0];
        volatile float ys = R[3] * X[0]; ys = ys + R[4] * X[1]; ys = ys + R[5] * X[2]; ys = ys + t[1];
        volatile float zs = R[6] * X[0]; zs = zs + R[7] * X[1]; zs = zs + R[8] * X[2]; zs = zs + t[2];
        if (zs < 0) continue;
        const float invzs = (float)(1.0 / zs);
        volatile float u = K[4 * ic] * xs; u = u * invzs; u = u + K[4 * ic + 2];
        volatile float v = K[4 * ic + 1] * ys; v = v * invzs; v = v + K[4 * ic + 3];
        const float* b = cur->bounds + 4 * ic;
        if (u < b[0] || u > b[1] || v < b[2] || v > b[3]) continue;
        const int oct = last->octave[i];
        if (oct < 0 || oct >= cur->n_levels) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_last: octave out of range at %d", i);
        q.valid = 1; q.cam = ic; q.u = u; q.v = v; q.r = th * cur->scale_factors[oct];
        q.minLevel = oct - 1; q.maxLevel = oct + 1; q.tag = i; q.obs_positive = last->obs_positive[i] != 0; q.angle = last->angle[i];
        memcpy(q.desc, last->desc + 32 * (size_t)i, 32);
        seq[i] = ic;
    }
    std::vector<int> out, sm;
    rc = run_window_search(m, cur, qs, seq, C, MODE_LAST, 0.f, check_orientation != 0, blocked, out, sm);
    if (rc != ORB_OK) return rc;
    // camera loop of SearchByProjection(cur, last): cameras after one with <= 20 matches are not searched (src/ORBmatcher.cc:660-669)
    int total = 0, first = 0;
    for (int ic = 0; ic < C; ic++) {
        if (per_cam) per_cam[ic] = 0;
    }
    for (int ic = 0; ic < C; ic++) {
        const int lo = first, hi = first + cur->n_kp[ic];
        first = hi;
        if (ic != 0 && !map_scaled) continue;
        for (int g = lo; g < hi; g++) { if (out[g] >= 0) kp_to_last[g] = out[g]; else if (out[g] == -2) kp_to_last[g] = -1; }
        if (per_cam) per_cam[ic] = sm[ic];
        if (sm[ic] <= 20) { total = sm[ic]; break; }
        total += sm[ic];
    }
    if (nmatches) *nmatches = total;
    return ORB_OK;
}

// merge-join of two DBoW2 feature vectors + sequential claims (src/ORBmatcher.cc:184-269, 323-391, 1288-1383): Q side = the loop's outer
// features (queries), T side = inner features (targets).  pairs: (camera of Q, camera of T) per sequence; q_ok / t_skip are indexed globally.
// out: MODE_BOW -> [totT] global target -> global query; MODE_BOW_KF / MODE_TRI (one sequence) -> [n_kp of Q camera] local query -> local target.
struct BowJob {
    const orbm_bowside_t* Q; const orbm_bowside_t* T;
    int n_seq; int cq[MAX_CAMS], ct[MAX_CAMS];
    const uint8_t* q_ok; int q_ok_invert;        // query i is searched iff (q_ok[i] != 0) != q_ok_invert
    const uint8_t* t_skip;                       // may be NULL
    int mode; float nnratio; int check_ori;
    TriDev tri; const orb_keypoint_t* kps1; const orb_keypoint_t* kps2;
};
static int bow_join(orbm_t* m, const BowJob& J, const char* who, int32_t* out, int n_out, int* seq_matches) {
    const orbm_bowside_t *Q = J.Q, *T = J.T;
    if (!Q->n_kp || !T->n_kp || !Q->node_first || !T->node_first || Q->n_cams < 1 || Q->n_cams > MAX_CAMS || T->n_cams < 1 || T->n_cams > MAX_CAMS)
        ORB_FAIL(ORB_E_INVALID, "%s: bad feature-vector side", who);
    std::vector<int> firstQ(Q->n_cams + 1, 0), firstT(T->n_cams + 1, 0);
    for (int c = 0; c < Q->n_cams; c++) firstQ[c + 1] = firstQ[c] + Q->n_kp[c];
    for (int c = 0; c < T->n_cams; c++) firstT[c + 1] = firstT[c] + T->n_kp[c];
    const int totQ = firstQ[Q->n_cams], totT = firstT[T->n_cams];
    if (totT > (int)KP_MASK || totQ > (int)KP_MASK) ORB_FAIL(ORB_E_INVALID, "%s: too many keypoints", who);
    if ((totQ && (!Q->desc || !Q->angle)) || (totT && (!T->desc || !T->angle))) ORB_FAIL(ORB_E_INVALID, "%s: NULL descriptors / angles", who);
    for (int g = 0; g < n_out; g++) out[g] = -1;
    std::vector<BowQuery> qs;
    std::vector<int> q_off(1, 0), q_seq, q_tag;
    std::vector<float> q_ang;
    std::vector<int> seq_base(J.n_seq, 0);
    int maxT = 0;
    for (int sq = 0; sq < J.n_seq; sq++) {
        const int cq = J.cq[sq], ct = J.ct[sq];
        seq_base[sq] = firstT[ct];          // targets are camera-local in the records; angles / MODE_BOW outputs are global
        maxT = std::max(maxT, T->n_kp[ct]);
        int kq = Q->node_first[cq], kqEnd = Q->node_first[cq + 1], kt = T->node_first[ct], ktEnd = T->node_first[ct + 1];
        while (kq != kqEnd && kt != ktEnd) {
            if (Q->node_id[kq] == T->node_id[kt]) {
                for (int a = Q->node_off[kq]; a < Q->node_off[kq + 1]; a++) {
                    if (Q->idx[a] < 0 || Q->idx[a] >= Q->n_kp[cq]) ORB_FAIL(ORB_E_INVALID, "%s: feature index out of range", who);
                    const int gQ = firstQ[cq] + Q->idx[a];
                    if ((J.q_ok[gQ] != 0) == (J.q_ok_invert != 0)) continue;
                    BowQuery q;
                    q.gKF = gQ; q.f_begin = T->node_off[kt]; q.f_end = T->node_off[kt + 1]; q.cam = sq; q.firstF = firstT[ct]; q.angle = Q->angle[gQ];
                    qs.push_back(q);
                    q_off.push_back(q_off.back() + (q.f_end - q.f_begin));
                    q_seq.push_back(sq); q_tag.push_back(J.mode == MODE_BOW ? gQ : Q->idx[a]); q_ang.push_back(Q->angle[gQ]);
                }
                kq++; kt++;
            } else if (Q->node_id[kq] < T->node_id[kt]) {
                while (kq != kqEnd && Q->node_id[kq] < T->node_id[kt]) kq++;          // lower_bound
            } else {
                while (kt != ktEnd && T->node_id[kt] < Q->node_id[kq]) kt++;
            }
        }
    }
    const int nq = (int)qs.size();
    for (int sq = 0; sq < J.n_seq; sq++) seq_matches[sq] = 0;
    if (nq == 0) return ORB_OK;
    const int nIdxT = T->node_off[T->node_first[T->n_cams]];
    for (int e = 0; e < nIdxT; e++) if (T->idx[e] < 0) ORB_FAIL(ORB_E_INVALID, "%s: feature index out of range", who);
    const int device = orbm_device_of(m);
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = orbm_stream_of(m);
    const bool tri = J.tri.on != 0;
    Bump B;
    const size_t o_q = B.add(sizeof(BowQuery) * (size_t)nq), o_off = B.add(4 * (size_t)(nq + 1)), o_seq = B.add(4 * (size_t)nq), o_tag = B.add(4 * (size_t)nq),
                 o_qang = B.add(4 * (size_t)nq), o_dQ = B.add(32 * (size_t)totQ), o_dT = B.add(32 * (size_t)totT), o_idx = B.add(4 * (size_t)std::max(nIdxT, 1)),
                 o_tang = B.add(4 * (size_t)totT), o_base = B.add(4 * (size_t)J.n_seq), o_skip = B.add(J.t_skip ? (size_t)totT : 0),
                 o_k1 = B.add(tri ? 16 * (size_t)totQ : 0), o_k2 = B.add(tri ? 16 * (size_t)totT : 0);
    const size_t staged = B.add(0);
    const size_t o_recs = B.add(4 * (size_t)std::max(q_off.back(), 1)), o_out = B.add(4 * (size_t)std::max(n_out, 1)), o_sm = B.add(4 * (size_t)J.n_seq),
                 o_mkp = B.add(4 * (size_t)nq * J.n_seq), o_mbin = B.add(4 * (size_t)nq * J.n_seq);
    const size_t total_bytes = B.add(0);
    Arena& A = g_arena;
    int rc = arena_reserve(A, device, total_bytes, staged);
    if (rc != ORB_OK) return rc;
    memcpy(A.h + o_q, qs.data(), sizeof(BowQuery) * (size_t)nq);
    memcpy(A.h + o_off, q_off.data(), 4 * (size_t)(nq + 1));
    memcpy(A.h + o_seq, q_seq.data(), 4 * (size_t)nq);
    memcpy(A.h + o_tag, q_tag.data(), 4 * (size_t)nq);
    memcpy(A.h + o_qang, q_ang.data(), 4 * (size_t)nq);
    memcpy(A.h + o_dQ, Q->desc, 32 * (size_t)totQ);
    memcpy(A.h + o_dT, T->desc, 32 * (size_t)totT);
    if (nIdxT) memcpy(A.h + o_idx, T->idx, 4 * (size_t)nIdxT);
    memcpy(A.h + o_tang, T->angle, 4 * (size_t)totT);
    memcpy(A.h + o_base, seq_base.data(), 4 * (size_t)J.n_seq);
    if (J.t_skip) memcpy(A.h + o_skip, J.t_skip, (size_t)totT);
    TriDev TD = J.tri;
    if (tri) {
        float4* k1 = (float4*)(A.h + o_k1);
        float4* k2 = (float4*)(A.h + o_k2);
        for (int g = 0; g < totQ; g++) { float w; const int oct = J.kps1[g].octave; memcpy(&w, &oct, 4); k1[g] = make_float4(J.kps1[g].x, J.kps1[g].y, J.kps1[g].angle, w); }
        for (int g = 0; g < totT; g++) { float w; const int oct = J.kps2[g].octave; memcpy(&w, &oct, 4); k2[g] = make_float4(J.kps2[g].x, J.kps2[g].y, J.kps2[g].angle, w); }
        TD.kp1 = (const float4*)(A.d + o_k1); TD.kp2 = (const float4*)(A.d + o_k2);
    }
    ORB_CUDA(cudaMemcpyAsync(A.d, A.h, staged, cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemsetAsync(A.d + o_out, 0xff, 4 * (size_t)std::max(n_out, 1), st));
    k_bow_cands<<<(nq + 3) / 4, 128, 0, st>>>((const BowQuery*)(A.d + o_q), nq, (const int*)(A.d + o_off), (const uint4*)(A.d + o_dQ), (const uint4*)(A.d + o_dT),
                                              (const int*)(A.d + o_idx), (uint32_t*)(A.d + o_recs), J.t_skip ? A.d + o_skip : nullptr, TD);
    ResolveArgs R;
    memset(&R, 0, sizeof(R));
    R.mode = J.mode; R.nq = nq; R.n_bits = maxT;
    R.q_off = (const int*)(A.d + o_off); R.recs = (const uint32_t*)(A.d + o_recs); R.q_seq = (const int*)(A.d + o_seq); R.q_tag = (const int*)(A.d + o_tag);
    R.q_obs = nullptr; R.q_angle = (const float*)(A.d + o_qang); R.kp_angle = (const float*)(A.d + o_tang); R.seq_base = (const int*)(A.d + o_base);
    R.blocked_in = nullptr; R.out = (int*)(A.d + o_out); R.seq_matches = (int*)(A.d + o_sm); R.m_kp = (int*)(A.d + o_mkp); R.m_bin = (int*)(A.d + o_mbin);
    R.nnratio = J.nnratio; R.check_ori = J.check_ori != 0; R.th_accept = ORBM_TH_LOW;
    ORB_CUDA(cudaMemsetAsync(A.d + o_sm, 0, 4 * (size_t)J.n_seq, st));
    k_resolve<<<J.n_seq, 32, ((maxT + 31) / 32 + 1) * 4, st>>>(R);
    ORB_CUDA(cudaGetLastError());
    if (n_out) ORB_CUDA(cudaMemcpyAsync(out, A.d + o_out, 4 * (size_t)n_out, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(seq_matches, A.d + o_sm, 4 * (size_t)J.n_seq, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    orbm_count_launches(m, 2);
    return ORB_OK;
}

int orbm_search_by_bow(orbm_t* m, const orbm_bowside_t* F, const orbm_bowside_t* KF, const uint8_t* kf_mp_valid, float nnratio, int check_orientation,
                       int map_scaled, int32_t* f_to_kf, int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: NULL handle");
    if (!F || !KF || !kf_mp_valid || !f_to_kf || F->n_cams != KF->n_cams || F->n_cams < 1 || F->n_cams > MAX_CAMS) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: bad argument");
    if (!F->n_kp || !KF->n_kp) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: NULL field");
    BowJob J;
    memset(&J, 0, sizeof(J));
    J.Q = KF; J.T = F; J.q_ok = kf_mp_valid; J.mode = MODE_BOW; J.nnratio = nnratio; J.check_ori = check_orientation;
    int totF = 0;
    for (int c = 0; c < F->n_cams; c++) {
        totF += F->n_kp[c];
        if (c != 0 && !map_scaled) continue;
        J.cq[J.n_seq] = c; J.ct[J.n_seq] = c; J.n_seq++;
    }
    int sm[MAX_CAMS];
    const int rc = bow_join(m, J, "orbm_search_by_bow", f_to_kf, totF, sm);
    if (rc != ORB_OK) return rc;
    int tot = 0;
    for (int sq = 0; sq < J.n_seq; sq++) tot += sm[sq];
    if (nmatches) *nmatches = tot;
    return ORB_OK;
}

int orbm_search_by_bow_kf(orbm_t* m, const orbm_bowside_t* K1, int c1, const orbm_bowside_t* K2, int c2, const uint8_t* mp_valid1, const uint8_t* mp_valid2,
                          float nnratio, int check_orientation, int32_t* matches12, int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow_kf: NULL handle");
    if (!K1 || !K2 || !mp_valid1 || !mp_valid2 || !matches12 || !K1->n_kp || !K2->n_kp || c1 < 0 || c1 >= K1->n_cams || c2 < 0 || c2 >= K2->n_cams)
        ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow_kf: bad argument");
    int tot2 = 0, first2 = 0;
    for (int c = 0; c < K2->n_cams && c < MAX_CAMS; c++) { if (c < c2) first2 += K2->n_kp[c]; tot2 += K2->n_kp[c]; }
    std::vector<uint8_t> skip((size_t)std::max(tot2, 1));
    for (int g = 0; g < tot2; g++) skip[g] = mp_valid2[g] ? 0 : 1;
    BowJob J;
    memset(&J, 0, sizeof(J));
    J.Q = K1; J.T = K2; J.q_ok = mp_valid1; J.t_skip = skip.data(); J.mode = MODE_BOW_KF; J.nnratio = nnratio; J.check_ori = check_orientation;
    J.n_seq = 1; J.cq[0] = c1; J.ct[0] = c2;
    int sm[MAX_CAMS];
    const int n1 = K1->n_kp[c1];
    const int rc = bow_join(m, J, "orbm_search_by_bow_kf", matches12, n1, sm);
    if (rc != ORB_OK) return rc;
    for (int i = 0; i < n1; i++) if (matches12[i] >= 0) matches12[i] += first2;       // GetGlobalIdxByLocal(bestIdx2local, c2)  :367
    if (nmatches) *nmatches = sm[0];
    return ORB_OK;
}

int orbm_search_for_triangulation(orbm_t* m, const orbm_bowside_t* K1, const orbm_bowside_t* K2, int cam, const orb_keypoint_t* kps1,
                                  const orb_keypoint_t* kps2, const uint8_t* has_mp1, const uint8_t* has_mp2, const float* F12, const float* C1sw,
                                  const float* R2sw, const float* t2sw, const float* K2cam, const float* scale_factors, int n_levels,
                                  int check_orientation, int32_t* matches12, int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_for_triangulation: NULL handle");
    if (!K1 || !K2 || !kps1 || !kps2 || !has_mp1 || !has_mp2 || !F12 || !C1sw || !R2sw || !t2sw || !K2cam || !scale_factors || !matches12 || !K1->n_kp ||
        !K2->n_kp || cam < 0 || cam >= K1->n_cams || cam >= K2->n_cams || n_levels < 1 || n_levels > 16)
        ORB_FAIL(ORB_E_INVALID, "orbm_search_for_triangulation: bad argument");
    int tot2 = 0;
    for (int c = 0; c < K2->n_cams && c < MAX_CAMS; c++) tot2 += K2->n_kp[c];
    for (int g = 0; g < tot2; g++) if (kps2[g].octave < 0 || kps2[g].octave >= n_levels) ORB_FAIL(ORB_E_INVALID, "orbm_search_for_triangulation: octave out of range");
    BowJob J;
    memset(&J, 0, sizeof(J));
    J.Q = K1; J.T = K2; J.q_ok = has_mp1; J.q_ok_invert = 1; J.t_skip = has_mp2; J.mode = MODE_TRI; J.check_ori = check_orientation;
    J.n_seq = 1; J.cq[0] = cam; J.ct[0] = cam; J.kps1 = kps1; J.kps2 = kps2;
    // epipole of camera 1 in image 2 (:1261-1268), FP32 left to right
    volatile float C0 = R2sw[0] * C1sw[0]; C0 = C0 + R2sw[1] * C1sw[1]; C0 = C0 + R2sw[2] * C1sw[2]; C0 = C0 + t2sw[0];
    volatile float C1 = R2sw[3] * C1sw[0]; C1 = C1 + R2sw[4] * C1sw[1]; C1 = C1 + R2sw[5] * C1sw[2]; C1 = C1 + t2sw[1];
    volatile float C2 = R2sw[6] * C1sw[0]; C2 = C2 + R2sw[7] * C1sw[1]; C2 = C2 + R2sw[8] * C1sw[2]; C2 = C2 + t2sw[2];
    const float invz = 1.0f / C2;
    volatile float ex = K2cam[0] * C0; ex = ex * invz; ex = ex + K2cam[2];
    volatile float ey = K2cam[1] * C1; ey = ey * invz; ey = ey + K2cam[3];
    J.tri.on = 1; J.tri.ex = ex; J.tri.ey = ey;
    memcpy(J.tri.F12, F12, 36);
    for (int l = 0; l < n_levels; l++) J.tri.scale[l] = scale_factors[l];
    int sm[MAX_CAMS];
    const int rc = bow_join(m, J, "orbm_search_for_triangulation", matches12, K1->n_kp[cam], sm);
    if (rc != ORB_OK) return rc;
    if (nmatches) *nmatches = sm[0];
    return ORB_OK;
}

// ---- key-frame flavoured projection searches.  claims == true: project -> count -> scan -> fill -> sequential resolve (one camera);
// claims == false: project -> best key point per (camera, map point), no interaction between map points.
static int run_projected(orbm_t* m, const orbm_frame_t* frame, const orbm_frustum_t* view, const orbm_points_t* P, const char* who, int cam0, int n_cam_run,
                         float th, int flags, int kf_quirk, bool claims, int th_accept, int check_ori, const uint8_t* blocked /* global, claims only */,
                         int32_t* out /* claims: [totalN] */, int32_t* nmatches, int32_t* best_kp, int32_t* best_dist) {
    int rc = check_frame(frame, who);
    if (rc != ORB_OK) return rc;
    if (!view || !P || !view->Rsw || !view->tsw || !view->Ow || !view->K || view->n_cams != frame->n_cams) ORB_FAIL(ORB_E_INVALID, "%s: bad view", who);
    if (P->n < 0 || (P->n && (!P->valid || !P->pos || !P->max_dist || !P->min_dist || !P->desc))) ORB_FAIL(ORB_E_INVALID, "%s: NULL map-point field", who);
    if (P->n && (flags & PS_VIEW_ANGLE) && !P->normal) ORB_FAIL(ORB_E_INVALID, "%s: map-point normals are required", who);
    if (P->n && check_ori && !P->angle) ORB_FAIL(ORB_E_INVALID, "%s: map-point key point angles are required for the orientation check", who);
    if (cam0 < 0 || cam0 + n_cam_run > frame->n_cams) ORB_FAIL(ORB_E_INVALID, "%s: camera out of range", who);
    const int n = P->n, nq = n * n_cam_run;
    int totalN = 0;
    for (int c = 0; c < frame->n_cams; c++) totalN += frame->n_kp[c];
    if (nmatches) *nmatches = 0;
    if (!claims) for (int q = 0; q < nq; q++) { best_kp[q] = -1; best_dist[q] = 256; }
    if (n == 0 || totalN == 0) return ORB_OK;
    const int device = orbm_device_of(m);
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = orbm_stream_of(m);
    Bump B;
    const size_t o_kp = B.add(16 * (size_t)totalN), o_desc = B.add(32 * (size_t)totalN), o_kang = B.add(4 * (size_t)totalN), o_blk = B.add((size_t)totalN);
    const size_t o_pv = B.add((size_t)n), o_pp = B.add(12 * (size_t)n), o_pn = B.add(12 * (size_t)n), o_pmax = B.add(4 * (size_t)n), o_pmin = B.add(4 * (size_t)n),
                 o_pd = B.add(32 * (size_t)n), o_pa = B.add(4 * (size_t)n);
    const size_t staged = B.add(0);
    const size_t o_q = B.add(sizeof(Query) * (size_t)nq), o_seq = B.add(4 * (size_t)nq), o_tag = B.add(4 * (size_t)nq), o_obs = B.add((size_t)nq), o_qang = B.add(4 * (size_t)nq);
    const size_t o_celloff = B.add(4 * (size_t)frame->n_cams * (GRID_CELLS + 1)), o_cellidx = B.add(4 * (size_t)totalN);
    const size_t o_cnt = B.add(4 * (size_t)(nq + 1)), o_off = B.add(4 * (size_t)(nq + 1)), o_out = B.add(4 * (size_t)totalN);
    const size_t o_sm = B.add(4), o_mkp = B.add(4 * (size_t)nq), o_mbin = B.add(4 * (size_t)nq), o_bkp = B.add(4 * (size_t)nq), o_bd = B.add(4 * (size_t)nq);
    const size_t total_bytes = B.add(0);
    Arena& A = g_arena;
    rc = arena_reserve(A, device, total_bytes, staged);
    if (rc != ORB_OK) return rc;
    FrameDev F;
    stage_frame(frame, F, A.h + o_kp, A.h + o_desc);
    F.kp = (const float4*)(A.d + o_kp); F.desc = (const uint4*)(A.d + o_desc);
    F.cell_off = (int*)(A.d + o_celloff); F.cell_idx = (int*)(A.d + o_cellidx); F.kf_quirk = kf_quirk != 0;
    for (int g = 0; g < totalN; g++) ((float*)(A.h + o_kang))[g] = frame->kps_un[g].angle;
    if (blocked) memcpy(A.h + o_blk, blocked, (size_t)totalN); else memset(A.h + o_blk, 0, (size_t)totalN);
    memcpy(A.h + o_pv, P->valid, (size_t)n); memcpy(A.h + o_pp, P->pos, 12 * (size_t)n);
    if (P->normal) memcpy(A.h + o_pn, P->normal, 12 * (size_t)n);
    memcpy(A.h + o_pmax, P->max_dist, 4 * (size_t)n); memcpy(A.h + o_pmin, P->min_dist, 4 * (size_t)n); memcpy(A.h + o_pd, P->desc, 32 * (size_t)n);
    if (P->angle) memcpy(A.h + o_pa, P->angle, 4 * (size_t)n);
    ProjDev D;
    memset(&D, 0, sizeof(D));
    D.n = n; D.n_levels = frame->n_levels; D.flags = flags; D.cam0 = cam0; D.th = th; D.log_scale = view->log_scale_factor;
    for (int c = 0; c < frame->n_cams; c++) {
        memcpy(D.R[c], view->Rsw + 9 * c, 36); memcpy(D.t[c], view->tsw + 3 * c, 12); memcpy(D.Ow[c], view->Ow + 3 * c, 12);
        memcpy(D.K[c], view->K + 4 * c, 16); memcpy(D.b[c], frame->bounds + 4 * c, 16);
    }
    for (int l = 0; l < frame->n_levels; l++) D.scale[l] = frame->scale_factors[l];
    D.valid = A.d + o_pv; D.pos = (const float*)(A.d + o_pp); D.normal = P->normal ? (const float*)(A.d + o_pn) : nullptr;
    D.max_dist = (const float*)(A.d + o_pmax); D.min_dist = (const float*)(A.d + o_pmin); D.desc = (const uint4*)(A.d + o_pd);
    D.angle = P->angle ? (const float*)(A.d + o_pa) : nullptr;
    ORB_CUDA(cudaMemcpyAsync(A.d, A.h, staged, cudaMemcpyHostToDevice, st));
    const size_t gsm = (GRID_CELLS + 1) * 4 + 2 * 32768;
    ORB_CUDA(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    k_grid_build<<<frame->n_cams, 256, gsm, st>>>(F);
    Query* dq = (Query*)(A.d + o_q);
    k_project_queries<<<(nq + 127) / 128, 128, 0, st>>>(D, nq, dq, claims ? (int*)(A.d + o_seq) : nullptr, (int*)(A.d + o_tag), A.d + o_obs, (float*)(A.d + o_qang));
    int launches = 2;
    if (!claims) {
        k_window_best<<<(nq + 3) / 4, 128, 0, st>>>(F, dq, nq, (flags & PS_CHI2) != 0, (int*)(A.d + o_bkp), (int*)(A.d + o_bd));
        ORB_CUDA(cudaGetLastError());
        ORB_CUDA(cudaMemcpyAsync(best_kp, A.d + o_bkp, 4 * (size_t)nq, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(best_dist, A.d + o_bd, 4 * (size_t)nq, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        orbm_count_launches(m, launches + 1);
        return ORB_OK;
    }
    ORB_CUDA(cudaMemsetAsync(A.d + o_out, 0xff, 4 * (size_t)totalN, st));
    ORB_CUDA(cudaMemsetAsync(A.d + o_sm, 0, 4, st));
    k_window_cands<false><<<(nq + 3) / 4, 128, 0, st>>>(F, dq, nq, (int*)(A.d + o_cnt), nullptr, nullptr);
    k_scan<<<1, 1024, 0, st>>>((const int*)(A.d + o_cnt), nq, (int*)(A.d + o_off), A.d_total);
    launches += 2;
    ORB_CUDA(cudaStreamSynchronize(st));
    const int total = *A.h_total;
    rc = recs_reserve(A, 4 * (size_t)std::max(total, 1));
    if (rc != ORB_OK) return rc;
    uint32_t* recs = (uint32_t*)A.recs;
    if (total > 0) { k_window_cands<true><<<(nq + 3) / 4, 128, 0, st>>>(F, dq, nq, nullptr, (const int*)(A.d + o_off), recs); launches++; }
    ResolveArgs R;
    memset(&R, 0, sizeof(R));
    R.mode = MODE_LAST; R.nq = nq; R.n_bits = totalN;
    R.q_off = (const int*)(A.d + o_off); R.recs = recs; R.q_seq = (const int*)(A.d + o_seq); R.q_tag = (const int*)(A.d + o_tag);
    R.q_obs = A.d + o_obs; R.q_angle = (const float*)(A.d + o_qang); R.kp_angle = (const float*)(A.d + o_kang);
    R.seq_base = nullptr; R.blocked_in = A.d + o_blk; R.out = (int*)(A.d + o_out); R.seq_matches = (int*)(A.d + o_sm);
    R.m_kp = (int*)(A.d + o_mkp); R.m_bin = (int*)(A.d + o_mbin); R.nnratio = 0.f; R.check_ori = check_ori != 0; R.th_accept = th_accept;
    k_resolve<<<1, 32, ((totalN + 31) / 32 + 1) * 4, st>>>(R);
    launches++;
    ORB_CUDA(cudaGetLastError());
    std::vector<int> res((size_t)totalN);
    int nm = 0;
    ORB_CUDA(cudaMemcpyAsync(res.data(), A.d + o_out, 4 * (size_t)totalN, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(&nm, A.d + o_sm, 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    orbm_count_launches(m, launches);
    for (int g = 0; g < totalN; g++) { if (res[g] >= 0) out[g] = res[g]; else if (res[g] == -2) out[g] = -1; }
    if (nmatches) *nmatches = nm;
    return ORB_OK;
}

int orbm_search_by_projection_reloc(orbm_t* m, const orbm_frame_t* F, const orbm_frustum_t* view, int cam, const orbm_points_t* P, float th, int orb_dist,
                                    int check_orientation, const uint8_t* blocked, int32_t* kp_to_point, int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_reloc: NULL handle");
    if (!blocked || !kp_to_point) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_reloc: bad argument");
    return run_projected(m, F, view, P, "orbm_search_by_projection_reloc", cam, 1, th, PS_LEVEL_UP, 0, true, orb_dist, check_orientation, blocked, kp_to_point,
                         nmatches, nullptr, nullptr);
}

int orbm_search_by_projection_sim3(orbm_t* m, const orbm_frame_t* KF, const orbm_frustum_t* view, int cam, const orbm_points_t* P, int th, int kf_index_quirk,
                                   const uint8_t* matched_local, int32_t* local_to_point, int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_sim3: NULL handle");
    if (!KF || !KF->n_kp || !matched_local || !local_to_point || cam < 0 || cam >= KF->n_cams) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_projection_sim3: bad argument");
    int first = 0, totalN = 0;
    for (int c = 0; c < KF->n_cams && c < MAX_CAMS; c++) { if (c < cam) first += KF->n_kp[c]; totalN += KF->n_kp[c]; }
    // the reference indexes vpMatched with the camera-local key point index (src/ORBmatcher.cc:504, 523): expand to the global layout and back
    std::vector<uint8_t> blocked((size_t)std::max(totalN, 1), 0);
    std::vector<int32_t> out((size_t)std::max(totalN, 1), -1);
    for (int l = 0; l < KF->n_kp[cam]; l++) blocked[first + l] = matched_local[l];
    const int rc = run_projected(m, KF, view, P, "orbm_search_by_projection_sim3", cam, 1, (float)th, PS_DEPTH_POS | PS_NORMALISE_FIRST | PS_HALF_OPEN | PS_VIEW_ANGLE,
                                 kf_index_quirk, true, ORBM_TH_LOW, 0, blocked.data(), out.data(), nmatches, nullptr, nullptr);
    if (rc != ORB_OK) return rc;
    for (int l = 0; l < KF->n_kp[cam]; l++) if (out[first + l] >= 0) local_to_point[l] = out[first + l];
    return ORB_OK;
}

int orbm_project_best(orbm_t* m, const orbm_frame_t* KF, const orbm_frustum_t* view, const orbm_points_t* P, float th, int variant, int kf_index_quirk,
                      int32_t* best_kp, int32_t* best_dist) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_project_best: NULL handle");
    if (!KF || !best_kp || !best_dist) ORB_FAIL(ORB_E_INVALID, "orbm_project_best: bad argument");
    int flags;
    if (variant == ORBM_KF_SEARCH) flags = PS_LEVEL_UP;
    else if (variant == ORBM_KF_FUSE) flags = PS_DEPTH_POS | PS_NORMALISE_FIRST | PS_HALF_OPEN | PS_VIEW_ANGLE | PS_CHI2;
    else if (variant == ORBM_KF_FUSE_SIM3) flags = PS_DEPTH_POS | PS_NORMALISE_FIRST | PS_HALF_OPEN | PS_VIEW_ANGLE;
    else ORB_FAIL(ORB_E_INVALID, "orbm_project_best: unknown variant %d", variant);
    return run_projected(m, KF, view, P, "orbm_project_best", 0, KF->n_cams, th, flags, kf_index_quirk, false, 0, 0, nullptr, nullptr, nullptr, best_kp, best_dist);
}

// One camera of orbm_project_best: best_kp / best_dist [P->n].  This is the form that reproduces the reference's loops exactly when the
// map changes between cameras: SearchByProjection(pKF, vpMapPoints, sFound, th, ORBdist) refreshes a matched point's normal, depth range
// and descriptor (UpdateNormalAndDepth / ComputeDistinctiveDescriptors, src/ORBmatcher.cc:783-787) before the next camera projects it
// again, and Fuse skips in camera 1 what camera 0 added (IsInKeyFrame, :1452).  The adaptor runs  for (s in cameras) { flatten the
// CURRENT state of the points; orbm_project_best_cam(s); apply the reference's own `if (bestDist <= ...)` block in list order }.
// Inside one camera a search reads only the point's own fields and the key frame's key points, which no earlier iteration of that
// camera's loop changes for a point still to be processed (a point replaced earlier is bad by then and is skipped when applying).
int orbm_project_best_cam(orbm_t* m, const orbm_frame_t* KF, const orbm_frustum_t* view, const orbm_points_t* P, float th, int variant, int kf_index_quirk, int cam,
                          int32_t* best_kp, int32_t* best_dist) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_project_best_cam: NULL handle");
    if (!KF || !best_kp || !best_dist) ORB_FAIL(ORB_E_INVALID, "orbm_project_best_cam: bad argument");
    int flags;
    if (variant == ORBM_KF_SEARCH) flags = PS_LEVEL_UP;
    else if (variant == ORBM_KF_FUSE) flags = PS_DEPTH_POS | PS_NORMALISE_FIRST | PS_HALF_OPEN | PS_VIEW_ANGLE | PS_CHI2;
    else if (variant == ORBM_KF_FUSE_SIM3) flags = PS_DEPTH_POS | PS_NORMALISE_FIRST | PS_HALF_OPEN | PS_VIEW_ANGLE;
    else ORB_FAIL(ORB_E_INVALID, "orbm_project_best_cam: unknown variant %d", variant);
    return run_projected(m, KF, view, P, "orbm_project_best_cam", cam, 1, th, flags, kf_index_quirk, false, 0, 0, nullptr, nullptr, nullptr, best_kp, best_dist);
}

int orbm_is_in_frustum(orbm_t* m, const orbm_frustum_t* fr, const float* pos, const float* normal, const float* max_dist, const float* min_dist, int n,
                       float viewing_cos_limit, int for_all_cams, int32_t* out, float* uvc) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_is_in_frustum: NULL handle");
    if (n == 0) return ORB_OK;
    if (!fr || !pos || !normal || !max_dist || !min_dist || !out || !uvc || n < 0) ORB_FAIL(ORB_E_INVALID, "orbm_is_in_frustum: bad argument");
    if (fr->n_cams < 1 || fr->n_cams > MAX_CAMS || !fr->Rsw || !fr->tsw || !fr->Ow || !fr->K || !fr->bounds) ORB_FAIL(ORB_E_INVALID, "orbm_is_in_frustum: bad frame");
    FrustumDev Q;
    memset(&Q, 0, sizeof(Q));
    Q.n_cams = fr->n_cams; Q.n_levels = fr->n_levels; Q.for_all = for_all_cams != 0; Q.log_scale = fr->log_scale_factor; Q.cos_limit = viewing_cos_limit;
    for (int c = 0; c < fr->n_cams; c++) {
        memcpy(Q.R[c], fr->Rsw + 9 * c, 36); memcpy(Q.t[c], fr->tsw + 3 * c, 12); memcpy(Q.Ow[c], fr->Ow + 3 * c, 12);
        memcpy(Q.K[c], fr->K + 4 * c, 16); memcpy(Q.b[c], fr->bounds + 4 * c, 16);
    }
    const int device = orbm_device_of(m);
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = orbm_stream_of(m);
    Bump B;
    const size_t o_pos = B.add(12 * (size_t)n), o_nrm = B.add(12 * (size_t)n), o_max = B.add(4 * (size_t)n), o_min = B.add(4 * (size_t)n);
    const size_t staged = B.add(0);
    const size_t o_out = B.add(12 * (size_t)n), o_uvc = B.add(12 * (size_t)n);
    Arena& A = g_arena;
    int rc = arena_reserve(A, device, B.add(0), staged);
    if (rc != ORB_OK) return rc;
    memcpy(A.h + o_pos, pos, 12 * (size_t)n); memcpy(A.h + o_nrm, normal, 12 * (size_t)n);
    memcpy(A.h + o_max, max_dist, 4 * (size_t)n); memcpy(A.h + o_min, min_dist, 4 * (size_t)n);
    ORB_CUDA(cudaMemcpyAsync(A.d, A.h, staged, cudaMemcpyHostToDevice, st));
    k_frustum<<<(n + 127) / 128, 128, 0, st>>>(Q, (const float*)(A.d + o_pos), (const float*)(A.d + o_nrm), (const float*)(A.d + o_max), (const float*)(A.d + o_min), n,
                                               (int*)(A.d + o_out), (float*)(A.d + o_uvc));
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(out, A.d + o_out, 12 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(uvc, A.d + o_uvc, 12 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    orbm_count_launches(m, 1);
    return ORB_OK;
}


static int undistort_run(orbm_t* m, const float* pts, int stride_floats, int n, const float* K4, const float* dist, int n_dist, float* out, int out_stride) {
    UndistDev U;
    U.fx = K4[0]; U.fy = K4[1]; U.cx = K4[2]; U.cy = K4[3];
    for (int i = 0; i < 12; i++) U.k[i] = i < n_dist ? (double)dist[i] : 0.0;
    const int device = orbm_device_of(m);
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = orbm_stream_of(m);
    const size_t in_b = 4 * (size_t)n * stride_floats, out_b = 4 * (size_t)n * out_stride;
    Arena& A = g_arena;
    int rc = arena_reserve(A, device, ((in_b + 255) & ~(size_t)255) + out_b + 512, in_b + out_b + 512);
    if (rc != ORB_OK) return rc;
    uint8_t* d_out = A.d + ((in_b + 255) & ~(size_t)255);
    memcpy(A.h, pts, in_b);
    if (out != pts) memcpy(A.h + in_b, out, out_b);      // fields the kernel does not touch keep the caller's values
    else memcpy(A.h + in_b, pts, out_b);
    ORB_CUDA(cudaMemcpyAsync(A.d, A.h, in_b, cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(d_out, A.h + in_b, out_b, cudaMemcpyHostToDevice, st));
    k_undistort<<<(n + 127) / 128, 128, 0, st>>>(U, (const float*)A.d, stride_floats, n, (float*)d_out, out_stride);
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(out, d_out, out_b, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    orbm_count_launches(m, 1);
    return ORB_OK;
}

int orbm_undistort_keypoints(orbm_t* m, const orb_keypoint_t* kps, int n, const float* K4, const float* dist, int n_dist, orb_keypoint_t* kps_un) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_undistort_keypoints: NULL handle");
    if (n == 0) return ORB_OK;
    if (!kps || !kps_un || !K4 || n < 0 || n_dist < 0 || n_dist > 12 || (n_dist && !dist)) ORB_FAIL(ORB_E_INVALID, "orbm_undistort_keypoints: bad argument");
    if (kps_un != kps) memcpy(kps_un, kps, sizeof(orb_keypoint_t) * (size_t)n);
    if (n_dist == 0 || dist[0] == 0.0f) return ORB_OK;                       // src/Frame.cc:414-418: no distortion -> copy
    return undistort_run(m, &kps->x, 7, n, K4, dist, n_dist, &kps_un->x, 7);
}

int orbm_image_bounds(orbm_t* m, int width, int height, const float* K4, const float* dist, int n_dist, float* bounds) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_image_bounds: NULL handle");
    if (!K4 || !bounds || n_dist < 0 || n_dist > 12 || (n_dist && !dist)) ORB_FAIL(ORB_E_INVALID, "orbm_image_bounds: bad argument");
    if (n_dist > 0 && dist[0] != 0.0f) {
        const float c[8] = {0.f, 0.f, (float)width, 0.f, 0.f, (float)height, (float)width, (float)height};
        float o[8];
        int rc = undistort_run(m, c, 2, 4, K4, dist, n_dist, o, 2);
        if (rc != ORB_OK) return rc;
        bounds[0] = std::min(o[0], o[4]); bounds[1] = std::max(o[2], o[6]);
        bounds[2] = std::min(o[1], o[3]); bounds[3] = std::max(o[5], o[7]);
    } else {
        bounds[0] = 0.f; bounds[1] = (float)width; bounds[2] = 0.f; bounds[3] = (float)height;
    }
    return ORB_OK;
}

}  // extern "C"
