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
                        const int oct = __float_as_int(k.w);
                        pred = oct >= Q.minLevel && oct <= Q.maxLevel && fabsf(__fsub_rn(k.x, x)) < r && fabsf(__fsub_rn(k.y, y)) < r;
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

// ------------------------------------------------------------------------------------------------ BoW candidates
struct BowQuery { int gKF, f_begin, f_end, cam, firstF; float angle; };   // f_begin..f_end: range in the frame's idx array
__global__ void __launch_bounds__(128) k_bow_cands(const BowQuery* __restrict__ qs, int nq, const int* __restrict__ q_off, const uint4* __restrict__ descKF,
                                                   const uint4* __restrict__ descF, const int* __restrict__ idxF, uint32_t* __restrict__ recs) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const BowQuery Q = qs[q];
    const uint4 qa = descKF[2 * (size_t)Q.gKF], qb = descKF[2 * (size_t)Q.gKF + 1];
    const uint32_t d[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    const int base = q_off[q];
    for (int e = Q.f_begin + lane; e < Q.f_end; e += 32) {
        const int local = idxF[e];
        const size_t g = (size_t)Q.firstF + local;
        recs[base + (e - Q.f_begin)] = ((uint32_t)hamming256(d, descF[2 * g], descF[2 * g + 1]) << 22) | (uint32_t)local;
    }
}

// ------------------------------------------------------------------------------------------------ resolve
enum { MODE_MP = 0, MODE_LAST = 1, MODE_BOW = 2 };
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
};

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}

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
            if ((s_claim[kp >> 5] >> (kp & 31)) & 1u) continue;
            const unsigned long long key = ((unsigned long long)(rec >> 22) << 48) | ((unsigned long long)pos << 24) | (rec & 0x3fffffu);
            if (key < b1) { b2 = b1; b1 = key; } else if (key < b2) b2 = key;
        }
        const unsigned long long best = warp_min_u64(b1);
        if (best == NONE) continue;
        const unsigned long long second = warp_min_u64(b1 == best ? b2 : b1);
        const int bd = (int)(best >> 48), kp = (int)(best & KP_MASK), lvl = (int)((best >> KP_BITS) & 31);
        const int bd2 = second == NONE ? 256 : (int)(second >> 48), lvl2 = second == NONE ? -1 : (int)((second >> KP_BITS) & 31);
        bool accept;
        if (R.mode == MODE_MP) accept = bd <= ORBM_TH_HIGH && !(lvl == lvl2 && (float)bd > R.nnratio * (float)bd2);
        else if (R.mode == MODE_LAST) accept = bd <= ORBM_TH_HIGH;
        else accept = bd <= ORBM_TH_LOW && (float)bd < R.nnratio * (float)bd2;
        if (!accept) continue;
        if (lane == 0) {
            R.out[tbase + kp] = R.q_tag[q];
            const bool claim = R.mode == MODE_BOW ? true : (R.q_obs[q] != 0);
            if (claim) s_claim[kp >> 5] |= 1u << (kp & 31); else s_claim[kp >> 5] &= ~(1u << (kp & 31));
            if (R.check_ori && R.mode != MODE_MP) {
                float rot = __fsub_rn(R.q_angle[q], R.kp_angle[tbase + kp]);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
                if (bin == HISTO_LENGTH) bin = 0;
                m_kp[nlist] = kp; m_bin[nlist] = bin;
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
            if (bin != ind1 && bin != ind2 && bin != ind3) { R.out[tbase + m_kp[i]] = -1; removed++; }
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
};
thread_local Arena g_arena;

int arena_reserve(Arena& A, int device, size_t dbytes, size_t hbytes) {
    if (A.device != device) {
        if (A.d) cudaFree(A.d);
        if (A.h) cudaFreeHost(A.h);
        if (A.recs) cudaFree(A.recs);
        if (A.h_total) cudaFreeHost(A.h_total);
        A = Arena();
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
        R.m_kp = (int*)(A.d + o_mkp); R.m_bin = (int*)(A.d + o_mbin); R.nnratio = nnratio; R.check_ori = check_ori;
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
        for (int g = lo; g < hi; g++) if (out[g] >= 0) kp_to_last[g] = out[g];
        if (per_cam) per_cam[ic] = sm[ic];
        if (sm[ic] <= 20) { total = sm[ic]; break; }
        total += sm[ic];
    }
    if (nmatches) *nmatches = total;
    return ORB_OK;
}

int orbm_search_by_bow(orbm_t* m, const orbm_bowside_t* F, const orbm_bowside_t* KF, const uint8_t* kf_mp_valid, float nnratio, int check_orientation,
                       int map_scaled, int32_t* f_to_kf, int32_t* nmatches) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: NULL handle");
    if (!F || !KF || !kf_mp_valid || !f_to_kf || F->n_cams != KF->n_cams || F->n_cams < 1 || F->n_cams > MAX_CAMS) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: bad argument");
    if (!F->n_kp || !KF->n_kp || !F->node_first || !KF->node_first) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: NULL field");
    const int C = F->n_cams;
    std::vector<int> firstF(C + 1, 0), firstK(C + 1, 0);
    for (int c = 0; c < C; c++) { firstF[c + 1] = firstF[c] + F->n_kp[c]; firstK[c + 1] = firstK[c] + KF->n_kp[c]; }
    const int totF = firstF[C], totK = firstK[C];
    if (totF > (int)KP_MASK) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: too many keypoints");
    for (int g = 0; g < totF; g++) f_to_kf[g] = -1;
    // merge-join of the two feature vectors (src/ORBmatcher.cc:184-269): the queries in the reference's visiting order
    std::vector<BowQuery> qs;
    std::vector<int> q_off(1, 0), q_seq, q_tag;
    std::vector<float> q_ang;
    for (int ic = 0; ic < C; ic++) {
        if (ic != 0 && !map_scaled) continue;
        int kf = KF->node_first[ic], kfEnd = KF->node_first[ic + 1], ff = F->node_first[ic], ffEnd = F->node_first[ic + 1];
        while (kf != kfEnd && ff != ffEnd) {
            if (KF->node_id[kf] == F->node_id[ff]) {
                for (int a = KF->node_off[kf]; a < KF->node_off[kf + 1]; a++) {
                    const int gKF = firstK[ic] + KF->idx[a];
                    if (KF->idx[a] < 0 || KF->idx[a] >= KF->n_kp[ic]) ORB_FAIL(ORB_E_INVALID, "orbm_search_by_bow: key-frame feature index out of range");
                    if (!kf_mp_valid[gKF]) continue;
                    BowQuery q;
                    q.gKF = gKF; q.f_begin = F->node_off[ff]; q.f_end = F->node_off[ff + 1]; q.cam = ic; q.firstF = firstF[ic]; q.angle = KF->angle[gKF];
                    qs.push_back(q);
                    q_off.push_back(q_off.back() + (q.f_end - q.f_begin));
                    q_seq.push_back(ic); q_tag.push_back(gKF); q_ang.push_back(KF->angle[gKF]);
                }
                kf++; ff++;
            } else if (KF->node_id[kf] < F->node_id[ff]) {
                while (kf != kfEnd && KF->node_id[kf] < F->node_id[ff]) kf++;
            } else {
                while (ff != ffEnd && F->node_id[ff] < KF->node_id[kf]) ff++;
            }
        }
    }
    const int nq = (int)qs.size();
    if (nmatches) *nmatches = 0;
    if (nq == 0) return ORB_OK;
    const int nIdxF = F->node_off[F->node_first[C]];
    const int device = orbm_device_of(m);
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = orbm_stream_of(m);
    Bump B;
    const size_t o_q = B.add(sizeof(BowQuery) * (size_t)nq), o_off = B.add(4 * (size_t)(nq + 1)), o_seq = B.add(4 * (size_t)nq), o_tag = B.add(4 * (size_t)nq),
                 o_qang = B.add(4 * (size_t)nq), o_dK = B.add(32 * (size_t)totK), o_dF = B.add(32 * (size_t)totF), o_idx = B.add(4 * (size_t)std::max(nIdxF, 1)),
                 o_fang = B.add(4 * (size_t)totF), o_base = B.add(4 * (size_t)C);
    const size_t staged = B.add(0);
    const size_t o_recs = B.add(4 * (size_t)std::max(q_off.back(), 1)), o_out = B.add(4 * (size_t)totF), o_sm = B.add(4 * (size_t)C),
                 o_mkp = B.add(4 * (size_t)nq * C), o_mbin = B.add(4 * (size_t)nq * C);
    const size_t total_bytes = B.add(0);
    Arena& A = g_arena;
    int rc = arena_reserve(A, device, total_bytes, staged);
    if (rc != ORB_OK) return rc;
    memcpy(A.h + o_q, qs.data(), sizeof(BowQuery) * (size_t)nq);
    memcpy(A.h + o_off, q_off.data(), 4 * (size_t)(nq + 1));
    memcpy(A.h + o_seq, q_seq.data(), 4 * (size_t)nq);
    memcpy(A.h + o_tag, q_tag.data(), 4 * (size_t)nq);
    memcpy(A.h + o_qang, q_ang.data(), 4 * (size_t)nq);
    memcpy(A.h + o_dK, KF->desc, 32 * (size_t)totK);
    memcpy(A.h + o_dF, F->desc, 32 * (size_t)totF);
    if (nIdxF) memcpy(A.h + o_idx, F->idx, 4 * (size_t)nIdxF);
    memcpy(A.h + o_fang, F->angle, 4 * (size_t)totF);
    memcpy(A.h + o_base, firstF.data(), 4 * (size_t)C);
    ORB_CUDA(cudaMemcpyAsync(A.d, A.h, staged, cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemsetAsync(A.d + o_out, 0xff, 4 * (size_t)totF, st));
    k_bow_cands<<<(nq + 3) / 4, 128, 0, st>>>((const BowQuery*)(A.d + o_q), nq, (const int*)(A.d + o_off), (const uint4*)(A.d + o_dK), (const uint4*)(A.d + o_dF),
                                              (const int*)(A.d + o_idx), (uint32_t*)(A.d + o_recs));
    ResolveArgs R;
    memset(&R, 0, sizeof(R));
    int maxF = 0;
    for (int c = 0; c < C; c++) maxF = std::max(maxF, F->n_kp[c]);
    R.mode = MODE_BOW; R.nq = nq; R.n_bits = maxF;
    R.q_off = (const int*)(A.d + o_off); R.recs = (const uint32_t*)(A.d + o_recs); R.q_seq = (const int*)(A.d + o_seq); R.q_tag = (const int*)(A.d + o_tag);
    R.q_obs = nullptr; R.q_angle = (const float*)(A.d + o_qang); R.kp_angle = (const float*)(A.d + o_fang); R.seq_base = (const int*)(A.d + o_base);
    R.blocked_in = nullptr; R.out = (int*)(A.d + o_out); R.seq_matches = (int*)(A.d + o_sm); R.m_kp = (int*)(A.d + o_mkp); R.m_bin = (int*)(A.d + o_mbin);
    R.nnratio = nnratio; R.check_ori = check_orientation != 0;
    ORB_CUDA(cudaMemsetAsync(A.d + o_sm, 0, 4 * (size_t)C, st));
    k_resolve<<<C, 32, ((maxF + 31) / 32 + 1) * 4, st>>>(R);
    ORB_CUDA(cudaGetLastError());
    std::vector<int> sm(C, 0);
    ORB_CUDA(cudaMemcpyAsync(f_to_kf, A.d + o_out, 4 * (size_t)totF, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(sm.data(), A.d + o_sm, 4 * (size_t)C, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    orbm_count_launches(m, 2);
    int tot = 0;
    for (int c = 0; c < C; c++) if (c == 0 || map_scaled) tot += sm[c];
    if (nmatches) *nmatches = tot;
    return ORB_OK;
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
