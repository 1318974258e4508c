// orb_gba.cu -- sm_100a GlobalBundleAdjustemnt for maps that do not fit the batched LocalBA path (thousands of key frames),
// on one GPU or landmark-partitioned over the GPUs of a node with one exchange step per LM trial.
//
// Reference path (file:line under /root/reference): Optimizer::GlobalBundleAdjustemnt / BundleAdjustment src/Optimizer.cc:62-248
// (all key frames and map points, one optimize(nIterations), Huber kernel thHuber2D when bRobust, no outlier pass) over the same g2o
// machinery as LocalBA (see orb_ba.cu for the per-block citations); the reduced camera system is solved by
// LinearSolverEigen = Eigen::SimplicialLDLT, a SPARSE LDL^T (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:60-121).
//
// Partitioning: key-frame poses are replicated, every rank owns the landmarks `point_id mod world == rank`
// together with their edges.  Landmark blocks are independent given the poses (Schur structure, block_solver.hpp:381-432), so a
// rank builds, for its landmarks only, Hll / bl, its share of Hpp / bp and its share of the reduced camera system
//     Hs = Hpp + lambda I - sum_l B_l Dinv_l B_l^T ,   bs = bp - sum_l B_l Dinv_l bl
// then ONE NCCL all-reduce (sum, FP64, over NVLink) of [Hs | bs] makes the system identical on every rank; the factorisation and the
// LM decision are replicated, the back-substitution and the trial errors are local again, and one all-reduce of three scalars
// (chi2, gain denominator, stop flag) closes the trial.
//
// The reduced camera system is kept as a BLOCK SKYLINE: block row i (one free pose, 6x6 blocks) holds the blocks first(i) .. i, where
// first(i) is the lowest pose that shares a landmark with i on ANY rank (an all-reduce(min) of K numbers at set-up).  Key frames see
// the same landmarks as their neighbours in time plus what loop closures connect, so the envelope of a map is a band with a few long
// rows: 2000 key frames x 8 neighbours = 4.6 MB instead of the 1.15 GB of the dense 11994^2 matrix, which is also what the
// all-reduce ships.  An LDL^T leaves the envelope unchanged (no fill outside it), so the factorisation is hand-written: right-looking
// over block columns, in place, with the forward substitution of the right-hand side carried along (k_sky).  Every sum of the build has
// one owner and a fixed order (pose blocks: CTA per pose over its edge list; Schur blocks: warp per block over its tuple list, built on
// the host from the landmark -> edges structure), so the result is bit-reproducible; no floating-point atomics, no library.
// NCCL is loaded with dlopen so that the library still loads on a box without it.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "orb_common.h"
#include "orb_ba_core.cuh"

#define G_T 128
#define SKY_T 1024             // threads of the factorisation CTA

struct GArgs {
    int nP, nL, nE, K, n;
    const int *e_pose, *e_pt, *e_cam, *pose_free, *pt_off;
    const double *e_obs, *e_info, *cam;
    double *pose[2], *pt[2], *err[2];
    double *rec, *B, *Y, *Hll, *bl, *Hpp, *bp, *x;
    double *Hsum;             // [K][36] pose blocks and [6 K] bp summed over the ranks (computeLambdaInit, computeScale); Hpp / bp keep this rank's share
    double *Hs, *bs;          // block skyline [NB][36] followed by the right-hand side [6 K]: one all-reduce
    double *part;             // per-block partial sums: [nb][2]
    double *red;              // [0] chi  [1] landmark scale  [2] stop flag (sum over ranks)  [3] active edges  [4] pose scale  [5] max diag (poses)  [6] max diag (landmarks)  [7] solve ok
    // skyline structure
    const int *first, *rowptr, *last;        // [K] first block column of row i; [K + 1] block offset of row i; [K] last row whose envelope reaches column j
    long long NB;
    // owner lists
    const int *pose_eoff, *pose_edge;        // CSR: free pose -> its edges (this rank)
    const long long* blk_toff;               // [NB + 1] CSR: block -> tuples
    const int2* blk_tup;                     // (edge of the row pose, edge of the column pose)
    double* panW;                            // [K][36] scratch of the factorisation (W = L D of the current block column)
};

__device__ __forceinline__ double g_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double g_block_sum(double v, double* red) {
    v = g_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < G_T / 32; w++) s += red[w];
    return s;
}

// computeActiveErrors + robust chi2 at estimate `buf` (thread per edge) -> per-block partials
__global__ void __launch_bounds__(G_T) g_errors(GArgs A, int buf, int robust, double delta) {
    __shared__ double red[G_T / 32];
    const int e = blockIdx.x * G_T + threadIdx.x;
    double chi = 0;
    if (e < A.nE) {
        const double* c = A.cam + BA_CAM_STRIDE * (size_t)A.e_cam[e];
        double pc[3], er[2];
        edge_project(A.pose[buf] + 7 * (size_t)A.e_pose[e], A.pt[buf] + 3 * (size_t)A.e_pt[e], c, pc);
        edge_error(pc, c, A.e_obs + 2 * (size_t)e, er);
        A.err[buf][2 * e] = er[0]; A.err[buf][2 * e + 1] = er[1];
        const double c2 = (er[0] * er[0] + er[1] * er[1]) * A.e_info[e];
        chi = robust ? huber_rho0(c2, delta, delta * delta) : c2;
    }
    const double s = g_block_sum(chi, red);
    if (threadIdx.x == 0) { A.part[2 * (size_t)blockIdx.x] = s; A.part[2 * (size_t)blockIdx.x + 1] = 0; }
}

// one CTA: red[slot0] = sum part[.][0], red[slot1] = sum part[.][1] in block order (deterministic)
__global__ void __launch_bounds__(256) g_reduce(GArgs A, int nb, int slot0, int slot1) {
    __shared__ double sa[256], sb[256];
    double a = 0, b = 0;
    const int per = (nb + 255) / 256;
    for (int i = threadIdx.x * per; i < min(nb, (threadIdx.x + 1) * per); i++) { a += A.part[2 * (size_t)i]; b += A.part[2 * (size_t)i + 1]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0, y = 0;
        for (int i = 0; i < 256; i++) { x += sa[i]; y += sb[i]; }
        A.red[slot0] = x;
        if (slot1 >= 0) A.red[slot1] = y;
    }
}

// linearizeOplus + the per-edge part of constructQuadraticForm at estimate `cur` (errors already in err[cur]):
// rec = { Jl[6], W, r0, r1, Jp[12] }, B = Jp^T W Jl (6x3)
__global__ void __launch_bounds__(G_T) g_lin(GArgs A, int cur, int robust, double delta) {
    const int e = blockIdx.x * G_T + threadIdx.x;
    if (e >= A.nE) return;
    const double dsqr = delta * delta;
    const double* c = A.cam + BA_CAM_STRIDE * (size_t)A.e_cam[e];
    const int pi = A.e_pose[e];
    const double* ps = A.pose[cur] + 7 * (size_t)pi;
    const double w = A.e_info[e];
    double pc[3];
    edge_project(ps, A.pt[cur] + 3 * (size_t)A.e_pt[e], c, pc);
    const double e0 = A.err[cur][2 * e], e1 = A.err[cur][2 * e + 1];
    const double X = pc[0], Y = pc[1], Z = pc[2], iz = -1. / Z;
    const double t00 = iz * c[0], t02 = iz * (-X / Z * c[0]), t11 = iz * c[1], t12 = iz * (-Y / Z * c[1]);
    double Jp[12];
    if (A.pose_free[pi] >= 0) {
        const double tJ[12] = {t02 * Y, t00 * Z - t02 * X, -t00 * Y, t00, 0, t02, -t11 * Z + t12 * Y, -t12 * X, t11 * X, 0, t11, t12};
        const double* Ad = c + BA_CAM_ADJ;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 6; k++) s += tJ[i * 6 + k] * Ad[k * 6 + j];
                Jp[i * 6 + j] = s;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 12; i++) Jp[i] = 0;
    }
    double q[4], Rm[9], Jl[6];
    q_mul(c + 4, ps, q);
    q_normalize(q);
    q_to_matrix(q, Rm);
#pragma unroll
    for (int j = 0; j < 3; j++) { Jl[j] = t00 * Rm[j] + t02 * Rm[6 + j]; Jl[3 + j] = t11 * Rm[3 + j] + t12 * Rm[6 + j]; }
    double wr = 1.0;
    if (robust) { const double c2 = (e0 * e0 + e1 * e1) * w; if (c2 > dsqr) wr = delta / sqrt(c2); }
    const double W = wr * w;
    double* R = A.rec + (size_t)BA_REC * e;
#pragma unroll
    for (int j = 0; j < 6; j++) R[j] = Jl[j];
    R[6] = W; R[7] = -w * e0 * wr; R[8] = -w * e1 * wr;
#pragma unroll
    for (int j = 0; j < 12; j++) R[9 + j] = Jp[j];
    double* Bm = A.B + 18 * (size_t)e;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++) Bm[r * 3 + cc] = W * (Jp[r] * Jl[cc] + Jp[6 + r] * Jl[3 + cc]);
}

// thread per landmark: Hll, bl; per-block max |diag|
__global__ void __launch_bounds__(G_T) g_build_lm(GArgs A) {
    __shared__ double red[G_T / 32];
    const int l = blockIdx.x * G_T + threadIdx.x;
    double md = 0;
    if (l < A.nL) {
        double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0, b0 = 0, b1 = 0, b2 = 0;
        for (int e = A.pt_off[l]; e < A.pt_off[l + 1]; e++) {
            const double* R = A.rec + (size_t)BA_REC * e;
            const double a0 = R[0], a1 = R[1], a2 = R[2], c0 = R[3], c1 = R[4], c2 = R[5], W = R[6], r0 = R[7], r1 = R[8];
            h00 += (a0 * a0 + c0 * c0) * W; h01 += (a0 * a1 + c0 * c1) * W; h02 += (a0 * a2 + c0 * c2) * W;
            h11 += (a1 * a1 + c1 * c1) * W; h12 += (a1 * a2 + c1 * c2) * W; h22 += (a2 * a2 + c2 * c2) * W;
            b0 += a0 * r0 + c0 * r1; b1 += a1 * r0 + c1 * r1; b2 += a2 * r0 + c2 * r1;
        }
        double* H = A.Hll + 6 * (size_t)l;
        H[0] = h00; H[1] = h01; H[2] = h02; H[3] = h11; H[4] = h12; H[5] = h22;
        A.bl[3 * (size_t)l] = b0; A.bl[3 * (size_t)l + 1] = b1; A.bl[3 * (size_t)l + 2] = b2;
        md = fmax(fabs(h00), fmax(fabs(h11), fabs(h22)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = md;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0;
        for (int w = 0; w < G_T / 32; w++) m = fmax(m, red[w]);
        atomicMax(reinterpret_cast<unsigned long long*>(A.red + 6), (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like integers; max is order-free
    }
}

// CTA per free pose: this rank's share of Hpp_k = sum Jp^T W Jp and bp_k = sum Jp^T r over the pose's edge list (fixed order)
__global__ void __launch_bounds__(G_T) g_build_pose(GArgs A) {
    __shared__ double s_W[(G_T / 32) * 27];
    const int k = blockIdx.x, tid = threadIdx.x;
    double h[27];
#pragma unroll
    for (int i = 0; i < 27; i++) h[i] = 0;
    for (int t = A.pose_eoff[k] + tid; t < A.pose_eoff[k + 1]; t += G_T) {
        const double* R = A.rec + (size_t)BA_REC * A.pose_edge[t];
        const double W = R[6], r0 = R[7], r1 = R[8];
        double Jp[12];
#pragma unroll
        for (int j = 0; j < 12; j++) Jp[j] = R[9 + j];
        int u = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            h[21 + i] += Jp[i] * r0 + Jp[6 + i] * r1;
#pragma unroll
            for (int j = i; j < 6; j++) h[u++] += (Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j]) * W;
        }
    }
#pragma unroll
    for (int i = 0; i < 27; i++) { const double s = g_warp_sum(h[i]); if ((tid & 31) == 0) s_W[(tid >> 5) * 27 + i] = s; }
    __syncthreads();
    if (tid < 36) {
        const int i = tid / 6, j = tid - 6 * i, lo = i < j ? i : j, hi = i < j ? j : i;
        const int u = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
        double s = 0;
#pragma unroll
        for (int w = 0; w < G_T / 32; w++) s += s_W[w * 27 + u];
        A.Hpp[36 * (size_t)k + tid] = s;
    } else if (tid < 42) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < G_T / 32; w++) s += s_W[w * 27 + 21 + tid - 36];
        A.bp[6 * (size_t)k + tid - 36] = s;
    }
}
// max |diag Hpp| after its all-reduce (one CTA)
__global__ void __launch_bounds__(256) g_pose_maxdiag(GArgs A) {
    __shared__ double sm[256];
    double m = 0;
    for (int i = threadIdx.x; i < 6 * A.K; i += 256) m = fmax(m, fabs(A.Hpp[36 * (size_t)(i / 6) + (i % 6) * 7]));
    sm[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < 256; i++) m = fmax(m, sm[i]); A.red[5] = m; }
}

__device__ __forceinline__ void g_dinv(const double* H, double lambda, double* d) {
    const double m0 = H[0] + lambda, m1 = H[1], m2 = H[2], m4 = H[3] + lambda, m5 = H[4], m8 = H[5] + lambda;
    const double c00 = m4 * m8 - m5 * m5, c01 = m5 * m2 - m1 * m8, c02 = m1 * m5 - m4 * m2;
    const double id = 1.0 / (m0 * c00 + m1 * c01 + m2 * c02);
    d[0] = c00 * id; d[1] = c01 * id; d[2] = c02 * id;
    d[3] = (m0 * m8 - m2 * m2) * id; d[4] = (m2 * m1 - m0 * m5) * id; d[5] = (m0 * m4 - m1 * m1) * id;
}

// thread per (edge, row): Y_e = B_e Dinv
__global__ void __launch_bounds__(256) g_trial(GArgs A, double lambda) {
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    if (q >= 6LL * A.nE) return;
    const int e = (int)(q / 6), r = (int)(q - 6LL * e);
    if (A.pose_free[A.e_pose[e]] < 0) return;
    double d[6];
    g_dinv(A.Hll + 6 * (size_t)A.e_pt[e], lambda, d);
    const double* Br = A.B + 18 * (size_t)e + 3 * r;
    const double x0 = Br[0], x1 = Br[1], x2 = Br[2];
    double* Yr = A.Y + 18 * (size_t)e + 3 * r;
    Yr[0] = x0 * d[0] + x1 * d[1] + x2 * d[2]; Yr[1] = x0 * d[1] + x1 * d[3] + x2 * d[4]; Yr[2] = x0 * d[2] + x1 * d[4] + x2 * d[5];
}

// warp per free pose: this rank's share of the Schur right-hand side, bs_k = bp_k - sum_e Y_e bl  (fixed order)
__global__ void __launch_bounds__(G_T) g_rhs(GArgs A) {
    const int k = blockIdx.x * (G_T / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= A.K) return;
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int t = A.pose_eoff[k] + lane; t < A.pose_eoff[k + 1]; t += 32) {
        const int e = A.pose_edge[t];
        const double* Ym = A.Y + 18 * (size_t)e;
        const double* bl = A.bl + 3 * (size_t)A.e_pt[e];
        const double b0 = bl[0], b1 = bl[1], b2 = bl[2];
#pragma unroll
        for (int r = 0; r < 6; r++) v[r] += Ym[3 * r] * b0 + Ym[3 * r + 1] * b1 + Ym[3 * r + 2] * b2;
    }
#pragma unroll
    for (int r = 0; r < 6; r++) v[r] = g_warp_sum(v[r]);
    if (lane < 6) {
        double s = v[0];
#pragma unroll
        for (int r = 1; r < 6; r++) if (lane == r) s = v[r];
        A.bs[6 * (size_t)k + lane] = A.bp[6 * (size_t)k + lane] - s;
    }
}

// warp per block (i, j) of the skyline: Hs_ij = [i == j] Hpp_i - sum over the block's tuples (edge a of pose i, edge b of pose j) of
// Y_a B_b^T.  Lanes split the tuples (36 accumulators each), fixed shuffle tree: deterministic.  Blocks without tuples are zero
// (fill-in room of the envelope), so nothing has to be cleared between trials.
__global__ void __launch_bounds__(G_T) g_schur(GArgs A) {
    const long long blk = (long long)blockIdx.x * (G_T / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (blk >= A.NB) return;
    double acc[36];
#pragma unroll
    for (int i = 0; i < 36; i++) acc[i] = 0;
    for (long long t = A.blk_toff[blk] + lane; t < A.blk_toff[blk + 1]; t += 32) {
        const int2 ab = A.blk_tup[t];
        double Ya[18], Bb[18];
        const double2* yp = reinterpret_cast<const double2*>(A.Y + 18 * (size_t)ab.x);
        const double2* bp = reinterpret_cast<const double2*>(A.B + 18 * (size_t)ab.y);
#pragma unroll
        for (int q = 0; q < 9; q++) { const double2 y = yp[q], b = bp[q]; Ya[2 * q] = y.x; Ya[2 * q + 1] = y.y; Bb[2 * q] = b.x; Bb[2 * q + 1] = b.y; }
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) acc[6 * r + c] += Ya[3 * r] * Bb[3 * c] + Ya[3 * r + 1] * Bb[3 * c + 1] + Ya[3 * r + 2] * Bb[3 * c + 2];
    }
#pragma unroll
    for (int i = 0; i < 36; i++) acc[i] = g_warp_sum(acc[i]);
    // which pose is this a diagonal block of?  The diagonal block of row i is the last block of the row: blk == rowptr[i + 1] - 1
    double out0 = 0, out1 = 0;             // entries lane and lane + 32
#pragma unroll
    for (int i = 0; i < 36; i++) { if (lane == (i & 31) && i < 32) out0 = acc[i]; if (i >= 32 && lane == i - 32) out1 = acc[i]; }
    // binary search of the row (rowptr is increasing)
    int lo = 0, hi = A.K - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (A.rowptr[mid] <= blk) lo = mid; else hi = mid - 1; }
    const bool diag = blk == (long long)A.rowptr[lo + 1] - 1;
    double* H = A.Hs + 36 * (size_t)blk;
    H[lane] = (diag ? A.Hpp[36 * (size_t)lo + lane] : 0.0) - out0;
    if (lane < 4) H[32 + lane] = (diag ? A.Hpp[36 * (size_t)lo + 32 + lane] : 0.0) - out1;
}

// ------------------------------------------------------------------------------------------------ block-skyline LDL^T
// One CTA.  In place, right-looking over block columns (one free pose = 6 scalar columns per step):
//   1. diagonal block (j, j) + lambda I: scalar LDL^T of the 6x6 by one warp; z_j = L_jj^-1 b_j (forward substitution of the rhs)
//   2. panel: the rows i in (j, last(j)] whose envelope reaches column j: thread per scalar row: W = L D from W L_jj^T = A_ij, L = W D^-1
//   3. trailing update inside the envelope: A_ik -= W_ij L_kj^T for j < k <= i, and b_i -= L_ij z_j
// then backward substitution L^T x = D^-1 z, block column by block column.  red[7] = 1 on success, 0 on a zero / non-finite pivot.
__device__ __forceinline__ double* sky_block(const GArgs& A, int i, int j) { return A.Hs + 36 * ((size_t)A.rowptr[i] + (j - A.first[i])); }

__global__ void __launch_bounds__(SKY_T) k_sky(GArgs A, double lambda) {
    __shared__ int s_rows[SKY_T];                 // rows of the current panel (when they fit; else recomputed from `first`)
    __shared__ int s_nrows, s_ok;
    __shared__ double s_Lj[36], s_dj[6], s_z[6];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = A.K;
    double* b = A.bs;
    if (tid == 0) s_ok = 1;
    __syncthreads();
    for (int j = 0; j < K; j++) {
        double* Djj = sky_block(A, j, j);
        if (tid == 0) s_nrows = 0;
        if (warp == 0) {
            if (lane < 6) Djj[7 * lane] += lambda;
            __syncwarp();
            for (int c = 0; c < 6; c++) {
                const double dc = Djj[7 * c];
                if (dc == 0.0 || !isfinite(dc)) { if (lane == 0) s_ok = 0; break; }
                if (lane > c && lane < 6) Djj[6 * lane + c] /= dc;
                __syncwarp();
                for (int en = lane; en < 36; en += 32) {    // H(r, k) -= L_rc d_c L_kc for c < k <= r < 6
                    const int r = en / 6, k = en - 6 * r;
                    if (k > c && k <= r) Djj[6 * r + k] -= Djj[6 * r + c] * dc * Djj[6 * k + c];
                }
                __syncwarp();
            }
            for (int en = lane; en < 36; en += 32) s_Lj[en] = Djj[en];
            if (lane < 6) s_dj[lane] = Djj[7 * lane];
            __syncwarp();
            if (lane == 0) {                                  // z_j = L_jj^-1 b_j
                double z[6];
                for (int c = 0; c < 6; c++) { double a = b[6 * j + c]; for (int k = 0; k < c; k++) a -= s_Lj[6 * c + k] * z[k]; z[c] = a; }
                for (int c = 0; c < 6; c++) { s_z[c] = z[c]; b[6 * j + c] = z[c]; }
            }
        }
        __syncthreads();
        if (!s_ok) break;
        // rows of the panel
        const int last = A.last[j];
        const int span = last - j;                             // candidate rows j + 1 .. last
        const bool listed = span <= SKY_T;
        if (listed) {
            if (tid < span && A.first[j + 1 + tid] <= j) s_rows[atomicAdd(&s_nrows, 1)] = j + 1 + tid;
            __syncthreads();
        }
        const int nrows = listed ? s_nrows : span;
        // 2. panel: scalar row (row slot, r) per thread
        for (int it = tid; it < 6 * nrows; it += SKY_T) {
            const int rs = it / 6, r = it - 6 * rs;
            const int i = listed ? s_rows[rs] : j + 1 + rs;
            if (!listed && A.first[i] > j) continue;
            double* row = sky_block(A, i, j) + 6 * r;
            double w[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double a = row[c];
#pragma unroll
                for (int k = 0; k < c; k++) a -= w[k] * s_Lj[6 * c + k];
                w[c] = a;
            }
            double* pw = A.panW + 36 * (size_t)(i - j - 1) + 6 * r;
            double bz = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) { const double l = w[c] / s_dj[c]; row[c] = l; pw[c] = w[c]; bz += l * s_z[c]; }
            b[6 * i + r] -= bz;                               // forward substitution of the rhs rides along
        }
        __syncthreads();
        // 3. trailing update: entries (i, k, r, c) with k <= i, both rows in the panel
        if (listed) {
            const int npair = nrows * nrows;                   // (row slot a, row slot b), used when s_rows[b] <= s_rows[a]
            for (int it = tid; it < npair * 36; it += SKY_T) {
                const int pr = it / 36, en = it - 36 * pr, ra = pr / nrows, rb = pr - ra * nrows;
                const int i = s_rows[ra], k = s_rows[rb];
                if (k > i) continue;
                const int r = en / 6, c = en - 6 * r;
                const double* W = A.panW + 36 * (size_t)(i - j - 1) + 6 * r;
                const double* Lk = sky_block(A, k, j) + 6 * c;
                double s = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) s += W[q] * Lk[q];
                sky_block(A, i, k)[en] -= s;
            }
        } else {
            // wide panel (a dense-ish envelope): rows i, k walked directly
            for (int i = j + 1 + warp; i <= last; i += SKY_T / 32) {
                if (A.first[i] > j) continue;
                for (int k = j + 1; k <= i; k++) {
                    if (A.first[k] > j) continue;
                    double* Hik = sky_block(A, i, k);
                    const double* Lk = sky_block(A, k, j);
                    for (int en = lane; en < 36; en += 32) {
                        const int r = en / 6, c = en - 6 * r;
                        const double* W = A.panW + 36 * (size_t)(i - j - 1) + 6 * r;
                        double s = 0;
#pragma unroll
                        for (int q = 0; q < 6; q++) s += W[q] * Lk[6 * c + q];
                        Hik[en] -= s;
                    }
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    if (tid == 0) A.red[7] = ok ? 1.0 : 0.0;
    if (!ok) { for (int i = tid; i < 6 * K; i += SKY_T) A.x[i] = 0.0; return; }
    // ---- backward: x_j = L_jj^-T (D_j^-1 z_j - sum_{i > j} L_ij^T x_i)
    for (int j = K - 1; j >= 0; j--) {
        const int last = A.last[j];
        // partial sums over the rows below: thread (row i, component c)
        double part = 0;
        const int NT6 = (SKY_T / 6) * 6;                       // a thread keeps its component: stride of a multiple of 6
        const int c_own = tid % 6;
        for (int it = tid; tid < NT6 && it < 6 * (last - j); it += NT6) {
            const int i = j + 1 + it / 6;
            if (A.first[i] > j) continue;
            const double* L = sky_block(A, i, j);
            const double* xi = A.x + 6 * (size_t)i;
            double s = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) s += L[6 * r + c_own] * xi[r];
            part += s;
        }
        // reduce over the threads with the same component: fixed order through shared memory
        __shared__ double s_part[SKY_T];
        s_part[tid] = part;
        __syncthreads();
        if (tid < 6) {
            double s = 0;
            const int nthr = min(NT6, 6 * (last - j));
            for (int t = tid; t < nthr; t += 6) s += s_part[t];
            s_z[tid] = b[6 * j + tid] / sky_block(A, j, j)[7 * tid] - s;
        }
        __syncthreads();
        if (tid == 0) {
            const double* Ljj = sky_block(A, j, j);
            double x[6];
            for (int c = 5; c >= 0; c--) { double a = s_z[c]; for (int k = c + 1; k < 6; k++) a -= Ljj[6 * k + c] * x[k]; x[c] = a; }
            for (int c = 0; c < 6; c++) A.x[6 * (size_t)j + c] = x[c];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ banded skyline in shared memory
// The same factorisation for envelopes of at most SKY_WMAX blocks below the diagonal (key frames linked to their neighbours in time):
// the rows j .. j + W that step j touches live in a ring of W + 3 rows in shared memory (block (i, k) at ring[i mod (W + 3)][k mod (W + 1)]),
// row j + W + 2 is fetched with cp.async two steps before it is first touched, row j goes back to global memory while the trailing update
// of step j runs.  A step costs three barriers and no global round trip.  The backward substitution streams block columns the same
// way (three-stage cp.async ring) with x in a ring.
#define SKY_WMAX 23
__device__ __forceinline__ void cp16(double* dst_smem, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// Blocks of the window that lie outside the envelope are explicit zeros: a zero block (i, j) gives L_ij = 0 and zero updates, so the
// steps run over the full band of W rows without consulting `first`.  The sequential chain of a step is what it costs (12 000 scalar
// pivots in a row for 2000 key frames), so: the 6x6 pivot block is factorised by every lane of warp 0 redundantly in registers (no
// exchange), reciprocals (__drcp_rn) instead of divisions, ring slots advanced by compare-and-subtract instead of modulo.
__global__ void __launch_bounds__(SKY_T) k_sky_band(GArgs A, double lambda, int W) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_ok;
    __shared__ double s_Lj[36], s_id[6], s_z[6], s_part[6 * (SKY_WMAX + 1)];
    __shared__ unsigned char s_pa[(SKY_WMAX * (SKY_WMAX + 1)) / 2], s_pb[(SKY_WMAX * (SKY_WMAX + 1)) / 2];   // (a, b <= a) row pairs of the trailing update, ordered by a
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = A.K, R = W + 1, RR = W + 3;
    double* win = sm;                               // [RR][R][36]
    double* panW = win + (size_t)RR * R * 36;       // [W][36]
    double* zb = panW + (size_t)max(W, 1) * 36;     // [RR][6]   right-hand side of the rows in the window
    double* colb = zb + RR * 6;                     // backward: [3][R][36] + [3][6]
    double* colz = colb + 3 * (size_t)R * 36;
    double* xr = colz + 18;                         // backward: x ring [RR][6]
    double* b = A.bs;
    // row i -> ring slot sl (= i mod RR), its diagonal block in column slot ci (= i mod R): blocks first(i) .. i with cp.async, zeros for
    // the columns i - W .. first(i) - 1, and its right-hand side
    auto load_row = [&](int i, int sl, int ci) {
        const int f = A.first[i], nb = i - f + 1;
        const double* src = A.Hs + 36 * (size_t)A.rowptr[i];
        double* rowp = win + (size_t)sl * R * 36;
        if (tid < nb * 18) {
            const int m = tid / 18, piece = tid - 18 * m;
            int ck = ci - (nb - 1 - m);              // column f + m is (i - f - m) left of the diagonal
            if (ck < 0) ck += R;
            cp16(rowp + 36 * ck + 2 * piece, src + 2 * tid);
        } else if (tid >= 512 && tid - 512 < (R - nb) * 18) {
            const int q = tid - 512, m = q / 18, piece = q - 18 * m;
            int ck = ci - nb - m;                    // the columns left of first(i)
            if (ck < 0) ck += R;
            rowp[36 * ck + 2 * piece] = 0.0; rowp[36 * ck + 2 * piece + 1] = 0.0;
        }
        if (tid >= 480 && tid < 483) cp16(zb + sl * 6 + 2 * (tid - 480), b + 6 * (size_t)i + 2 * (tid - 480));
    };
    if (tid == 0) {
        s_ok = 1;
        int p = 0;
        for (int a = 0; a < W; a++) for (int q = 0; q <= a; q++) { s_pa[p] = (unsigned char)a; s_pb[p] = (unsigned char)q; p++; }
    }
    for (int i = 0; i <= min(R, K - 1); i++) load_row(i, i % RR, i % R);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    int sj = 0, cj = 0;                             // j mod RR, j mod R
    for (int j = 0; j < K; j++) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");          // rows up to j + R - 1 have landed (this thread's pieces); only row j + R may be in flight
        __syncthreads();                                               // ... and everybody is done with step j - 1, the write-back of row j - 1 included
        // the row first touched two steps from now goes into the slot that held row j - 1
        if (j + R + 1 < K) load_row(j + R + 1, sj == 0 ? RR - 1 : sj - 1, cj + 1 >= R ? cj + 1 - R : cj + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        double* Djj = win + ((size_t)sj * R + cj) * 36;
        if (warp == 0) {
            // scalar LDL^T of the pivot block, every lane on its own copy (lower triangle in registers)
            double h[6][6], dd[6], id[6];
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int k = 0; k <= r; k++) h[r][k] = Djj[6 * r + k] + (r == k ? lambda : 0.0);
            bool good = true;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                dd[c] = h[c][c];
                if (dd[c] == 0.0 || !isfinite(dd[c])) good = false;
                id[c] = __drcp_rn(dd[c]);
                double t[6];                                           // column c before scaling: L_rc d_c
#pragma unroll
                for (int r = c + 1; r < 6; r++) t[r] = h[r][c];
#pragma unroll
                for (int r = c + 1; r < 6; r++) {
                    const double l = t[r] * id[c];
#pragma unroll
                    for (int k = c + 1; k <= r; k++) h[r][k] -= l * t[k];
                    h[r][c] = l;
                }
            }
            double z[6];
            const double* bj = zb + sj * 6;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double a = bj[c];
#pragma unroll
                for (int k = 0; k < c; k++) a -= h[c][k] * z[k];
                z[c] = a;
            }
            __syncwarp();                                              // every lane has read the pivot block before lane 0 overwrites it
            if (tid == 0) {
                if (!good) s_ok = 0;
#pragma unroll
                for (int r = 0; r < 6; r++) {
#pragma unroll
                    for (int k = 0; k < r; k++) { s_Lj[6 * r + k] = h[r][k]; Djj[6 * r + k] = h[r][k]; }
                    Djj[7 * r] = id[r];                                // the diagonal goes back as 1 / d (what the backward sweep multiplies with)
                    s_id[r] = id[r]; s_z[r] = z[r]; b[6 * (size_t)j + r] = z[r];
                }
            }
        }
        __syncthreads();
        if (!s_ok) break;
        const int nr = min(W, K - 1 - j);                              // rows j + 1 .. j + nr
        if (tid < 6 * nr) {
            const int a = tid / 6, r = tid - 6 * a;
            int si = sj + 1 + a;
            if (si >= RR) si -= RR;
            double* row = win + ((size_t)si * R + cj) * 36 + 6 * r;
            double w[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double v = row[c];
#pragma unroll
                for (int k = 0; k < c; k++) v -= w[k] * s_Lj[6 * c + k];
                w[c] = v;
            }
            double bz = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) { const double l = w[c] * s_id[c]; row[c] = l; panW[36 * a + 6 * r + c] = w[c]; bz += l * s_z[c]; }
            zb[si * 6 + r] -= bz;
        }
        __syncthreads();
        const int np36 = (nr * (nr + 1) / 2) * 36;
        for (int it = tid; it < np36; it += SKY_T) {
            const int pr = it / 36, en = it - 36 * pr, a = s_pa[pr], bq = s_pb[pr];
            int si = sj + 1 + a, sk = sj + 1 + bq, ck = cj + 1 + bq;
            if (si >= RR) si -= RR;
            if (sk >= RR) sk -= RR;
            if (ck >= R) ck -= R;
            const int r = en / 6, c = en - 6 * r;
            const double* Wp = panW + 36 * a + 6 * r;
            const double* Lk = win + ((size_t)sk * R + cj) * 36 + 6 * c;
            double sacc = 0;
#pragma unroll
            for (int q = 0; q < 6; q++) sacc += Wp[q] * Lk[q];
            win[((size_t)si * R + ck) * 36 + en] -= sacc;
        }
        {   // row j is final: back to global memory (the blocks inside its envelope)
            const int f = A.first[j], nb = j - f + 1;
            double* dst = A.Hs + 36 * (size_t)A.rowptr[j];
            const double* rowp = win + (size_t)sj * R * 36;
            for (int q = tid; q < nb * 36; q += SKY_T) {
                const int m = q / 36;
                int ck = cj - (nb - 1 - m);
                if (ck < 0) ck += R;
                dst[q] = rowp[36 * ck + (q - 36 * m)];
            }
        }
        if (++sj == RR) sj = 0;
        if (++cj == R) cj = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const bool ok = s_ok != 0;
    if (tid == 0) A.red[7] = ok ? 1.0 : 0.0;
    if (!ok) { for (int i = tid; i < 6 * K; i += SKY_T) A.x[i] = 0.0; return; }
    // ---- backward: x_j = L_jj^-T (D_j^-1 z_j - sum_{i > j} L_ij^T x_i); block column j = (j, j), (j + 1, j) .. (j + W, j) staged two
    //      steps ahead (zeros where the envelope of a row does not reach column j)
    auto load_col = [&](int j, int stage) {
        double* dst = colb + (size_t)stage * R * 36;
        const int nr = min(W, K - 1 - j);
        if (tid < (nr + 1) * 18) {
            const int a = tid / 18, piece = tid - 18 * a, i = j + a, f = A.first[i];
            if (f <= j) cp16(dst + 36 * a + 2 * piece, A.Hs + 36 * ((size_t)A.rowptr[i] + (j - f)) + 2 * piece);
            else { dst[36 * a + 2 * piece] = 0.0; dst[36 * a + 2 * piece + 1] = 0.0; }
        }
        if (tid >= 480 && tid < 483) cp16(colz + stage * 6 + 2 * (tid - 480), b + 6 * (size_t)j + 2 * (tid - 480));
    };
    int st3 = (K - 1) % 3;                          // stage of column j
    if (K - 1 >= 0) load_col(K - 1, st3);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (K - 2 >= 0) load_col(K - 2, (K - 2) % 3);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int xj = (K - 1) % RR;                          // ring slot of x_j
    for (int j = K - 1; j >= 0; j--) {
        if (j - 2 >= 0) load_col(j - 2, st3 == 2 ? 0 : st3 + 1);       // (j - 2) mod 3 = (j + 1) mod 3
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncthreads();
        const double* col = colb + (size_t)st3 * R * 36;
        const int nr = min(W, K - 1 - j);
        if (tid < 6 * nr) {
            const int a = tid / 6 + 1, c = tid - 6 * (a - 1);
            int xi = xj + a;
            if (xi >= RR) xi -= RR;
            const double* L = col + 36 * a;
            const double* xv = xr + xi * 6;
            double sacc = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) sacc += L[6 * r + c] * xv[r];
            s_part[tid] = sacc;
        }
        __syncthreads();
        if (warp == 0) {                            // every lane redundantly: no exchange
            double x[6];
#pragma unroll
            for (int c = 5; c >= 0; c--) {
                double sacc = 0;
                for (int a = 0; a < nr; a++) sacc += s_part[6 * a + c];
                double v = colz[st3 * 6 + c] * col[7 * c] - sacc;        // col[7 c] = 1 / d_c
#pragma unroll
                for (int k = c + 1; k < 6; k++) v -= col[6 * k + c] * x[k];
                x[c] = v;
            }
            if (tid < 6) {
                double v = x[0];
#pragma unroll
                for (int c = 1; c < 6; c++) if (tid == c) v = x[c];
                xr[xj * 6 + tid] = v; A.x[6 * (size_t)j + tid] = v;
            }
        }
        __syncthreads();
        st3 = st3 == 0 ? 2 : st3 - 1;
        xj = xj == 0 ? RR - 1 : xj - 1;
    }
}

// ------------------------------------------------------------------------------------------------ substructured banded solve
// The single-CTA band factorisation is a chain of 6 K sequential pivots (3.5 us per key frame: 7 ms for 2000).  For long bands the rows are
// cut into P segments separated by W-row separators (W = band width in blocks, so two interiors never touch):
//   k_seg_fwd   CTA per segment: the same windowed LDL^T over the segment's interior columns only.  The rows of the separator BELOW the
//               interior ride along as trailing rows (they end up holding L towards the interior and the interior's Schur term on
//               themselves); the coupling to the separator ABOVE is carried as W extra block columns ("spike"): Zt_j = L_jj^-1 Z_j per
//               interior row, trailing rows Z_i -= L_ij Zt_j, and the Schur terms on that separator C -= Zt^T D^-1 Zt, g -= Zt^T D^-1 z.
//   k_red_asm   the separators' reduced system: a band of width 2 W - 1 over (P - 1) W block rows
//   k_sky_band  ... factorised and solved by the band kernel itself
//   k_seg_bwd   CTA per segment: backward sweep of the interior with the two separator solutions known.
// Sequential depth K / P + (P - 1) W + K / P instead of 2 K; every sum keeps one owner and a fixed order.
#define SEG_T 512
#define SEG_WMAX 11            // reduced band 2 W - 1 <= SKY_WMAX
#define SEG_PMAX 64
struct SegArgs {
    int P, W;
    int s[SEG_PMAX], e[SEG_PMAX];     // interior [s, e) of segment p; separator p = rows [e_p, e_p + W), p < P - 1
    double* Zt;                       // [K][W][36]
    double* SPK;                      // [P][W][W][36]  block (row a of separator p, column c of separator p - 1)
    double* CLL;                      // [P][SEG_NCH][W][W][36]  partial Schur terms of interior p on the separator above it (p - 1)
    double* GL;                       // [P][SEG_NCH][W][6]
    int* ok;                          // [P]
    // reduced system
    double *Rs, *rb, *rx;
    const int *rfirst, *rrowptr, *rlast;
    long long rNB;
};

__global__ void __launch_bounds__(SEG_T) k_seg_fwd(GArgs A, const __grid_constant__ SegArgs G, double lambda) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_ok;
    __shared__ double s_Lj[36], s_id[6], s_z[6], s_zt[SEG_WMAX * 36];
    __shared__ unsigned char s_pa[(SEG_WMAX * (SEG_WMAX + 1)) / 2], s_pb[(SEG_WMAX * (SEG_WMAX + 1)) / 2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, p = blockIdx.x;
    const int W = G.W, R = W + 1, RR = W + 3;
    const int s0 = G.s[p], e0 = G.e[p];
    const bool has_left = p > 0, has_right = p < G.P - 1;
    const int rows_end = has_right ? e0 + W : e0;         // rows this CTA owns: interior + the separator below it
    const int lcol0 = s0 - W;                             // first column of the separator above
    double* win = sm;                                     // [RR][R][36]
    double* zsp = win + (size_t)RR * R * 36;              // [RR][W][36]  spike block row of the rows in the window
    double* panW = zsp + (size_t)RR * W * 36;             // [W][36]
    double* zb = panW + (size_t)W * 36;                   // [RR][6]
    double* b = A.bs;
    auto load_row = [&](int i, int sl, int ci) {
        const int f = A.first[i], nb = i - f + 1;
        const double* src = A.Hs + 36 * (size_t)A.rowptr[i];
        double* rowp = win + (size_t)sl * R * 36;
        double* zrow = zsp + (size_t)sl * W * 36;
        if (tid < nb * 18) {
            const int m = tid / 18, piece = tid - 18 * m, k = f + m;
            if (k >= s0) {
                int ck = ci - (i - k);
                if (ck < 0) ck += R;
                cp16(rowp + 36 * ck + 2 * piece, src + 2 * tid);
            } else if (has_left && k >= lcol0) {
                cp16(zrow + 36 * (k - lcol0) + 2 * piece, src + 2 * tid);
            }
        }
        // explicit zeros for what the envelope (or the segment) does not hold
        for (int q = tid; q < (R + W) * 18; q += SEG_T) {
            const int blk = q / 18, piece = q - 18 * blk;
            if (blk < R) {                                // window column i - W + blk
                const int k = i - W + blk;
                if (k < f || k < s0) {
                    int ck = ci - (W - blk);
                    if (ck < 0) ck += R;
                    rowp[36 * ck + 2 * piece] = 0.0; rowp[36 * ck + 2 * piece + 1] = 0.0;
                }
            } else {                                      // spike column lcol0 + (blk - R)
                const int c = blk - R, k = lcol0 + c;
                if (!(has_left && k >= f && k >= 0)) { zrow[36 * c + 2 * piece] = 0.0; zrow[36 * c + 2 * piece + 1] = 0.0; }
            }
        }
        if (tid >= 480 && tid < 483) cp16(zb + sl * 6 + 2 * (tid - 480), b + 6 * (size_t)i + 2 * (tid - 480));
    };
    if (tid == 0) {
        s_ok = 1;
        int q = 0;
        for (int a = 0; a < W; a++) for (int c = 0; c <= a; c++) { s_pa[q] = (unsigned char)a; s_pb[q] = (unsigned char)c; q++; }
    }
    for (int i = s0; i <= min(s0 + R, rows_end - 1); i++) load_row(i, i % RR, i % R);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    int sj = s0 % RR, cj = s0 % R;
    for (int j = s0; j < e0; j++) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (j + R + 1 < rows_end) load_row(j + R + 1, sj == 0 ? RR - 1 : sj - 1, cj + 1 >= R ? cj + 1 - R : cj + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        double* Djj = win + ((size_t)sj * R + cj) * 36;
        if (warp == 0) {
            double h[6][6], id[6];
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int k = 0; k <= r; k++) h[r][k] = Djj[6 * r + k] + (r == k ? lambda : 0.0);
            bool good = true;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double dc = h[c][c];
                if (dc == 0.0 || !isfinite(dc)) good = false;
                id[c] = __drcp_rn(dc);
                double t[6];
#pragma unroll
                for (int r = c + 1; r < 6; r++) t[r] = h[r][c];
#pragma unroll
                for (int r = c + 1; r < 6; r++) {
                    const double l = t[r] * id[c];
#pragma unroll
                    for (int k = c + 1; k <= r; k++) h[r][k] -= l * t[k];
                    h[r][c] = l;
                }
            }
            double z[6];
            const double* bj = zb + sj * 6;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double a = bj[c];
#pragma unroll
                for (int k = 0; k < c; k++) a -= h[c][k] * z[k];
                z[c] = a;
            }
            __syncwarp();                                              // every lane has read the pivot block before lane 0 overwrites it
            if (tid == 0) {
                if (!good) s_ok = 0;
#pragma unroll
                for (int r = 0; r < 6; r++) {
#pragma unroll
                    for (int k = 0; k < r; k++) { s_Lj[6 * r + k] = h[r][k]; Djj[6 * r + k] = h[r][k]; }
                    Djj[7 * r] = id[r];
                    s_id[r] = id[r]; s_z[r] = z[r]; b[6 * (size_t)j + r] = z[r];
                }
            }
        }
        __syncthreads();
        if (!s_ok) break;
        const int nr = min(W, rows_end - 1 - j);
        if (tid < 6 * nr) {
            const int a = tid / 6, r = tid - 6 * a;
            int si = sj + 1 + a;
            if (si >= RR) si -= RR;
            double* row = win + ((size_t)si * R + cj) * 36 + 6 * r;
            double w[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double v = row[c];
#pragma unroll
                for (int k = 0; k < c; k++) v -= w[k] * s_Lj[6 * c + k];
                w[c] = v;
            }
            double bz = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) { const double l = w[c] * s_id[c]; row[c] = l; panW[36 * a + 6 * r + c] = w[c]; bz += l * s_z[c]; }
            zb[si * 6 + r] -= bz;
        } else if (has_left && tid >= 128 && tid < 128 + 6 * W) {      // Zt_j = L_jj^-1 Z_j, a thread per column of the spike block row
            const int col = tid - 128, c = col / 6, cc = col - 6 * c;
            const double* Zj = zsp + (size_t)sj * W * 36 + 36 * c + cc;
            double zt[6];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double a = Zj[6 * r];
#pragma unroll
                for (int k = 0; k < r; k++) a -= s_Lj[6 * r + k] * zt[k];
                zt[r] = a;
            }
#pragma unroll
            for (int r = 0; r < 6; r++) s_zt[36 * c + 6 * r + cc] = zt[r];
        }
        __syncthreads();
        // ---- everything that only reads the finished column: trailing update, spike rows, Schur terms on the separator above, write-backs
        const int nT = (nr * (nr + 1) / 2) * 36;
        const int nS = has_left ? nr * W * 36 : 0;
        for (int it = tid; it < nT + nS; it += SEG_T) {
            if (it < nT) {
                const int pr = it / 36, en = it - 36 * pr, a = s_pa[pr], bq = s_pb[pr];
                int si = sj + 1 + a, sk = sj + 1 + bq, ck = cj + 1 + bq;
                if (si >= RR) si -= RR;
                if (sk >= RR) sk -= RR;
                if (ck >= R) ck -= R;
                const int r = en / 6, c = en - 6 * r;
                const double* Wp = panW + 36 * a + 6 * r;
                const double* Lk = win + ((size_t)sk * R + cj) * 36 + 6 * c;
                double acc = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) acc += Wp[q] * Lk[q];
                win[((size_t)si * R + ck) * 36 + en] -= acc;
            } else {                                      // Z_i -= L_ij Zt_j
                const int q0 = it - nT, a = q0 / (W * 36), rem = q0 - a * W * 36, c = rem / 36, en = rem - 36 * c, r = en / 6, cc = en - 6 * r;
                int si = sj + 1 + a;
                if (si >= RR) si -= RR;
                const double* Li = win + ((size_t)si * R + cj) * 36 + 6 * r;
                double acc = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) acc += Li[q] * s_zt[36 * c + 6 * q + cc];
                zsp[((size_t)si * W + c) * 36 + en] -= acc;
            }
        }
        {   // row j is final: its blocks at or right of the segment start go back to global memory, and its Zt
            const int f = max(A.first[j], s0), nb = j - f + 1;
            double* dst = A.Hs + 36 * ((size_t)A.rowptr[j] + (f - A.first[j]));
            const double* rowp = win + (size_t)sj * R * 36;
            for (int q = tid; q < nb * 36; q += SEG_T) {
                const int m = q / 36;
                int ck = cj - (nb - 1 - m);
                if (ck < 0) ck += R;
                dst[q] = rowp[36 * ck + (q - 36 * m)];
            }
            if (has_left) for (int q = tid; q < W * 36; q += SEG_T) G.Zt[(size_t)j * W * 36 + q] = s_zt[q];
        }
        if (++sj == RR) sj = 0;
        if (++cj == R) cj = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (tid == 0) G.ok[p] = s_ok;
    // ---- the separator below: updated rows, spike rows and right-hand side; the Schur terms on the separator above
    if (has_right && s_ok) {
        for (int a = 0; a < W; a++) {
            const int i = e0 + a;
            int si = sj + a, ci = cj + a;
            if (si >= RR) si -= RR;
            if (ci >= R) ci -= R;
            const int f = A.first[i], nb = i - f + 1;
            double* dst = A.Hs + 36 * (size_t)A.rowptr[i];
            const double* rowp = win + (size_t)si * R * 36;
            for (int q = tid; q < nb * 36; q += SEG_T) {
                const int m = q / 36;
                int ck = ci - (nb - 1 - m);
                if (ck < 0) ck += R;
                dst[q] = rowp[36 * ck + (q - 36 * m)];
            }
            for (int q = tid; q < W * 36; q += SEG_T) G.SPK[((size_t)p * W + a) * W * 36 + q] = has_left ? zsp[(size_t)si * W * 36 + q] : 0.0;
            if (tid < 6) b[6 * (size_t)i + tid] = zb[si * 6 + tid];
        }
    }
}

// Schur terms of interior p on the separator above it, out of the sequential sweep: C = sum_j Zt_j^T D_j^-1 Zt_j and g = sum_j Zt_j^T D_j^-1 z_j
// over the interior rows.  CTA (p, chunk of rows), thread per entry, rows in order; k_red_asm adds the SEG_NCH partial sums in order.
#define SEG_NCH 4
__global__ void __launch_bounds__(512) k_seg_schur(GArgs A, const __grid_constant__ SegArgs G) {
    const int p = blockIdx.x + 1, ch = blockIdx.y, W = G.W;
    const int s0 = G.s[p], e0 = G.e[p], m = e0 - s0;
    const int j0 = s0 + (int)((long long)m * ch / SEG_NCH), j1 = s0 + (int)((long long)m * (ch + 1) / SEG_NCH);
    double* Cp = G.CLL + ((size_t)p * SEG_NCH + ch) * W * W * 36;
    double* Gp = G.GL + ((size_t)p * SEG_NCH + ch) * W * 6;
    for (int q0 = threadIdx.x; q0 < W * W * 36 + W * 6; q0 += 512) {
        double acc = 0;
        if (q0 < W * W * 36) {
            const int c1 = q0 / (W * 36), rem = q0 - c1 * W * 36, c2 = rem / 36, en = rem - 36 * c2, r = en / 6, cc = en - 6 * r;
            for (int j = j0; j < j1; j++) {
                const double* Z = G.Zt + (size_t)j * W * 36;
                const double* Dj = A.Hs + 36 * ((size_t)A.rowptr[j] + (j - A.first[j]));
#pragma unroll
                for (int q = 0; q < 6; q++) acc += Z[36 * c1 + 6 * q + r] * Dj[7 * q] * Z[36 * c2 + 6 * q + cc];
            }
            Cp[q0] = acc;
        } else {
            const int q1 = q0 - W * W * 36, c = q1 / 6, r = q1 - 6 * c;
            for (int j = j0; j < j1; j++) {
                const double* Z = G.Zt + (size_t)j * W * 36;
                const double* Dj = A.Hs + 36 * ((size_t)A.rowptr[j] + (j - A.first[j]));
#pragma unroll
                for (int q = 0; q < 6; q++) acc += Z[36 * c + 6 * q + r] * Dj[7 * q] * A.bs[6 * (size_t)j + q];
            }
            Gp[q1] = acc;
        }
    }
}

// reduced system of the separators: block row rho = p W + a <-> row e_p + a; block (rho, p W + b) = A(e_p + a, e_p + b) - C^(p+1)[a][b],
// block (rho, (p - 1) W + b) = spike of row e_p + a; right-hand side b(e_p + a) + g^(p+1)[a].  thread per entry.
__global__ void __launch_bounds__(256) k_red_asm(GArgs A, const __grid_constant__ SegArgs G) {
    const int W = G.W, nsep = G.P - 1;
    const long long tot = G.rNB * 36;
    for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < tot; q += (long long)gridDim.x * 256) {
        const long long blk = q / 36;
        const int en = (int)(q - 36 * blk);
        int lo = 0, hi = nsep * W - 1;                       // row of the block (rrowptr is increasing)
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (G.rrowptr[mid] <= blk) lo = mid; else hi = mid - 1; }
        const int rho = lo, col = G.rfirst[rho] + (int)(blk - G.rrowptr[rho]);
        const int p = rho / W, a = rho - p * W, pc = col / W, bq = col - pc * W;
        double v;
        if (pc == p) {
            const int i = G.e[p] + a, k = G.e[p] + bq;
            v = k >= A.first[i] ? A.Hs[36 * ((size_t)A.rowptr[i] + (k - A.first[i])) + en] : 0.0;
            for (int ch = 0; ch < SEG_NCH; ch++) v -= G.CLL[((((size_t)(p + 1) * SEG_NCH + ch) * W + a) * W + bq) * 36 + en];
        } else {
            v = G.SPK[(((size_t)p * W + a) * W + bq) * 36 + en];
        }
        G.Rs[q] = v;
    }
    for (int q = blockIdx.x * 256 + threadIdx.x; q < nsep * W * 6; q += gridDim.x * 256) {
        const int rho = q / 6, r = q - 6 * rho, p = rho / W, a = rho - p * W;
        double v = A.bs[6 * (size_t)(G.e[p] + a) + r];
        for (int ch = 0; ch < SEG_NCH; ch++) v -= G.GL[(((size_t)(p + 1) * SEG_NCH + ch) * W + a) * 6 + r];
        G.rb[q] = v;
    }
}

// CTA per segment: x_j = L_jj^-T (D_j^-1 (z_j - Zt_j x_above) - sum_{i in (j, j + W]} L_ij^T x_i), the separator solutions known
__global__ void __launch_bounds__(256) k_seg_bwd(GArgs A, const __grid_constant__ SegArgs G) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_part[6 * (SEG_WMAX + 1)], s_xl[6 * SEG_WMAX];
    __shared__ int s_fail;
    const int tid = threadIdx.x, warp = tid >> 5, p = blockIdx.x;
    const int W = G.W, R = W + 1, RR = W + 3;
    const int s0 = G.s[p], e0 = G.e[p];
    const bool has_left = p > 0, has_right = p < G.P - 1;
    const int rows_end = has_right ? e0 + W : e0;
    double* colb = sm;                                    // [3][R][36]
    double* colz = colb + 3 * (size_t)R * 36;             // [3][6]
    double* ztb = colz + 18;                              // [3][W][36]
    double* xr = ztb + 3 * (size_t)W * 36;                // [RR][6]
    if (tid == 0) {
        int fail = A.red[7] == 0.0;
        for (int q = 0; q < G.P; q++) fail |= G.ok[q] == 0;
        s_fail = fail;
    }
    __syncthreads();
    if (s_fail) {                                         // a pivot failed somewhere: g2o applies no update (the trial is rejected)
        if (p == 0 && tid == 0) A.red[7] = 0.0;
        for (int i = 6 * s0 + tid; i < 6 * rows_end; i += 256) A.x[i] = 0.0;
        return;
    }
    // separator solutions: above -> s_xl, below -> the x ring (and A.x)
    if (has_left) for (int q = tid; q < 6 * W; q += 256) s_xl[q] = G.rx[(size_t)(p - 1) * W * 6 + q];
    if (has_right)
        for (int q = tid; q < 6 * W; q += 256) {
            const double v = G.rx[(size_t)p * W * 6 + q];
            const int i = e0 + q / 6;
            xr[(i % RR) * 6 + q % 6] = v;
            A.x[6 * (size_t)e0 + q] = v;
        }
    auto load_col = [&](int j, int stage) {
        double* dst = colb + (size_t)stage * R * 36;
        const int nr = min(W, rows_end - 1 - j);
        if (tid < (nr + 1) * 18) {
            const int a = tid / 18, piece = tid - 18 * a, i = j + a, f = A.first[i];
            if (f <= j) cp16(dst + 36 * a + 2 * piece, A.Hs + 36 * ((size_t)A.rowptr[i] + (j - f)) + 2 * piece);
            else { dst[36 * a + 2 * piece] = 0.0; dst[36 * a + 2 * piece + 1] = 0.0; }
        }
        if (has_left) for (int q = tid; q < W * 18; q += 256) cp16(ztb + (size_t)stage * W * 36 + 2 * q, G.Zt + (size_t)j * W * 36 + 2 * q);
        if (tid >= 240 && tid < 243) cp16(colz + stage * 6 + 2 * (tid - 240), A.bs + 6 * (size_t)j + 2 * (tid - 240));
    };
    int st3 = (e0 - 1) % 3;
    if (e0 - 1 >= s0) load_col(e0 - 1, st3);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (e0 - 2 >= s0) load_col(e0 - 2, (e0 - 2) % 3);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int xj = (e0 - 1) % RR;
    for (int j = e0 - 1; j >= s0; j--) {
        if (j - 2 >= s0) load_col(j - 2, st3 == 2 ? 0 : st3 + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncthreads();
        const double* col = colb + (size_t)st3 * R * 36;
        const int nr = min(W, rows_end - 1 - j);
        if (tid < 6 * nr) {
            const int a = tid / 6 + 1, c = tid - 6 * (a - 1);
            int xi = xj + a;
            if (xi >= RR) xi -= RR;
            const double* L = col + 36 * a;
            const double* xv = xr + xi * 6;
            double acc = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) acc += L[6 * r + c] * xv[r];
            s_part[tid] = acc;
        } else if (has_left && tid >= 128 && tid < 134) {       // (Zt_j x_above)[r], r = tid - 128
            const int r = tid - 128;
            const double* Z = ztb + (size_t)st3 * W * 36;
            double acc = 0;
            for (int c = 0; c < W; c++)
#pragma unroll
                for (int cc = 0; cc < 6; cc++) acc += Z[36 * c + 6 * r + cc] * s_xl[6 * c + cc];
            s_part[6 * W + r] = acc;
        }
        __syncthreads();
        if (warp == 0) {
            double x[6];
#pragma unroll
            for (int c = 5; c >= 0; c--) {
                double acc = 0;
                for (int a = 0; a < nr; a++) acc += s_part[6 * a + c];
                const double zl = has_left ? s_part[6 * W + c] : 0.0;
                double v = (colz[st3 * 6 + c] - zl) * col[7 * c] - acc;
#pragma unroll
                for (int k = c + 1; k < 6; k++) v -= col[6 * k + c] * x[k];
                x[c] = v;
            }
            if (tid < 6) {
                double v = x[0];
#pragma unroll
                for (int c = 1; c < 6; c++) if (tid == c) v = x[c];
                xr[xj * 6 + tid] = v; A.x[6 * (size_t)j + tid] = v;
            }
        }
        __syncthreads();
        st3 = st3 == 0 ? 2 : st3 - 1;
        xj = xj == 0 ? RR - 1 : xj - 1;
    }
}

// trial poses: exp(x) * pose for free poses, copy for fixed; pose part of computeScale() -> red[4]
__global__ void __launch_bounds__(256) g_pose_update(GArgs A, int cur, double lambda) {
    __shared__ double sm[256];
    const int ok = A.red[7] != 0.0;
    double sc = 0;
    for (int i = threadIdx.x; i < A.nP; i += 256) {
        const int k = A.pose_free[i];
        const double* src = A.pose[cur] + 7 * (size_t)i;
        double* dst = A.pose[cur ^ 1] + 7 * (size_t)i;
        if (k >= 0 && ok) {
            const double* x = A.x + 6 * (size_t)k;
            se3_oplus(x, src, dst);
            for (int j = 0; j < 6; j++) sc += x[j] * (lambda * x[j] + A.Hsum[36 * (size_t)A.K + 6 * (size_t)k + j]);
        } else {
            for (int q = 0; q < 7; q++) dst[q] = src[q];
        }
    }
    sm[threadIdx.x] = sc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < 256; i++) s += sm[i]; A.red[4] = s; }
}

// thread per landmark: increment, trial point, trial errors and chi2 of its edges, landmark part of computeScale()
__global__ void __launch_bounds__(G_T) g_back(GArgs A, int cur, double lambda, int robust, double delta) {
    __shared__ double red[G_T / 32];
    const int ok = A.red[7] != 0.0;
    const int l = blockIdx.x * G_T + threadIdx.x;
    double chi = 0, sc = 0;
    if (l < A.nL) {
        const double* bl = A.bl + 3 * (size_t)l;
        double x0 = 0, x1 = 0, x2 = 0;
        if (ok) {
            double d[6];
            g_dinv(A.Hll + 6 * (size_t)l, lambda, d);
            x0 = d[0] * bl[0] + d[1] * bl[1] + d[2] * bl[2]; x1 = d[1] * bl[0] + d[3] * bl[1] + d[4] * bl[2]; x2 = d[2] * bl[0] + d[4] * bl[1] + d[5] * bl[2];
            for (int e = A.pt_off[l]; e < A.pt_off[l + 1]; e++) {
                const int k = A.pose_free[A.e_pose[e]];
                if (k < 0) continue;
                const double* Ym = A.Y + 18 * (size_t)e;
                const double* xp = A.x + 6 * (size_t)k;
#pragma unroll
                for (int r = 0; r < 6; r++) { x0 -= Ym[r * 3] * xp[r]; x1 -= Ym[r * 3 + 1] * xp[r]; x2 -= Ym[r * 3 + 2] * xp[r]; }
            }
        }
        const double* po = A.pt[cur] + 3 * (size_t)l;
        const double pn[3] = {po[0] + x0, po[1] + x1, po[2] + x2};
        double* pw = A.pt[cur ^ 1] + 3 * (size_t)l;
        pw[0] = pn[0]; pw[1] = pn[1]; pw[2] = pn[2];
        sc = x0 * (lambda * x0 + bl[0]) + x1 * (lambda * x1 + bl[1]) + x2 * (lambda * x2 + bl[2]);
        const double dsqr = delta * delta;
        for (int e = A.pt_off[l]; e < A.pt_off[l + 1]; e++) {
            const double* c = A.cam + BA_CAM_STRIDE * (size_t)A.e_cam[e];
            double pc[3], er[2];
            edge_project(A.pose[cur ^ 1] + 7 * (size_t)A.e_pose[e], pn, c, pc);
            edge_error(pc, c, A.e_obs + 2 * (size_t)e, er);
            A.err[cur ^ 1][2 * e] = er[0]; A.err[cur ^ 1][2 * e + 1] = er[1];
            const double c2 = (er[0] * er[0] + er[1] * er[1]) * A.e_info[e];
            chi += robust ? huber_rho0(c2, delta, dsqr) : c2;
        }
    }
    const double cs = g_block_sum(chi, red);
    const double ss = g_block_sum(sc, red);
    if (threadIdx.x == 0) { A.part[2 * (size_t)blockIdx.x] = cs; A.part[2 * (size_t)blockIdx.x + 1] = ss; }
}

__global__ void g_outputs(GArgs A, int cur, double* poses_out, double* points_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.nP) {
        double R[9];
        const double* s = A.pose[cur] + 7 * (size_t)i;
        q_to_matrix(s, R);
        double* o = poses_out + 12 * (size_t)i;
        for (int r = 0; r < 3; r++) { o[r * 4] = R[r * 3]; o[r * 4 + 1] = R[r * 3 + 1]; o[r * 4 + 2] = R[r * 3 + 2]; o[r * 4 + 3] = s[4 + r]; }
    }
    for (int j = i; j < 3 * A.nL; j += gridDim.x * blockDim.x) points_out[j] = A.pt[cur][j];
}

// ================================================================================================ NCCL (dlopen)
namespace {
struct Id128 { char internal[128]; };      // ncclUniqueId
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        void* h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (!h) continue;
        g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
        g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
        g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
        g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        if (g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy) { g_nccl.lib = h; return true; }
        dlclose(h);
    }
    return false;
}
const int NCCL_F64 = 8, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3;
}  // namespace

namespace {
// host threads for the set-up passes: ORB_HOST_THREADS (e.g. cores / ranks when several processes share a node), else up to 16
int host_thread_count(int want) {
    static const int cap = []() { const char* e = getenv("ORB_HOST_THREADS"); const int v = e ? atoi(e) : 0; return v > 0 ? std::min(v, 64) : 16; }();
    return std::max(1, std::min({want, cap, (int)std::max(1u, std::thread::hardware_concurrency())}));
}
// fn(t) on threads t = 0 .. T-1
template <class F>
void host_threads(int T, F fn) {
    if (T <= 1) { fn(0); return; }
    std::vector<std::thread> ts;
    for (int t = 0; t < T; t++) ts.emplace_back([&, t]() { fn(t); });
    for (std::thread& th : ts) th.join();
}
}  // namespace

struct orbgba {
    int device = 0, rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    void* comm = nullptr;
    uint8_t* arena = nullptr; size_t arena_cap = 0;
    uint8_t* h_stage = nullptr; size_t h_stage_cap = 0;   // pinned staging buffer of the inputs (grow-only)
    double* h_red = nullptr;        // pinned
    long long launches = 0;
    double allreduce_ms = 0, solve_ms = 0, loop_ms = 0;   // accumulated over the last optimize call (events)
    size_t allreduce_bytes = 0;
    long long sky_blocks = 0;
    int segments = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

static void gba_free(orbgba* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g->comm);
    cudaFree(g->arena);
    if (g->h_red) cudaFreeHost(g->h_red);
    if (g->h_stage) cudaFreeHost(g->h_stage);
    for (cudaEvent_t e : g->ev) if (e) cudaEventDestroy(e);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
}

#define NCCL_CALL(g, call)                                                                                              \
    do {                                                                                                                \
        const int _r = (call);                                                                                          \
        if (_r != 0) ORB_FAIL(ORB_E_CUDA, "NCCL error %d (%s) at %s:%d", _r, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", __FILE__, __LINE__); \
    } while (0)

static int all_reduce(orbgba* g, double* buf, size_t count, int op) {
    if (g->world == 1) return ORB_OK;
    NCCL_CALL(g, g_nccl.AllReduce(buf, buf, count, NCCL_F64, op, g->comm, g->stream));
    g->allreduce_bytes += count * 8;
    return ORB_OK;
}

extern "C" {

int orbba_dist_unique_id(uint8_t* id128) {
    if (!id128) ORB_FAIL(ORB_E_INVALID, "orbba_dist_unique_id: NULL");
    if (!load_nccl()) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_dist_unique_id: libnccl.so.2 not found");
    Id128 id;
    memset(&id, 0, sizeof(id));
    NCCL_CALL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return ORB_OK;
}

int orbba_dist_create(orbgba_t** out, int device, int rank, int world, const uint8_t* id128) {
    if (!out) ORB_FAIL(ORB_E_INVALID, "orbba_dist_create: out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) ORB_FAIL(ORB_E_INVALID, "orbba_dist_create: bad rank / world / id");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_dist_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_dist_create: device %d not present", device);
    cudaDeviceProp prop;
    ORB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_dist_create: device %d is sm_%d%d, the kernels are built for sm_100a only", device, prop.major, prop.minor);
    if (world > 1 && !load_nccl()) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_dist_create: libnccl.so.2 not found");
    ORB_CUDA(cudaSetDevice(device));
    orbgba* g = new (std::nothrow) orbgba();
    if (!g) ORB_FAIL(ORB_E_INVALID, "orbba_dist_create: out of host memory");
    g->device = device; g->rank = rank; g->world = world;
    cudaError_t ce = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaHostAlloc((void**)&g->h_red, 64 * sizeof(double), cudaHostAllocDefault);
    for (int i = 0; i < 6 && ce == cudaSuccess; i++) ce = cudaEventCreate(&g->ev[i]);
    if (ce != cudaSuccess) { int rc = orbhost::check_cuda(ce, "orbba_dist_create", __FILE__, __LINE__); gba_free(g); return rc; }
    if (world > 1) {
        Id128 id;
        memcpy(&id, id128, 128);
        const int r = g_nccl.CommInitRank(&g->comm, world, id, rank);
        if (r != 0) { gba_free(g); ORB_FAIL(ORB_E_CUDA, "orbba_dist_create: ncclCommInitRank failed (%d)", r); }
    }
    *out = g;
    return ORB_OK;
}

void orbba_dist_destroy(orbgba_t* g) { gba_free(g); }
long long orbba_dist_launch_count(const orbgba_t* g) { return g ? g->launches : 0; }
int orbba_dist_timing(const orbgba_t* g, double* allreduce_ms, double* solve_ms, double* allreduce_bytes) {
    if (!g) ORB_FAIL(ORB_E_INVALID, "orbba_dist_timing: NULL handle");
    if (allreduce_ms) *allreduce_ms = g->allreduce_ms;
    if (solve_ms) *solve_ms = g->solve_ms;
    if (allreduce_bytes) *allreduce_bytes = (double)g->allreduce_bytes;
    return ORB_OK;
}
int orbba_dist_loop_ms(const orbgba_t* g, double* loop_ms, long long* skyline_blocks) {
    if (!g) ORB_FAIL(ORB_E_INVALID, "orbba_dist_loop_ms: NULL handle");
    if (loop_ms) *loop_ms = g->loop_ms;
    if (skyline_blocks) *skyline_blocks = g->sky_blocks;
    return ORB_OK;
}

int orbba_dist_segments(const orbgba_t* g) { return g ? g->segments : 0; }

// Optimizer::BundleAdjustment on this rank's shard: ALL poses (replicated, identical on every rank), this rank's landmarks
// (points [n_points][3]) and their edges (edge_point indexes the local landmark array).  Collective: every rank of the
// communicator must call it with the same poses / iterations / huber_delta.  Everything a rank decides on its own (input
// validation, the stop flag) is exchanged before it is acted on, so that no rank leaves the collective sequence alone.
int orbba_dist_optimize(orbgba_t* g, const orbba_problem_t* Q, int iterations, double huber_delta, const volatile uint8_t* stop,
                        double* poses_out, double* points_out, orbba_stats_t* stats) {
    if (!g || !Q) ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: bad argument");
    const int nP = Q->n_poses, nL = Q->n_points, nE = Q->n_edges, nC = Q->n_cams;
    ORB_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = g->stream;
    // ---- local validation; the verdict is summed over the ranks before anybody returns
    const char* bad = nullptr;
    if (nP < 1 || nL < 0 || nE < 0 || nC < 1 || iterations < 0) bad = "bad sizes";
    else if (!Q->poses || !Q->pose_fixed || (nL && !Q->points) || (nE && (!Q->edge_pose || !Q->edge_point || !Q->edge_cam || !Q->edge_obs || !Q->edge_inv_sigma2)) ||
             !Q->cam_K || !Q->cam_ext || !Q->cam_adj) bad = "NULL array";
    if (!bad)
        for (int e = 0; e < nE; e++)
            if (Q->edge_pose[e] < 0 || Q->edge_pose[e] >= nP || Q->edge_point[e] < 0 || Q->edge_point[e] >= nL || Q->edge_cam[e] < 0 || Q->edge_cam[e] >= nC) { bad = "edge indexes out of range"; break; }
    double* d_flag = nullptr;
    if (g->world > 1) {
        if (g->arena_cap < 256) { cudaFree(g->arena); g->arena = nullptr; g->arena_cap = 0; ORB_CUDA(cudaMalloc((void**)&g->arena, 1 << 20)); g->arena_cap = 1 << 20; }
        d_flag = (double*)g->arena;
        g->h_red[0] = bad ? 1.0 : 0.0;
        ORB_CUDA(cudaMemcpyAsync(d_flag, g->h_red, 8, cudaMemcpyHostToDevice, st));
        int rc0 = all_reduce(g, d_flag, 1, NCCL_SUM);
        if (rc0 != ORB_OK) return rc0;
        ORB_CUDA(cudaMemcpyAsync(g->h_red, d_flag, 8, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        if (g->h_red[0] > 0 && !bad) bad = "another rank rejected its shard";
    }
    if (bad) ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: %s", bad);
    // ---- host-side preparation: free-pose numbering, edges grouped by landmark (stable), CSR, envelope
    const bool timing = getenv("ORBGBA_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto tick = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[orbgba rank %d] %-28s %8.2f ms\n", g->rank, what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    std::vector<int> pose_free(nP, -1);
    int K = 0;
    for (int i = 0; i < nP; i++) if (!Q->pose_fixed[i]) pose_free[i] = K++;
    const long long n = 6LL * K;
    std::vector<int> cnt(nL + 1, 0), perm(nE);
    for (int e = 0; e < nE; e++) cnt[Q->edge_point[e] + 1]++;
    for (int l = 0; l < nL; l++) cnt[l + 1] += cnt[l];
    std::vector<int> pt_off(cnt);
    for (int e = 0; e < nE; e++) perm[cnt[Q->edge_point[e]]++] = e;
    // first(i): lowest free pose sharing a landmark with i (this rank's landmarks; min over ranks below)
    std::vector<double> first_d((size_t)std::max(K, 1));
    for (int k = 0; k < K; k++) first_d[k] = k;
    for (int l = 0; l < nL; l++) {
        int m = K;
        for (int s = pt_off[l]; s < pt_off[l + 1]; s++) { const int k = pose_free[Q->edge_pose[perm[s]]]; if (k >= 0) m = std::min(m, k); }
        for (int s = pt_off[l]; s < pt_off[l + 1]; s++) { const int k = pose_free[Q->edge_pose[perm[s]]]; if (k >= 0 && m < first_d[k]) first_d[k] = m; }
    }
    if (g->world > 1 && K > 0) {
        const size_t need = 8 * (size_t)K + 256;
        if (need > g->arena_cap) { cudaFree(g->arena); g->arena = nullptr; g->arena_cap = 0; ORB_CUDA(cudaMalloc((void**)&g->arena, need)); g->arena_cap = need; }
        ORB_CUDA(cudaMemcpyAsync(g->arena, first_d.data(), 8 * (size_t)K, cudaMemcpyHostToDevice, st));
        int rc0 = all_reduce(g, (double*)g->arena, (size_t)K, NCCL_MIN);
        if (rc0 != ORB_OK) return rc0;
        ORB_CUDA(cudaMemcpyAsync(first_d.data(), g->arena, 8 * (size_t)K, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    tick("sort by landmark, envelope");
    std::vector<int> first(std::max(K, 1), 0), rowptr(K + 1, 0), last(std::max(K, 1), 0);
    long long NB = 0;
    for (int k = 0; k < K; k++) { first[k] = (int)first_d[k]; rowptr[k] = (int)NB; NB += k - first[k] + 1; if (NB > 0x3fffffffLL) ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: the envelope of the reduced camera system is too large"); }
    rowptr[K] = (int)NB;
    for (int k = 0; k < K; k++) last[k] = k;
    for (int i = 0; i < K; i++) for (int j = first[i]; j < i && last[j] < i; j++) last[j] = i;   // (rows in ascending order: last[] only grows)
    g->sky_blocks = NB;
    int sky_w = 0;                           // widest row of the envelope (blocks below the diagonal)
    for (int k = 0; k < K; k++) sky_w = std::max(sky_w, k - first[k]);
    const size_t sky_smem = 8 * ((size_t)(sky_w + 3) * (sky_w + 1) * 36 + (size_t)std::max(sky_w, 1) * 36 + (size_t)(sky_w + 3) * 6 + 3 * (size_t)(sky_w + 1) * 36 + 18 + (size_t)(sky_w + 3) * 6);
    if (sky_w <= SKY_WMAX) ORB_CUDA(cudaFuncSetAttribute(k_sky_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sky_smem));
    // substructuring (k_seg_fwd): P segments of >= max(W, 8) interior rows separated by W-row separators.  P ~ sqrt(K / W) balances the
    // interior sweeps (2 K / P steps) against the reduced band ((P - 1) W steps at width 2 W - 1).  ORBGBA_SEGMENTS overrides (0: off).
    SegArgs SG;
    memset(&SG, 0, sizeof(SG));
    int seg_P = 0;
    if (sky_w >= 1 && sky_w <= SEG_WMAX) {
        const int W = sky_w;
        int P = K >= 192 ? (int)lround(sqrt(2.0 * K / (2.5 * W))) : 0;
        if (const char* ev = getenv("ORBGBA_SEGMENTS")) P = atoi(ev);
        P = std::min(P, SEG_PMAX);
        while (P >= 2 && (K - (P - 1) * W) / P < std::max(W, 8)) P--;
        if (P >= 2) {
            seg_P = P;
            SG.P = P; SG.W = W;
            const int inner = K - (P - 1) * W;
            int row = 0;
            for (int q = 0; q < P; q++) {
                const int m = inner / P + (q < inner % P ? 1 : 0);
                SG.s[q] = row; SG.e[q] = row + m;
                row += m + W;
            }
        }
    }
    const int rK = seg_P ? (seg_P - 1) * sky_w : 0, rW = 2 * sky_w - 1;
    std::vector<int> rfirst(std::max(rK, 1), 0), rrowptr(rK + 1, 0);
    long long rNB = 0;
    for (int r = 0; r < rK; r++) { const int q = r / sky_w; rfirst[r] = q > 0 ? (q - 1) * sky_w : 0; rrowptr[r] = (int)rNB; rNB += r - rfirst[r] + 1; }
    rrowptr[rK] = (int)rNB;
    const size_t seg_smem_f = 8 * ((size_t)(sky_w + 3) * (sky_w + 1) * 36 + (size_t)(sky_w + 3) * sky_w * 36 + (size_t)sky_w * 36 + (size_t)(sky_w + 3) * 6);
    const size_t seg_smem_b = 8 * (3 * (size_t)(sky_w + 1) * 36 + 18 + 3 * (size_t)sky_w * 36 + (size_t)(sky_w + 3) * 6);
    const size_t red_smem = 8 * ((size_t)(rW + 3) * (rW + 1) * 36 + (size_t)std::max(rW, 1) * 36 + (size_t)(rW + 3) * 6 + 3 * (size_t)(rW + 1) * 36 + 18 + (size_t)(rW + 3) * 6);
    if (seg_P) {
        ORB_CUDA(cudaFuncSetAttribute(k_seg_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem_f));
        ORB_CUDA(cudaFuncSetAttribute(k_seg_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem_b));
        ORB_CUDA(cudaFuncSetAttribute(k_sky_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(red_smem, sky_smem)));
    }
    g->segments = seg_P;
    // owner lists: edges per free pose; tuples per skyline block
    std::vector<int> pose_eoff(K + 1, 0), pose_edge;
    std::vector<long long> blk_toff((size_t)NB + 1, 0);
    // tuples per block, ordered by landmark: a parallel counting sort.  Thread t owns a contiguous range of landmarks (balanced by the
    // number of tuples), counts its tuples per block, the per-block prefix over (block, thread) gives every thread its write positions,
    // so the order inside a block is the landmark order whatever the thread count (bit-reproducible sums downstream).
    std::vector<long long> lm_tup((size_t)nL + 1, 0);
    {
        for (int s = 0; s < nE; s++) { const int k = pose_free[Q->edge_pose[perm[s]]]; if (k >= 0) pose_eoff[k + 1]++; }
        for (int k = 0; k < K; k++) pose_eoff[k + 1] += pose_eoff[k];
        pose_edge.resize((size_t)pose_eoff[K]);
        std::vector<int> pos(pose_eoff.begin(), pose_eoff.end() - 1);
        for (int s = 0; s < nE; s++) { const int k = pose_free[Q->edge_pose[perm[s]]]; if (k >= 0) pose_edge[pos[k]++] = s; }
        for (int l = 0; l < nL; l++) {
            long long f = 0;
            for (int s = pt_off[l]; s < pt_off[l + 1]; s++) f += pose_free[Q->edge_pose[perm[s]]] >= 0;
            lm_tup[l + 1] = lm_tup[l] + f * (f + 1) / 2;
        }
    }
    const long long nT = lm_tup[nL];
    auto block_of = [&](int ka, int kb) { return (long long)rowptr[ka] + (kb - first[ka]); };   // ka >= kb
    const int nth = host_thread_count(nT > 200000 && NB < (1 << 22) ? 64 : 1);
    std::vector<int> lm_cut;                              // landmark ranges of the threads
    std::vector<std::vector<int>> cntT;                   // [thread][block]
    lm_cut.assign(nth + 1, nL);
    lm_cut[0] = 0;
    for (int t = 1; t < nth; t++) lm_cut[t] = (int)(std::lower_bound(lm_tup.begin(), lm_tup.end(), nT * t / nth) - lm_tup.begin());
    for (int t = 1; t <= nth; t++) lm_cut[t] = std::max(std::min(lm_cut[t], nL), lm_cut[t - 1]);
    lm_cut[nth] = nL;
    cntT.assign(nth, std::vector<int>());
    auto for_tuples = [&](int l0, int l1, auto&& emit) {
        for (int l = l0; l < l1; l++)
            for (int sa = pt_off[l]; sa < pt_off[l + 1]; sa++) {
                const int ka = pose_free[Q->edge_pose[perm[sa]]];
                if (ka < 0) continue;
                for (int sb = pt_off[l]; sb < pt_off[l + 1]; sb++) {
                    const int kb = pose_free[Q->edge_pose[perm[sb]]];
                    if (kb < 0 || kb > ka || (kb == ka && sb != sa)) continue;
                    emit(block_of(ka, kb), sa, sb);
                }
            }
    };
    host_threads(nth, [&](int t) {
        std::vector<int>& c = cntT[t];
        c.assign((size_t)NB, 0);
        for_tuples(lm_cut[t], lm_cut[t + 1], [&](long long bid, int, int) { c[(size_t)bid]++; });
    });
    for (long long b2 = 0; b2 < NB; b2++) {
        long long tot = 0;
        for (int t = 0; t < nth; t++) { const int c = cntT[t][(size_t)b2]; cntT[t][(size_t)b2] = (int)tot; tot += c; }   // -> offset of thread t inside the block
        blk_toff[b2 + 1] = blk_toff[b2] + tot;
    }
    tick("owner lists, tuple lists");
    // ---- layout
    size_t cur_off = 0;
    auto add = [&](size_t b) { const size_t o = (cur_off + 255) & ~(size_t)255; cur_off = o + b; return o; };
    const size_t o_epose = add(4 * (size_t)nE), o_ept = add(4 * (size_t)nE), o_ecam = add(4 * (size_t)nE), o_pfree = add(4 * (size_t)nP), o_ptoff = add(4 * (size_t)(nL + 1));
    const size_t o_eobs = add(16 * (size_t)nE), o_einfo = add(8 * (size_t)nE), o_cam = add(8 * BA_CAM_STRIDE * (size_t)nC);
    const size_t o_pose0 = add(56 * (size_t)nP), o_pt0 = add(24 * (size_t)nL);
    const size_t o_first = add(4 * (size_t)std::max(K, 1)), o_rowptr = add(4 * (size_t)(K + 1)), o_last = add(4 * (size_t)std::max(K, 1));
    const size_t o_peoff = add(4 * (size_t)(K + 1)), o_pedge = add(4 * std::max<size_t>(pose_edge.size(), 1));
    const size_t o_btoff = add(8 * (size_t)(NB + 1)), o_btup = add(8 * (size_t)std::max<long long>(nT, 1));
    const size_t o_rfirst = add(4 * (size_t)std::max(rK, 1)), o_rrowptr = add(4 * (size_t)(rK + 1));
    const size_t staged = add(0);
    const size_t o_pose1 = add(56 * (size_t)nP), o_pt1 = add(24 * (size_t)nL), o_err0 = add(16 * (size_t)nE), o_err1 = add(16 * (size_t)nE);
    const size_t o_rec = add(8 * BA_REC * (size_t)nE), o_B = add(144 * (size_t)nE), o_Y = add(144 * (size_t)nE);
    const size_t o_Hll = add(48 * (size_t)nL), o_bl = add(24 * (size_t)nL);
    const size_t o_Hpp = add(8 * (size_t)(36 + 6) * std::max(K, 1));                 // Hpp | bp: this rank's share
    const size_t o_Hsum = add(8 * (size_t)(36 + 6) * std::max(K, 1));                // Hpp | bp summed over the ranks: one all-reduce per LM iteration
    const size_t o_Hs = add(8 * (size_t)(36 * NB + n + 8));                          // skyline | bs contiguous: one all-reduce
    const size_t o_x = add(8 * (size_t)std::max<long long>(n, 1)), o_panW = add(288 * (size_t)std::max(K, 1));
    const int nbE = std::max(1, (nE + G_T - 1) / G_T), nbL = std::max(1, (nL + G_T - 1) / G_T);
    const size_t o_part = add(16 * (size_t)std::max(nbE, nbL)), o_red = add(64 * 8);
    const size_t o_pout = add(96 * (size_t)nP), o_lout = add(24 * (size_t)std::max(nL, 1));
    const size_t sW = (size_t)std::max(sky_w, 1), sP = (size_t)std::max(seg_P, 1);
    const size_t o_zt = add(seg_P ? 288 * sW * (size_t)K : 8), o_spk = add(seg_P ? 288 * sW * sW * sP : 8), o_cll = add(seg_P ? 288 * SEG_NCH * sW * sW * (sP + 1) : 8);
    const size_t o_gl = add(seg_P ? 48 * SEG_NCH * sW * (sP + 1) : 8), o_segok = add(4 * sP);
    const size_t o_Rs = add(8 * (size_t)(36 * rNB + 8)), o_rb = add(48 * (size_t)std::max(rK, 1)), o_rx = add(48 * (size_t)std::max(rK, 1));
    const size_t total = add(0) + 256;
    if (total > g->arena_cap) {
        cudaFree(g->arena); g->arena = nullptr; g->arena_cap = 0;
        ORB_CUDA(cudaMalloc((void**)&g->arena, total));
        g->arena_cap = total;
    }
    if (staged > g->h_stage_cap) {
        if (g->h_stage) cudaFreeHost(g->h_stage);
        g->h_stage = nullptr; g->h_stage_cap = 0;
        ORB_CUDA(cudaHostAlloc((void**)&g->h_stage, staged + (staged >> 3), cudaHostAllocDefault));
        g->h_stage_cap = staged + (staged >> 3);
    }
    struct { uint8_t* p; uint8_t* data() const { return p; } } H{g->h_stage};
    {   // second pass of the counting sort: the tuples go straight into the pinned buffer
        int2* tup = (int2*)(H.data() + o_btup);
        host_threads(nth, [&](int t) {
            std::vector<int>& c = cntT[t];
            for_tuples(lm_cut[t], lm_cut[t + 1], [&](long long bid, int sa, int sb) { tup[(size_t)(blk_toff[bid] + c[(size_t)bid]++)] = make_int2(sa, sb); });
        });
    }
    int *h_epose = (int*)(H.data() + o_epose), *h_ept = (int*)(H.data() + o_ept), *h_ecam = (int*)(H.data() + o_ecam);
    double *h_eobs = (double*)(H.data() + o_eobs), *h_einfo = (double*)(H.data() + o_einfo), *h_cam = (double*)(H.data() + o_cam), *h_pose0 = (double*)(H.data() + o_pose0);
    const int nth_e = host_thread_count(nE > 100000 ? 64 : 1);
    memset(h_cam, 0, 8 * BA_CAM_STRIDE * (size_t)nC);
    host_threads(nth_e, [&](int t) {
        const int T = nth_e, s0 = (int)((long long)nE * t / T), s1 = (int)((long long)nE * (t + 1) / T);
        for (int s = s0; s < s1; s++) {
            const int e = perm[s];
            h_epose[s] = Q->edge_pose[e]; h_ept[s] = Q->edge_point[e]; h_ecam[s] = Q->edge_cam[e];
            h_eobs[2 * s] = Q->edge_obs[2 * e]; h_eobs[2 * s + 1] = Q->edge_obs[2 * e + 1]; h_einfo[s] = Q->edge_inv_sigma2[e];
        }
    });
    memcpy(H.data() + o_pfree, pose_free.data(), 4 * (size_t)nP);
    memcpy(H.data() + o_ptoff, pt_off.data(), 4 * (size_t)(nL + 1));
    memcpy(H.data() + o_first, first.data(), 4 * (size_t)std::max(K, 1));
    memcpy(H.data() + o_rowptr, rowptr.data(), 4 * (size_t)(K + 1));
    memcpy(H.data() + o_last, last.data(), 4 * (size_t)std::max(K, 1));
    memcpy(H.data() + o_rfirst, rfirst.data(), 4 * (size_t)std::max(rK, 1));
    memcpy(H.data() + o_rrowptr, rrowptr.data(), 4 * (size_t)(rK + 1));
    memcpy(H.data() + o_peoff, pose_eoff.data(), 4 * (size_t)(K + 1));
    if (!pose_edge.empty()) memcpy(H.data() + o_pedge, pose_edge.data(), 4 * pose_edge.size());
    memcpy(H.data() + o_btoff, blk_toff.data(), 8 * (size_t)(NB + 1));
    for (int c = 0; c < nC; c++) {
        double* Dc = h_cam + (size_t)BA_CAM_STRIDE * c;
        for (int i = 0; i < 4; i++) Dc[i] = Q->cam_K[4 * c + i];
        const double* T = Q->cam_ext + 12 * c;
        const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
        q_from_matrix(R, Dc + 4);
        if (Dc[7] < 0) for (int i = 4; i < 8; i++) Dc[i] = -Dc[i];
        const double nn = sqrt(Dc[4] * Dc[4] + Dc[5] * Dc[5] + Dc[6] * Dc[6] + Dc[7] * Dc[7]);
        for (int i = 4; i < 8; i++) Dc[i] /= nn;
        Dc[8] = T[3]; Dc[9] = T[7]; Dc[10] = T[11];
        for (int i = 0; i < 36; i++) Dc[BA_CAM_ADJ + i] = Q->cam_adj[36 * c + i];
    }
    for (int i = 0; i < nP; i++) {
        const double* T = Q->poses + 12 * i;
        double* Dp = h_pose0 + 7 * (size_t)i;
        const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
        q_from_matrix(R, Dp);
        if (Dp[3] < 0) for (int k = 0; k < 4; k++) Dp[k] = -Dp[k];
        const double nn = sqrt(Dp[0] * Dp[0] + Dp[1] * Dp[1] + Dp[2] * Dp[2] + Dp[3] * Dp[3]);
        for (int k = 0; k < 4; k++) Dp[k] /= nn;
        Dp[4] = T[3]; Dp[5] = T[7]; Dp[6] = T[11];
    }
    if (nL) memcpy(H.data() + o_pt0, Q->points, 24 * (size_t)nL);
    tick("staging buffer fill");
    uint8_t* D = g->arena;
    ORB_CUDA(cudaMemcpyAsync(D, H.data(), staged, cudaMemcpyHostToDevice, st));
    GArgs A;
    memset(&A, 0, sizeof(A));
    A.nP = nP; A.nL = nL; A.nE = nE; A.K = K; A.n = (int)n;
    A.e_pose = (const int*)(D + o_epose); A.e_pt = (const int*)(D + o_ept); A.e_cam = (const int*)(D + o_ecam); A.pose_free = (const int*)(D + o_pfree);
    A.pt_off = (const int*)(D + o_ptoff); A.e_obs = (const double*)(D + o_eobs); A.e_info = (const double*)(D + o_einfo); A.cam = (const double*)(D + o_cam);
    A.pose[0] = (double*)(D + o_pose0); A.pose[1] = (double*)(D + o_pose1); A.pt[0] = (double*)(D + o_pt0); A.pt[1] = (double*)(D + o_pt1);
    A.err[0] = (double*)(D + o_err0); A.err[1] = (double*)(D + o_err1);
    A.rec = (double*)(D + o_rec); A.B = (double*)(D + o_B); A.Y = (double*)(D + o_Y);
    A.Hll = (double*)(D + o_Hll); A.bl = (double*)(D + o_bl); A.Hpp = (double*)(D + o_Hpp); A.bp = A.Hpp + 36 * (size_t)std::max(K, 1); A.Hsum = (double*)(D + o_Hsum);
    A.Hs = (double*)(D + o_Hs); A.bs = A.Hs + 36 * (size_t)NB; A.x = (double*)(D + o_x); A.panW = (double*)(D + o_panW);
    A.part = (double*)(D + o_part); A.red = (double*)(D + o_red);
    A.first = (const int*)(D + o_first); A.rowptr = (const int*)(D + o_rowptr); A.last = (const int*)(D + o_last); A.NB = NB;
    A.pose_eoff = (const int*)(D + o_peoff); A.pose_edge = (const int*)(D + o_pedge);
    SG.Zt = (double*)(D + o_zt); SG.SPK = (double*)(D + o_spk); SG.CLL = (double*)(D + o_cll); SG.GL = (double*)(D + o_gl); SG.ok = (int*)(D + o_segok);
    SG.Rs = (double*)(D + o_Rs); SG.rb = (double*)(D + o_rb); SG.rx = (double*)(D + o_rx);
    SG.rfirst = (const int*)(D + o_rfirst); SG.rrowptr = (const int*)(D + o_rrowptr); SG.rNB = rNB;
    GArgs AR = A;                          // the separators' reduced system, as k_sky_band sees it
    AR.K = rK; AR.n = 6 * rK; AR.Hs = SG.Rs; AR.bs = SG.rb; AR.x = SG.rx; AR.first = SG.rfirst; AR.rowptr = SG.rrowptr; AR.NB = rNB;
    A.blk_toff = (const long long*)(D + o_btoff); A.blk_tup = (const int2*)(D + o_btup);
    double* d_pout = (double*)(D + o_pout);
    double* d_lout = (double*)(D + o_lout);
    g->allreduce_ms = 0; g->solve_ms = 0; g->loop_ms = 0; g->allreduce_bytes = 0;
    const bool robust = huber_delta > 0;
    auto read_red = [&](int count) -> int {   // device red[] -> host
        ORB_CUDA(cudaMemcpyAsync(g->h_red, A.red, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        return ORB_OK;
    };
    auto put_stop = [&]() -> int {            // this rank's view of the stop flag -> red[2]; it is summed with the two trial scalars
        const double f = (stop && *stop) ? 1.0 : 0.0;
        g->h_red[32] = f;
        ORB_CUDA(cudaMemcpyAsync(A.red + 2, g->h_red + 32, 8, cudaMemcpyHostToDevice, st));
        return ORB_OK;
    };
    int rc;
    orbba_stats_t S;
    memset(&S, 0, sizeof(S));
    int cur = 0;
    // ---- initial errors: chi2, the global number of edges and the stop flag on entry (collective decision)
    ORB_CUDA(cudaMemsetAsync(A.red, 0, 64 * 8, st));
    g_errors<<<nbE, G_T, 0, st>>>(A, cur, robust, huber_delta);
    g_reduce<<<1, 256, 0, st>>>(A, nbE, 0, -1);
    g->launches += 2;
    {
        g->h_red[33] = (double)nE;
        ORB_CUDA(cudaMemcpyAsync(A.red + 3, g->h_red + 33, 8, cudaMemcpyHostToDevice, st));
        if ((rc = put_stop()) != ORB_OK) return rc;
    }
    if ((rc = all_reduce(g, A.red, 4, NCCL_SUM)) != ORB_OK) return rc;
    if ((rc = read_red(4)) != ORB_OK) return rc;
    double currentChi = g->h_red[0];
    const double totalEdges = g->h_red[3];
    const bool stopped0 = g->h_red[2] > 0;
    S.initial_chi2 = currentChi;
    double lambda = 0, ni = 2, rho = 0;
    int nBad = 0;
    bool ok_iter = !stopped0 && totalEdges > 0;
    bool stopped = stopped0;
    tick("upload, initial errors");
    ORB_CUDA(cudaEventRecord(g->ev[4], st));
    for (int it = 0; it < iterations && ok_iter; it++) {
        if (stopped) break;                  // the value every rank agreed on at the end of the previous trial
        // ---- buildSystem
        ORB_CUDA(cudaMemsetAsync(A.red + 6, 0, 8, st));
        g_lin<<<nbE, G_T, 0, st>>>(A, cur, robust, huber_delta);
        g_build_lm<<<nbL, G_T, 0, st>>>(A);
        if (K > 0) g_build_pose<<<K, G_T, 0, st>>>(A);
        g->launches += 2 + (K > 0);
        // [Hpp | bp] summed over the ranks on a copy (the per-rank shares stay in place for the diagonal blocks / the rhs of the skyline)
        if (K > 0) ORB_CUDA(cudaMemcpyAsync(A.Hsum, A.Hpp, 8 * 42 * (size_t)K, cudaMemcpyDeviceToDevice, st));
        if ((rc = all_reduce(g, A.Hsum, (size_t)42 * K, NCCL_SUM)) != ORB_OK) return rc;
        if (it == 0) {                       // computeLambdaInit: max |diag| of the summed pose blocks and of the landmark blocks
            if ((rc = all_reduce(g, A.red + 6, 1, NCCL_MAX)) != ORB_OK) return rc;   // non-negative doubles: max is order-free
            {
                GArgs T2 = A;
                T2.Hpp = A.Hsum;
                g_pose_maxdiag<<<1, 256, 0, st>>>(T2);
            }
            g->launches++;
            if ((rc = read_red(8)) != ORB_OK) return rc;
            lambda = 1e-5 * std::max(g->h_red[6], g->h_red[5]);     // computeLambdaInit
            ni = 2; nBad = 0;
        }
        const double iniChi = currentChi;
        int qmax = 0;
        rho = 0;
        bool again = true;
        while (again) {
            // ---- setLambda + Schur complement (this rank's landmarks) into the skyline
            if (nE > 0) { g_trial<<<(unsigned)((6LL * nE + 255) / 256), 256, 0, st>>>(A, lambda); g->launches++; }
            if (K > 0) {
                g_rhs<<<(K + G_T / 32 - 1) / (G_T / 32), G_T, 0, st>>>(A);
                g_schur<<<(unsigned)((NB + G_T / 32 - 1) / (G_T / 32)), G_T, 0, st>>>(A);
                g->launches += 2;
            }
            // ---- the exchange step: [Hs | bs] summed over ranks
            ORB_CUDA(cudaEventRecord(g->ev[0], st));
            if ((rc = all_reduce(g, A.Hs, (size_t)(36 * NB + n), NCCL_SUM)) != ORB_OK) return rc;
            ORB_CUDA(cudaEventRecord(g->ev[1], st));
            if (K > 0) {
                if (seg_P) {
                    k_seg_fwd<<<seg_P, SEG_T, seg_smem_f, st>>>(A, SG, lambda);
                    k_seg_schur<<<dim3(seg_P - 1, SEG_NCH), 512, 0, st>>>(A, SG);
                    k_red_asm<<<(unsigned)std::min<long long>((36 * rNB + 255) / 256, 148 * 8), 256, 0, st>>>(A, SG);
                    k_sky_band<<<1, SKY_T, red_smem, st>>>(AR, lambda, rW);
                    k_seg_bwd<<<seg_P, 256, seg_smem_b, st>>>(A, SG);
                    g->launches += 4;
                }
                else if (sky_w <= SKY_WMAX) k_sky_band<<<1, SKY_T, sky_smem, st>>>(A, lambda, sky_w);
                else k_sky<<<1, SKY_T, 0, st>>>(A, lambda);
                g->launches++;
            }
            else ORB_CUDA(cudaMemsetAsync(A.red + 7, 0, 8, st));
            ORB_CUDA(cudaEventRecord(g->ev[2], st));
            // ---- update + trial errors (local), then the scalars
            g_pose_update<<<1, 256, 0, st>>>(A, cur, lambda);
            g_back<<<nbL, G_T, 0, st>>>(A, cur, lambda, robust, huber_delta);
            g_reduce<<<1, 256, 0, st>>>(A, nbL, 0, 1);
            g->launches += 3;
            if ((rc = put_stop()) != ORB_OK) return rc;
            if ((rc = all_reduce(g, A.red, 3, NCCL_SUM)) != ORB_OK) return rc;
            if ((rc = read_red(8)) != ORB_OK) return rc;
            {
                float ms = 0;
                cudaEventElapsedTime(&ms, g->ev[0], g->ev[1]); g->allreduce_ms += ms;
                cudaEventElapsedTime(&ms, g->ev[1], g->ev[2]); g->solve_ms += ms;
            }
            const bool ok = g->h_red[7] != 0.0;
            double tempChi = g->h_red[0];
            if (!ok) tempChi = 1.7976931348623157e308;
            const double scale = (g->h_red[4] + g->h_red[1]) + 1e-3;
            rho = (currentChi - tempChi) / scale;
            S.trials++;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3.0);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                cur ^= 1;
            } else {
                lambda *= ni;
                ni *= 2;
            }
            qmax++;
            stopped = g->h_red[2] > 0;       // summed over the ranks: the same on every rank
            again = rho < 0 && qmax < 10 && !stopped;
        }
        S.iterations++;
        if (qmax == 10 || rho == 0) ok_iter = false;
        else {
            if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
            if (nBad >= 3) ok_iter = false;
        }
    }
    ORB_CUDA(cudaEventRecord(g->ev[5], st));
    g_outputs<<<std::max(1, (std::max(nP, 3 * nL / 8 + 1) + 255) / 256), 256, 0, st>>>(A, cur, d_pout, d_lout);
    g->launches++;
    ORB_CUDA(cudaGetLastError());
    if (poses_out) ORB_CUDA(cudaMemcpyAsync(poses_out, d_pout, 96 * (size_t)nP, cudaMemcpyDeviceToHost, st));
    if (points_out && nL) ORB_CUDA(cudaMemcpyAsync(points_out, d_lout, 24 * (size_t)nL, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, g->ev[4], g->ev[5]); g->loop_ms = ms; }
    tick("LM loop, download");
    S.final_chi2 = currentChi; S.final_lambda = lambda; S.outliers = 0; S.status = stopped0 ? ORB_E_ABORTED : ORB_OK;
    if (stats) *stats = S;
    return S.status;
}

}  // extern "C"
