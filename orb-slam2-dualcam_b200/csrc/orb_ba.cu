// orb_ba.cu -- sm_100a Levenberg-Marquardt bundle adjustment behind the orbba_* C-ABI.
//
// Reference path (file:line under /root/reference):
//   Optimizer::LocalBundleAdjustment src/Optimizer.cc:407-696, Optimizer::BundleAdjustment :70-248
//   EdgeSE3ProjectXYZ::{computeError,linearizeOplus,isDepthPositive} Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:109-169
//   BaseBinaryEdge::constructQuadraticForm .../core/base_binary_edge.hpp:55-120, RobustKernelHuber .../core/robust_kernel_impl.cpp:78-91
//   BlockSolver<6,3>::{buildSystem,setLambda,solve} .../core/block_solver.hpp:354-604 (Schur complement)
//   OptimizationAlgorithmLevenberg::solve .../core/optimization_algorithm_levenberg.cpp:61-189, SE3Quat .../types/se3quat.h
//
// B200 formulation ("flat" batched LM).  A batch of independent problems (one per keyframe / per sequence) is concatenated
// into one edge / landmark / pose index space and advanced in lock-step: one STEP = one LM trial of every unfinished
// problem = six kernel launches over the whole batch; the LM policy (lambda, rho, trial / iteration counters, the
// Huber -> outlier-removal -> plain round switch, the stop flag) lives in a per-problem state block on the device and is
// advanced by the last CTA of the step, so the host never reads anything back between steps.  Everything is FP64 and
// every sum has one owner and a fixed order (no floating-point atomics): results are bit-reproducible.
//
//   k_lin, k_build   only on the first step of a round (Huber round, plain round): robust chi2 of the round's first estimate,
//               outlier marking, and max |diag H| for computeLambdaInit -- exit at once on every other step
//   k_land      CTA / group of whole landmarks (<= 128 edges): linearisation of every edge (residual, Huber weight, Jl), the
//               landmark blocks Hll, bl, (Hll + lambda I)^-1, and per edge the 128-byte record {x y 1/z W V VD} + {r, VD bl};
//               nothing Jacobian-sized is written, the per-edge intermediates never leave the SM                          (per trial)
//   k_pairs     warp / chunk of (edge, edge) tuples of one pose pair and camera pair: partial Schur product in "tJ space"
//               (the diagonal pairs carry Hpp with them: U (VD V^T - W I) U^T); warp / pose: bp and the Schur right-hand side  (per trial)
//   k_solve     CTA / problem       reduced camera system in shared memory, LDL^T, pose increments, exp-map update
//   k_back      thread / landmark   landmark increment, trial errors and chi2 of its edges; last CTA: LM decision
// Blocks are laid out problem-major, so the CTAs resident at any time work on a handful of neighbouring problems and the
// gathers of k_pairs (4x re-use of every Y_e / B_e block) are served by the 126 MB L2.
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "orb_common.h"
#include "orb_ba_core.cuh"

#define BA_TE 128              // threads per edge block
#define BA_TG 128              // threads (= edge slots) per landmark-group block (k_land, k_back)
#define BA_TAB 16              // (pose, camera) table entry: R[9] t[3] fx fy cx cy
#define BA_TL 128              // threads per landmark block
#define BA_TP 128              // threads per pose block (k_build)
#define BA_CH 512              // tuples per chunk (k_pairs)
#define BA_MAXCC 16            // camera pairs per pose pair: rigs of up to 4 cameras
#define BA_TS 512              // threads of k_solve
#define BA_HS_SMEM_N 216       // reduced camera system (packed lower triangle) in shared memory up to 216 x 216 doubles (36 free poses)

struct BAProb {                // static description of one problem inside the batch
    int e0, nE, l0, nL, p0, nP, c0, nC, k0, K, n;
    int pair0, nPairs;         // K (K + 1) / 2 pair slots
    int chunk0, nChunksMax;    // chunk slots (upper bound)
    int item0, nItems;         // k_pairs work items: nChunksMax chunk warps + K pose warps
    int blkI0;                 // first k_pairs block of the problem
    int rt0, CC, pc0, ut0;     // first (pose, camera) rotation, nC * nC, first (pair, camera pair) slot, first (free pose, camera) slot
    int blkE0, nbE, blkL0, nbL;
    int blkG0, nbG;            // landmark-group blocks of k_land
    long long tup0, eof0, hs_off;
    long long bm0; int bmW;    // landmark bitmaps of the (free pose, camera) pairs: bmW words each
};

struct BAState {
    double lambda, ni, currentChi, iniChi, rho, initial_chi2, poseScale;
    unsigned long long maxdiag_bits;
    int cur, last;             // double-buffer index of the accepted estimate / of the errors computed last
    int round, it, qmax, nBad, trials, iterations;
    int need_build, round_start, mark, lambda_pending, solve_ok, skip, done, stopped, aborted;
    unsigned ticket;
    int nChunks, nTuples;
};

struct BABatch {               // kernel argument (by value)
    int nProb;
    const BAProb* prob;
    BAState* state;
    // static
    const int *e_pose, *e_pt, *e_cam;          // global indices
    const double *e_obs, *e_info, *cam;
    const int* pt_off;                          // [Ltot + 1] global CSR (edges are grouped by landmark)
    const int* pose_free;                       // [Ptot] global free index or -1
    const double *pose0, *pt0;
    const int *blkE_prob, *blkL_prob, *item_prob;
    const int4* blkG_desc;                      // k_land block -> {problem, first landmark (global), first edge (global), landmarks | edges << 8}: one load instead of a chain of four
    const int *blkG_prob, *blkG_l0, *blkG_nl;  // k_land block -> problem, first landmark (global), landmarks (whole landmarks, <= BA_TG edges; a landmark with more gets a block of its own)
    const int* free_pose;                       // [Ktot] global pose index of every free pose
    // index built on the device
    int* edge_of;                               // [free pose of the problem][landmark] -> edge << 2 | camera, or -1
    int* e_kf;                                  // per edge: free-pose index of its key frame (global), -1 if fixed: spares the consumers a dependent gather
    const int* blkL_l0;                         // k_back block -> its first landmark (global)
    unsigned* bm;                               // [(free pose, camera) of the problem][bmW]: bit l = the pose observes landmark l with that camera
    int *pair_cnt, *pair_off;                    // per pair: tuples, first tuple
    int *pc_cnt, *pc_off, *pc_fchunk, *pc_nchunk; // per (pair, camera pair): tuples, first tuple, first chunk, chunks
    int *chunk_pair, *chunk_start, *chunk_len;
    int* pairs_counter;                         // work-queue head of the persistent k_pairs
    int4* item_rec;                             // per k_pairs warp (block * 4 + warp), two int4: {problem, first tuple (absolute), tuples, chunk} {camera a, camera b, -, -}
    int2* tuples;
    // dynamic
    double *pose[2], *pt[2], *err[2];
    unsigned char* level;
    double *er;                                 // per edge: X Y Z 1/Z W r0 r1 - (point in the camera frame, weights)
    double *yr, *rw;                            // per edge and trial: {x y 1/z W V[6] VD[6]} (16), {r0 r1 (VD bl)0 (VD bl)1} (4)
    double *tab[2];                             // per estimate buffer and (pose, camera): [R | t] of ext_c * pose and the intrinsics (BA_TAB doubles, one 128-byte line)
    double *ut_u, *ut_b, *ut_y;                 // per (free pose, camera): Schur rhs partial / bp partial in tJ space (6 each), Adj_c x_k (6)
    int* rs_flag;                               // != 0: some problem starts a round on this step (k_lin / k_build have work)
    int* stop_dev;                              // device copy of the caller's stop flag, refreshed once per LM step (k_land) from the mapped host word
    double *Hll, *bl;                           // per landmark: 6 / 3
    double *Hpp, *bp, *bs, *xp;                 // per free pose: 36 / 6 / 6 / 6
    double *partial, *prhs;                     // per chunk: 36 ; per chunk of a diagonal pair: bp and Schur rhs partials in tJ space (6 + 6)
    double *partE, *partL;                      // per edge block: chi, active ; per landmark block: chi, scale
    double* Hs;                                 // reduced systems that do not fit in shared memory
    // outputs
    double *poses_out, *points_out;
    unsigned char* outlier;
    orbba_stats_t* stats;
    int* n_active;                              // mapped host memory
    // run parameters
    int its1, its2;
    double delta, chi2_th;
    const volatile int* stop;
};

// ------------------------------------------------------------------------------------------------ block reductions (fixed order)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int NW>
__device__ __forceinline__ double block_sum(double v, double* red) {   // red: NW doubles of shared memory
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) s += red[w];
    return s;
}
__device__ __forceinline__ double lambda_eff(const BAState& S) {
    return S.lambda_pending ? 1e-5 * __longlong_as_double((long long)S.maxdiag_bits) : S.lambda;
}

// ------------------------------------------------------------------------------------------------ index construction
__global__ void k_edge_of(BABatch A) {
    const int b = blockIdx.x;
    const int p = A.blkE_prob[b];
    const BAProb P = A.prob[p];
    const int e = P.e0 + (b - P.blkE0) * BA_TE + threadIdx.x;
    if (e >= P.e0 + P.nE) return;
    const int k = A.pose_free[A.e_pose[e]];
    A.e_kf[e] = k;
    if (k >= 0) {
        const int l = A.e_pt[e] - P.l0, c = A.e_cam[e] - P.c0;
        A.edge_of[P.eof0 + (long long)(k - P.k0) * P.nL + l] = (e << 2) | c;   // edge and its camera (rigs of <= 4 cameras)
        atomicOr(&A.bm[P.bm0 + ((long long)(k - P.k0) * P.nC + c) * P.bmW + (l >> 5)], 1u << (l & 31));
    }
}
__device__ __forceinline__ void pair_decode(int pid, int K, int& i, int& j) {
    i = 0;
    int rem = pid;
    while (rem >= K - i) { rem -= K - i; i++; }
    j = i + rem;
}
// warp per pose pair (i <= j): number of landmarks observed by both, per camera pair (camera of i's edge, camera of j's edge):
// popcount of the AND of the two landmark bitmaps (a scan of the pose-major edge table took 0.75 ms per 256 windows, this takes 1/10)
__global__ void k_pair_count(BABatch A, const int* blkP_prob, const int* blkP_first) {
    const int p = blkP_prob[blockIdx.x];
    const BAProb P = A.prob[p];
    const int pid = (blockIdx.x - blkP_first[blockIdx.x]) * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (pid >= P.nPairs) return;
    int i, j;
    pair_decode(pid, P.K, i, j);
    int tot = 0;
    for (int q = 0; q < P.CC; q++) {
        const unsigned* bi = A.bm + P.bm0 + ((long long)i * P.nC + q / P.nC) * P.bmW;
        const unsigned* bj = A.bm + P.bm0 + ((long long)j * P.nC + q % P.nC) * P.bmW;
        int c = 0;
        for (int w = lane; w < P.bmW; w += 32) c += __popc(bi[w] & bj[w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) A.pc_cnt[P.pc0 + pid * P.CC + q] = c;
        tot += c;
    }
    if (lane == 0) A.pair_cnt[P.pair0 + pid] = tot;
}
// CTA per problem: tuple offsets of the (pair, camera pair) slots and the chunk table (block-wide exclusive scan over the slots, 256 per sweep)
__global__ void __launch_bounds__(256) k_pair_scan(BABatch A) {
    __shared__ int s_wt[8], s_wc[8], s_carry_t, s_carry_c;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BAProb P = A.prob[p];
    const int nslot = P.nPairs * P.CC;
    if (tid == 0) { s_carry_t = 0; s_carry_c = 0; }
    __syncthreads();
    for (int s0 = 0; s0 < nslot; s0 += 256) {
        const int slot = s0 + tid;
        const int c = slot < nslot ? A.pc_cnt[P.pc0 + slot] : 0;
        const int nch = (c + BA_CH - 1) / BA_CH;
        int it = c, ic = nch;                                 // inclusive scans inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int vt = __shfl_up_sync(0xffffffffu, it, o), vc = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) { it += vt; ic += vc; }
        }
        if (lane == 31) { s_wt[warp] = it; s_wc[warp] = ic; }
        __syncthreads();
        int bt = s_carry_t, bc = s_carry_c;
        for (int w = 0; w < warp; w++) { bt += s_wt[w]; bc += s_wc[w]; }
        const int off = bt + it - c, fch = bc + ic - nch;     // exclusive
        if (slot < nslot) {
            const int pid = slot / P.CC, q = slot - pid * P.CC;
            A.pc_off[P.pc0 + slot] = off;
            A.pc_fchunk[P.pc0 + slot] = fch;
            A.pc_nchunk[P.pc0 + slot] = nch;
            if (q == 0) A.pair_off[P.pair0 + pid] = off;
            int pi, pj;
            pair_decode(pid, P.K, pi, pj);
            for (int k = 0; k < nch; k++) {
                const int nc = fch + k;
                if (nc >= P.nChunksMax) break;
                const int st = off + k * BA_CH, len = min(BA_CH, c - k * BA_CH);
                A.chunk_pair[P.chunk0 + nc] = slot;
                A.chunk_start[P.chunk0 + nc] = st;
                A.chunk_len[P.chunk0 + nc] = len;
                int4* rec = A.item_rec + 2 * ((size_t)P.blkI0 * 4 + nc);      // everything the chunk's warp needs, in one 32-byte read
                rec[0] = make_int4(p, (int)(P.tup0 + st), len, P.chunk0 + nc);
                rec[1] = make_int4(P.c0 + q / P.nC, P.c0 + q % P.nC, pi == pj, P.rt0 + (A.free_pose[P.k0 + pj] - P.p0) * P.nC + q % P.nC);   // .z: diagonal pair (its tuples are (e, e): Hpp rides along); .w: projection-table entry of (pose j, camera b)
            }
        }
        __syncthreads();
        if (tid == 255) { s_carry_t = bt + it; s_carry_c = bc + ic; }
        __syncthreads();
    }
    if (tid == 0) {
        A.state[p].nChunks = min(s_carry_c, P.nChunksMax);
        A.state[p].nTuples = s_carry_t;
    }
}
// tuples of a pair: grouped by camera pair, sorted by landmark inside a group (lanes take 32-landmark words of the ANDed bitmaps,
// an exclusive scan of their popcounts places each lane's tuples)
__global__ void k_pair_fill(BABatch A, const int* blkP_prob, const int* blkP_first) {
    const int p = blkP_prob[blockIdx.x];
    const BAProb P = A.prob[p];
    const int pid = (blockIdx.x - blkP_first[blockIdx.x]) * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (pid >= P.nPairs) return;
    int i, j;
    pair_decode(pid, P.K, i, j);
    const int* Ti = A.edge_of + P.eof0 + (long long)i * P.nL;
    const int* Tj = A.edge_of + P.eof0 + (long long)j * P.nL;
    int2* out = A.tuples + P.tup0;
    for (int q = 0; q < P.CC; q++) {
        if (A.pc_cnt[P.pc0 + pid * P.CC + q] == 0) continue;
        const unsigned* bi = A.bm + P.bm0 + ((long long)i * P.nC + q / P.nC) * P.bmW;
        const unsigned* bj = A.bm + P.bm0 + ((long long)j * P.nC + q % P.nC) * P.bmW;
        int pos = A.pc_off[P.pc0 + pid * P.CC + q];
        for (int w0 = 0; w0 < P.bmW; w0 += 32) {
            const int w = w0 + lane;
            unsigned m = w < P.bmW ? (bi[w] & bj[w]) : 0u;
            const int cnt = __popc(m);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int at = pos + incl - cnt;
            while (m) {
                const int l = 32 * w + (__ffs(m) - 1);
                m &= m - 1;
                out[at++] = make_int2(Ti[l] >> 2, Tj[l] >> 2);
            }
            pos += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// The projection of every edge goes through a per-(pose, camera) table: T = ext_c * pose as [R | t] plus the camera intrinsics
// (se3quat.h:104-128 composition, then toRotationMatrix), 16 doubles = one 128-byte line.  One definition of the reprojection and of
// the residual for every kernel of this file, so that they all produce the same bits.  R is also the factor of
// Jl = -1/z tmp R(ext_c pose) (types_six_dof_expmap.cpp:155-159).  tab[b] always belongs to pose[b].
__device__ __forceinline__ void tab_entry(const double* cam, const double* pose7, double* o) {
    double q[4], Rm[9], t[3];
    q_mul(cam + 4, pose7, q);
    q_normalize(q);
    q_to_matrix(q, Rm);
    se3_map(cam + 4, pose7 + 4, t);
#pragma unroll
    for (int k = 0; k < 9; k++) o[k] = Rm[k];
    o[9] = t[0]; o[10] = t[1]; o[11] = t[2];
    o[12] = cam[0]; o[13] = cam[1]; o[14] = cam[2]; o[15] = cam[3];
}
__device__ __forceinline__ void tab_project(const double* T, const double* X, double* pc) {
    pc[0] = fma(T[0], X[0], fma(T[1], X[1], fma(T[2], X[2], T[9])));
    pc[1] = fma(T[3], X[0], fma(T[4], X[1], fma(T[5], X[2], T[10])));
    pc[2] = fma(T[6], X[0], fma(T[7], X[1], fma(T[8], X[2], T[11])));
}
__device__ __forceinline__ void tab_error(const double* pc, const double* K4, double ox, double oy, double* e2) {   // K4 = fx fy cx cy
    e2[0] = ox - fma(pc[0] / pc[2], K4[0], K4[2]);
    e2[1] = oy - fma(pc[1] / pc[2], K4[1], K4[3]);
}
__device__ __forceinline__ void write_tab(const BABatch& A, const BAProb& P, const double* pose, double* tab, int tid, int nthreads) {
    for (int i = tid; i < P.nP * P.nC; i += nthreads) {
        const int pl = i / P.nC, cl = i - pl * P.nC;
        double o[BA_TAB];
        tab_entry(A.cam + BA_CAM_STRIDE * (size_t)(P.c0 + cl), pose + 7 * (size_t)(P.p0 + pl), o);
        double2* d = reinterpret_cast<double2*>(tab + BA_TAB * (size_t)(P.rt0 + i));
#pragma unroll
        for (int k = 0; k < BA_TAB / 2; k++) d[k] = make_double2(o[2 * k], o[2 * k + 1]);
    }
}
// [R | t] of an entry: six 128-bit gathers.  The intrinsics are read from the camera block instead (a window has two or three cameras,
// so those loads are broadcasts, while the table entries of a warp's edges lie in up to 32 different lines)
__device__ __forceinline__ void load_tab(const double* T, double* o) {
#pragma unroll
    for (int k = 0; k < 6; k++) { const double2 v = reinterpret_cast<const double2*>(T)[k]; o[2 * k] = v.x; o[2 * k + 1] = v.y; }
}

// ------------------------------------------------------------------------------------------------ reset
__global__ void k_reset(BABatch A, int stopped0) {
    const int b = blockIdx.x;
    const int p = A.blkE_prob[b];
    const BAProb P = A.prob[p];
    const int lb = b - P.blkE0, tid = threadIdx.x;
    for (int e = P.e0 + lb * BA_TE + tid; e < min(P.e0 + (lb + 1) * BA_TE, P.e0 + P.nE); e += BA_TE) {
        A.level[e] = 0;
        A.err[0][2 * e] = 0; A.err[0][2 * e + 1] = 0; A.err[1][2 * e] = 0; A.err[1][2 * e + 1] = 0;
    }
    // poses / points: strided over the problem's edge blocks (at least one block exists per problem)
    for (int i = lb * BA_TE + tid; i < 7 * P.nP; i += P.nbE * BA_TE) { const double v = A.pose0[7 * (size_t)P.p0 + i]; A.pose[0][7 * (size_t)P.p0 + i] = v; A.pose[1][7 * (size_t)P.p0 + i] = v; }
    for (int i = lb * BA_TE + tid; i < 3 * P.nL; i += P.nbE * BA_TE) { const double v = A.pt0[3 * (size_t)P.l0 + i]; A.pt[0][3 * (size_t)P.l0 + i] = v; A.pt[1][3 * (size_t)P.l0 + i] = v; }
    if (lb == 0) { write_tab(A, P, A.pose0, A.tab[0], tid, BA_TE); write_tab(A, P, A.pose0, A.tab[1], tid, BA_TE); }
    if (b == 0 && tid == 0) { *A.rs_flag = 1; *A.stop_dev = stopped0; }
    if (lb == 0 && tid == 0) {
        BAState& S = A.state[p];
        const int nChunks = S.nChunks, nTuples = S.nTuples;
        memset(&S, 0, sizeof(S));
        S.nChunks = nChunks; S.nTuples = nTuples;
        S.need_build = 1; S.round_start = 1; S.lambda_pending = 1; S.ni = 2;
        if (stopped0) { S.done = 1; S.aborted = 1; }
        if (P.nE == 0) S.done = 1;
    }
}

// ------------------------------------------------------------------------------------------------ per-edge geometry
// Nothing Jacobian-sized is stored per edge.  k_lin keeps only the point in the camera frame and the weights
//     er[e] = { X, Y, Z, 1/Z, W, r0, r1, - }      (W = rho1 * invSigma2, r = -invSigma2 * e * rho1; all zero for a level-1 edge)
// and every consumer rebuilds what it needs from them (FP64 flops are cheap next to 300 bytes per edge and launch):
//     tJ_e (2x6) = -1/z * tmp * [-skew(p) | I]                       the pose Jacobian BEFORE the camera adjoint:  Jp_e = tJ_e * Adj_c
//     Jl_e (2x3) = -1/z * tmp * R(ext_c * pose)                      (types_six_dof_expmap.cpp:136-159)
// The 6x6 adjoint of a camera is the same for all its edges, so every sum over edges is formed in "tJ space" per camera (or per
// camera pair) and the adjoint is applied once to the sum:  sum_e Jp_e^T X_e Jp'_e = Adj_c^T (sum_e tJ_e^T X_e tJ'_e) Adj_c'.
__device__ __forceinline__ void edge_tj(double X, double Y, double Z, double iz, double fx, double fy, double* tJ, double* t4) {
    const double t00 = -iz * fx, t02 = iz * iz * X * fx, t11 = -iz * fy, t12 = iz * iz * Y * fy;
    t4[0] = t00; t4[1] = t02; t4[2] = t11; t4[3] = t12;
    tJ[0] = t02 * Y; tJ[1] = t00 * Z - t02 * X; tJ[2] = -t00 * Y; tJ[3] = t00; tJ[4] = 0; tJ[5] = t02;
    tJ[6] = -t11 * Z + t12 * Y; tJ[7] = -t12 * X; tJ[8] = t11 * X; tJ[9] = 0; tJ[10] = t11; tJ[11] = t12;
}
__device__ __forceinline__ void edge_jl(const double* t4, const double* R, double* Jl) {
#pragma unroll
    for (int j = 0; j < 3; j++) { Jl[j] = t4[0] * R[j] + t4[1] * R[6 + j]; Jl[3 + j] = t4[2] * R[3 + j] + t4[3] * R[6 + j]; }
}
// (Hll + lambda I)^-1 of a landmark: cofactor inverse of the symmetric 3x3 (Eigen's fixed-size inverse in
// BlockSolver::solve, block_solver.hpp:381-395).  d = {d00, d01, d02, d11, d12, d22}
__device__ __forceinline__ void landmark_dinv(const double* H, double lambda, double* d) {
    const double m0 = H[0] + lambda, m1 = H[1], m2 = H[2], m4 = H[3] + lambda, m5 = H[4], m8 = H[5] + lambda;
    const double c00 = m4 * m8 - m5 * m5, c01 = m5 * m2 - m1 * m8, c02 = m1 * m5 - m4 * m2;
    const double id = 1.0 / (m0 * c00 + m1 * c01 + m2 * c02);
    d[0] = c00 * id; d[1] = c01 * id; d[2] = c02 * id;
    d[3] = (m0 * m8 - m2 * m2) * id; d[4] = (m2 * m1 - m0 * m5) * id; d[5] = (m0 * m4 - m1 * m1) * id;
}
__device__ __forceinline__ void load8(const double* p, double* o) {   // 64-byte record, 16-byte aligned
#pragma unroll
    for (int i = 0; i < 4; i++) { const double2 q = reinterpret_cast<const double2*>(p)[i]; o[2 * i] = q.x; o[2 * i + 1] = q.y; }
}
__device__ __forceinline__ void load6(const double* p, double* o) {   // 48-byte record, 16-byte aligned
#pragma unroll
    for (int i = 0; i < 3; i++) { const double2 q = reinterpret_cast<const double2*>(p)[i]; o[2 * i] = q.x; o[2 * i + 1] = q.y; }
}

// ------------------------------------------------------------------------------------------------ k_lin
// (k_lin and k_build run over a capped grid with a block-stride loop: on the steps where no problem starts a round -- all but two of the
// fifteen of a LocalBundleAdjustment -- the launch costs one wave of CTAs that read a flag, not 60 000 empty ones)
__device__ __forceinline__ void k_lin_body(const BABatch& A, int b) {
    __shared__ double red[BA_TE / 32];
    const int p = A.blkE_prob[b];
    const BAState& S = A.state[p];
    if (S.done || !S.round_start) return;
    const BAProb& P = A.prob[p];
    const int tid = threadIdx.x;
    const int e = P.e0 + (b - P.blkE0) * BA_TE + tid;
    const bool valid = e < P.e0 + P.nE;
    const int cur = S.cur;
    const bool robust = S.round == 0 && A.delta > 0;
    const double delta = A.delta, dsqr = delta * delta;
    double chi = 0, act = 0;
    if (valid) {
        const double* T = A.tab[cur] + BA_TAB * (size_t)(P.rt0 + (A.e_pose[e] - P.p0) * P.nC + (A.e_cam[e] - P.c0));
        const double w = A.e_info[e];
        double pc[3];
        tab_project(T, A.pt[cur] + 3 * (size_t)A.e_pt[e], pc);
        int lvl = A.level[e];
        if (S.mark) {   // e->chi2() > th || !e->isDepthPositive() -> level 1  (src/Optimizer.cc:598-613); chi2 from the errors computed last
            const double l0 = A.err[S.last][2 * e], l1 = A.err[S.last][2 * e + 1];
            lvl = ((l0 * l0 + l1 * l1) * w > A.chi2_th || !(pc[2] > 0.0)) ? 1 : 0;
            A.level[e] = (unsigned char)lvl;
            if (lvl) { A.err[0][2 * e] = l0; A.err[0][2 * e + 1] = l1; A.err[1][2 * e] = l0; A.err[1][2 * e + 1] = l1; }
        }
        double2* R = reinterpret_cast<double2*>(A.er + 8 * (size_t)e);
        if (!lvl) {
            double er[2];
            if (S.round_start) {
                tab_error(pc, A.cam + BA_CAM_STRIDE * (size_t)A.e_cam[e], A.e_obs[2 * (size_t)e], A.e_obs[2 * (size_t)e + 1], er);
                A.err[cur][2 * e] = er[0]; A.err[cur][2 * e + 1] = er[1];
                const double c2 = (er[0] * er[0] + er[1] * er[1]) * w;
                chi = robust ? huber_rho0(c2, delta, dsqr) : c2;
            } else {
                er[0] = A.err[cur][2 * e]; er[1] = A.err[cur][2 * e + 1];
            }
            act = 1;
            double wr = 1.0;
            if (robust) {
                const double c2 = (er[0] * er[0] + er[1] * er[1]) * w;
                if (c2 > dsqr) wr = delta / sqrt(c2);
            }
            R[0] = make_double2(pc[0], pc[1]);
            R[1] = make_double2(pc[2], 1.0 / pc[2]);
            R[2] = make_double2(wr * w, -w * er[0] * wr);
            R[3] = make_double2(-w * er[1] * wr, 0.0);
        } else {
            R[0] = make_double2(0.0, 0.0);
            R[1] = make_double2(1.0, 1.0);
            R[2] = make_double2(0.0, 0.0);
            R[3] = make_double2(0.0, 0.0);
        }
    }
    const double cs = block_sum<BA_TE / 32>(chi, red);
    const double as = block_sum<BA_TE / 32>(act, red);
    if (tid == 0) { A.partE[2 * (size_t)b] = cs; A.partE[2 * (size_t)b + 1] = as; }
    if (b == P.blkE0 && tid == 0 && S.round_start) A.state[p].maxdiag_bits = 0ull;
}
__global__ void __launch_bounds__(BA_TE) k_lin(BABatch A, int nb) {
    if (*A.rs_flag == 0) return;                      // no problem starts a round on this step
    for (int b = blockIdx.x; b < nb; b += gridDim.x) { k_lin_body(A, b); __syncthreads(); }
}

// ------------------------------------------------------------------------------------------------ k_build
// blocks [0, nLandmarkBlocks): thread per landmark -> Hll, bl ;  blocks beyond: CTA per free pose -> Hpp, bp
// Two instantiations, launched back to back: the landmark part needs half the registers of the pose part, and as one kernel it ran at
// the pose part's occupancy (126 registers, 16 warps per SM).
template <int PART>
__device__ __forceinline__ void k_build_body(const BABatch& A, int blk, const int* pose_prob) {
    __shared__ double s_N[27], s_W[(BA_TP / 32) * 27];
    const int tid = threadIdx.x;
    if (PART == 0) {
        const int b = blk;
        const int p = A.blkL_prob[b];
        const BAState& S = A.state[p];
        if (S.done || !S.round_start) return;
        const BAProb& P = A.prob[p];
        const int l = P.l0 + (b - P.blkL0) * BA_TL + tid;
        double md = 0;
        if (l < P.l0 + P.nL) {
            double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0, b0 = 0, b1 = 0, b2 = 0;
            for (int e = A.pt_off[l]; e < A.pt_off[l + 1]; e++) {
                double r[8], t4[4], tJ[12], Jl[6];
                load8(A.er + 8 * (size_t)e, r);
                const int cg = A.e_cam[e];
                const double* c = A.cam + BA_CAM_STRIDE * (size_t)cg;
                edge_tj(r[0], r[1], r[2], r[3], c[0], c[1], tJ, t4);
                edge_jl(t4, A.tab[S.cur] + BA_TAB * (size_t)(P.rt0 + (A.e_pose[e] - P.p0) * P.nC + (cg - P.c0)), Jl);
                const double a0 = Jl[0], a1 = Jl[1], a2 = Jl[2], c0 = Jl[3], c1 = Jl[4], c2 = Jl[5], W = r[4], r0 = r[5], r1 = r[6];
                h00 += (a0 * a0 + c0 * c0) * W; h01 += (a0 * a1 + c0 * c1) * W; h02 += (a0 * a2 + c0 * c2) * W;
                h11 += (a1 * a1 + c1 * c1) * W; h12 += (a1 * a2 + c1 * c2) * W; h22 += (a2 * a2 + c2 * c2) * W;
                b0 += a0 * r0 + c0 * r1; b1 += a1 * r0 + c1 * r1; b2 += a2 * r0 + c2 * r1;
            }
            double* H = A.Hll + 6 * (size_t)l;
            H[0] = h00; H[1] = h01; H[2] = h02; H[3] = h11; H[4] = h12; H[5] = h22;
            A.bl[3 * (size_t)l] = b0; A.bl[3 * (size_t)l + 1] = b1; A.bl[3 * (size_t)l + 2] = b2;
            md = fmax(fabs(h00), fmax(fabs(h11), fabs(h22)));
        }
        if (S.it == 0) {   // computeLambdaInit: max |diag| over all free vertices
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
            if ((tid & 31) == 0 && md > 0) atomicMax(&A.state[p].maxdiag_bits, (unsigned long long)__double_as_longlong(md));
        }
        return;
    }
    // ---- pose part: per camera c, N_c = sum tJ^T W tJ and u_c = sum tJ^T r over the pose's edges seen by camera c (they are the
    //      (c, c) slice of the diagonal pair's tuple list), then Hpp += Adj_c^T N_c Adj_c, bp += Adj_c^T u_c
    const int kg = blk;                                // global free-pose index
    const int p = pose_prob[kg];
    const BAState& S = A.state[p];
    if (S.done || !S.round_start) return;
    const BAProb& P = A.prob[p];
    const int k = kg - P.k0;
    const int pidd = k * P.K - k * (k - 1) / 2;         // diagonal pair (k, k)
    double out = 0;                                     // thread t < 36: Hpp[t / 6][t % 6];  36 <= t < 42: bp[t - 36]
    for (int cl = 0; cl < P.nC; cl++) {
        const int pc = P.pc0 + pidd * P.CC + cl * P.nC + cl;
        const int cnt = A.pc_cnt[pc];
        if (cnt == 0) continue;                         // uniform over the CTA
        const int2* T = A.tuples + P.tup0 + A.pc_off[pc];
        const double* c = A.cam + BA_CAM_STRIDE * (size_t)(P.c0 + cl);
        const double fx = c[0], fy = c[1];
        double h[27];
#pragma unroll
        for (int i = 0; i < 27; i++) h[i] = 0;
        for (int t = tid; t < cnt; t += BA_TP) {
            double r[8], t4[4], tJ[12];
            load8(A.er + 8 * (size_t)T[t].x, r);
            edge_tj(r[0], r[1], r[2], r[3], fx, fy, tJ, t4);
            const double W = r[4], r0 = r[5], r1 = r[6];
            int u = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                h[21 + i] += tJ[i] * r0 + tJ[6 + i] * r1;
#pragma unroll
                for (int j = i; j < 6; j++) h[u++] += (tJ[i] * tJ[j] + tJ[6 + i] * tJ[6 + j]) * W;
            }
        }
        // fixed-order reduction: shuffle tree inside each warp, then the four warp sums in warp order
#pragma unroll
        for (int i = 0; i < 27; i++) { const double s = warp_sum(h[i]); if ((tid & 31) == 0) s_W[(tid >> 5) * 27 + i] = s; }
        __syncthreads();
        if (tid < 27) {
            double s = 0;
#pragma unroll
            for (int w = 0; w < BA_TP / 32; w++) s += s_W[w * 27 + tid];
            s_N[tid] = s;
        }
        __syncthreads();
        const double* Ad = c + BA_CAM_ADJ;
        if (tid < 36) {
            const int i = tid / 6, j = tid - 6 * i;
            double s = 0;
            for (int a = 0; a < 6; a++) {
                double row = 0;                          // (N Adj)[a][j]
                for (int q = 0; q < 6; q++) {
                    const int lo = a < q ? a : q, hi = a < q ? q : a;
                    row += s_N[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)] * Ad[q * 6 + j];
                }
                s += Ad[a * 6 + i] * row;
            }
            out += s;
        } else if (tid < 42) {
            const int i = tid - 36;
            double s = 0;
            for (int a = 0; a < 6; a++) s += Ad[a * 6 + i] * s_N[21 + a];
            out += s;
        }
        __syncthreads();
    }
    if (tid < 36) {
        A.Hpp[36 * (size_t)kg + tid] = out;
        if (S.it == 0 && tid % 7 == 0 && fabs(out) > 0) atomicMax(&A.state[p].maxdiag_bits, (unsigned long long)__double_as_longlong(fabs(out)));
    } else if (tid < 42) {
        A.bp[6 * (size_t)kg + tid - 36] = out;
    }
}
template <int PART>
__global__ void __launch_bounds__(BA_TL) k_build(BABatch A, int nb, const int* pose_prob) {
    if (*A.rs_flag == 0) return;                      // only the first step of a round needs max |diag H| (computeLambdaInit)
    for (int b = blockIdx.x; b < nb; b += gridDim.x) { k_build_body<PART>(A, b, pose_prob); __syncthreads(); }
}

// ------------------------------------------------------------------------------------------------ k_land
// One CTA = a group of WHOLE landmarks with at most BA_TG edges between them (edges are stored grouped by landmark), one LM trial.
// Everything that is local to a landmark happens here without leaving the SM:
//   phase A  thread / edge      reprojection at pose[cur], pt[cur]; residual (the errors of the accepted estimate, err[cur]); Huber
//                               weight; Jl = -1/z tmp R(ext_c pose) (types_six_dof_expmap.cpp:136-159); the edge's terms of Hll, bl
//   phase B  thread / landmark  Hll, bl summed in edge order (fixed order, no atomics); D = (Hll + lambda I)^-1, D bl
//   phase C  thread / edge      the 6x3 block of an edge in tJ space factors through the 2-d residual:
//                                   Bt_e = tJ_e^T W_e Jl_e = U_e V_e          U_e = tJ_e^T (6x2, polynomials in x = X/Z, y = Y/Z, w = 1/Z),  V_e = W_e Jl_e (2x3)
//                                   Yt_e = Bt_e D = U_e VD_e                  VD_e = V_e D (2x3)
//                               so the Schur product of a tuple is U_a (VD_a V_b^T) U_b^T and the per-edge record is
//                                   yr[e] = { x, y, w, W, V[6], VD[6] }   (128 bytes = 4 sectors; k_pairs gathers x y w W + VD of edge a and
//                                                                           x y w W + V of edge b, three sectors each; k_back reads x y w W + V)
//                                   rw[e] = { r0, r1, (VD bl)0, (VD bl)1 } (bp and the Schur right-hand side, summed per pose in k_pairs)
//                               staged in shared memory and written out coalesced.
// A landmark observed by more than BA_TG key frames has a block of its own (`big` path: strided loops and block reductions).
#define BA_YR 16
#define BA_YRS 18                 // row stride of the record in the staging buffer (a 128-byte stride would put every lane on the same banks)

struct EdgeLin { double x, y, iz, W, r0, r1, Jl[6]; bool free_pose; };

// linearisation of edge e at the accepted estimate (level-1 edges and their zero weights included)
__device__ __forceinline__ void edge_lin(const BABatch& A, const BAProb& P, int e, int cur, bool robust, double delta, double dsqr,
                                         const double* tb, int ts, EdgeLin& L) {
    const int pg = A.e_pose[e], cg = A.e_cam[e];
    double T[12], pc[3];
    const double* K4 = A.cam + BA_CAM_STRIDE * (size_t)cg;
    load_tab(tb + ts * ((pg - P.p0) * P.nC + (cg - P.c0)), T);
    tab_project(T, A.pt[cur] + 3 * (size_t)A.e_pt[e], pc);
    L.free_pose = A.e_kf[e] >= 0;
    if (A.level[e]) {                                     // removed from the optimisation (src/Optimizer.cc:598-613): contributes nothing
        L.x = 0; L.y = 0; L.iz = 1; L.W = 0; L.r0 = 0; L.r1 = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) L.Jl[j] = 0;
        return;
    }
    const double w = A.e_info[e];
    const double2 er = reinterpret_cast<const double2*>(A.err[cur])[e];
    double wr = 1.0;
    if (robust) {
        const double c2 = (er.x * er.x + er.y * er.y) * w;
        if (c2 > dsqr) wr = delta / sqrt(c2);
    }
    const double iz = 1.0 / pc[2];
    L.x = pc[0] * iz; L.y = pc[1] * iz; L.iz = iz; L.W = wr * w; L.r0 = -w * er.x * wr; L.r1 = -w * er.y * wr;
    const double t00 = -iz * K4[0], t02 = iz * iz * pc[0] * K4[0], t11 = -iz * K4[1], t12 = iz * iz * pc[1] * K4[1];
    const double t4[4] = {t00, t02, t11, t12};
    edge_jl(t4, T, L.Jl);
}
__device__ __forceinline__ void edge_terms(const EdgeLin& L, double* t9) {   // Jl^T W Jl (upper triangle) and Jl^T r
    const double a0 = L.Jl[0], a1 = L.Jl[1], a2 = L.Jl[2], c0 = L.Jl[3], c1 = L.Jl[4], c2 = L.Jl[5], W = L.W;
    t9[0] = (a0 * a0 + c0 * c0) * W; t9[1] = (a0 * a1 + c0 * c1) * W; t9[2] = (a0 * a2 + c0 * c2) * W;
    t9[3] = (a1 * a1 + c1 * c1) * W; t9[4] = (a1 * a2 + c1 * c2) * W; t9[5] = (a2 * a2 + c2 * c2) * W;
    t9[6] = a0 * L.r0 + c0 * L.r1; t9[7] = a1 * L.r0 + c1 * L.r1; t9[8] = a2 * L.r0 + c2 * L.r1;
}
// d9 = { D (6, symmetric), D bl (3) } -> the record of the edge
__device__ __forceinline__ void edge_record(const EdgeLin& L, const double* d9, double* yr, double* rw) {
    yr[0] = L.x; yr[1] = L.y; yr[2] = L.iz; yr[3] = L.free_pose ? L.W : 0.0;
    rw[0] = L.r0; rw[1] = L.r1;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        double x0 = L.W * L.Jl[3 * k], x1 = L.W * L.Jl[3 * k + 1], x2 = L.W * L.Jl[3 * k + 2];
        if (!L.free_pose) { x0 = 0; x1 = 0; x2 = 0; }     // no pose block: the edge is in no tuple, k_back skips it
        yr[4 + 3 * k] = x0; yr[5 + 3 * k] = x1; yr[6 + 3 * k] = x2;
        yr[10 + 3 * k] = x0 * d9[0] + x1 * d9[1] + x2 * d9[2];
        yr[11 + 3 * k] = x0 * d9[1] + x1 * d9[3] + x2 * d9[4];
        yr[12 + 3 * k] = x0 * d9[2] + x1 * d9[4] + x2 * d9[5];
        rw[2 + k] = x0 * d9[6] + x1 * d9[7] + x2 * d9[8];
    }
}

__global__ void __launch_bounds__(BA_TG, 7) k_land(BABatch A) {
    __shared__ __align__(16) double s_blk[BA_TG * BA_YRS];   // phase A: 9 terms per edge; phase C: record staging
    __shared__ __align__(16) double s_lm[BA_TG * 9];         // per landmark of the group: D (6), D bl (3)
    __shared__ double red[BA_TG / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b == 0 && tid == 0) { *A.pairs_counter = 0; *A.rs_flag = 0; }   // work queue of the k_pairs launch that follows; k_lin / k_build are done with the flag
    // the caller's stop flag lives in mapped host memory: ONE read per LM step crosses PCIe here, off the critical path; the LM decisions of
    // k_back (one per window, at the tail of the step) read the device copy
    if (b == 0 && tid == 32 && A.stop) *A.stop_dev = *(volatile int*)A.stop;
    const int4 gd = A.blkG_desc[b];
    const int p = gd.x, l0 = gd.y, e0 = gd.z, nl = gd.w & 255, ne = gd.w >> 8;
    // the edge ids do not depend on the problem's state: their lines are requested before the state is read (one level less in the chain
    // of dependent loads that a 128-edge CTA spends a third of its life in)
    if (tid < ne) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.e_pose + e0 + tid));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.e_cam + e0 + tid));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.e_pt + e0 + tid));
    }
    const BAState& S = A.state[p];
    if (S.done) return;
    const BAProb& P = A.prob[p];
    const int cur = S.cur;
    const double lambda = lambda_eff(S);
    const bool robust = S.round == 0 && A.delta > 0;
    const double delta = A.delta, dsqr = delta * delta;
    // the (pose, camera) table is gathered straight from global memory (one 128-byte line per edge, L1 / L2 hits): a per-CTA copy in
    // shared memory was measured slower (the kernel is bound by LSU wavefronts, and the copy adds 40 lines per 128 edges)
    const int ts = BA_TAB;
    const double* tb = A.tab[cur] + BA_TAB * (size_t)P.rt0;
    if (ne <= BA_TG) {
        // ---- phase A
        EdgeLin L;
        const int e = e0 + tid;
        if (tid < ne) {
            edge_lin(A, P, e, cur, robust, delta, dsqr, tb, ts, L);
            double t9[9];
            edge_terms(L, t9);
#pragma unroll
            for (int q = 0; q < 9; q++) s_blk[9 * tid + q] = t9[q];
        }
        __syncthreads();
        // ---- phase B
        if (tid < nl) {
            const int l = l0 + tid;
            double h[9];
#pragma unroll
            for (int q = 0; q < 9; q++) h[q] = 0;
            for (int i = A.pt_off[l] - e0; i < A.pt_off[l + 1] - e0; i++) {
#pragma unroll
                for (int q = 0; q < 9; q++) h[q] += s_blk[9 * i + q];
            }
            double* H = A.Hll + 6 * (size_t)l;
#pragma unroll
            for (int q = 0; q < 6; q++) H[q] = h[q];
            A.bl[3 * (size_t)l] = h[6]; A.bl[3 * (size_t)l + 1] = h[7]; A.bl[3 * (size_t)l + 2] = h[8];
            double d[6];
            landmark_dinv(h, lambda, d);
            double* o = s_lm + 9 * tid;
#pragma unroll
            for (int q = 0; q < 6; q++) o[q] = d[q];
            o[6] = d[0] * h[6] + d[1] * h[7] + d[2] * h[8]; o[7] = d[1] * h[6] + d[3] * h[7] + d[4] * h[8]; o[8] = d[2] * h[6] + d[4] * h[7] + d[5] * h[8];
        }
        __syncthreads();
        // ---- phase C
        double yr[BA_YR], rw[4];
        if (tid < ne) {
            edge_record(L, s_lm + 9 * (A.e_pt[e] - l0), yr, rw);
#pragma unroll
            for (int q = 0; q < BA_YR / 2; q++) reinterpret_cast<double2*>(s_blk + BA_YRS * tid)[q] = make_double2(yr[2 * q], yr[2 * q + 1]);
        }
        __syncthreads();
        for (int i = tid; i < (BA_YR / 2) * ne; i += BA_TG)
            reinterpret_cast<double2*>(A.yr + BA_YR * (size_t)e0)[i] = reinterpret_cast<const double2*>(s_blk + BA_YRS * (i >> 3))[i & 7];
        if (tid < ne) {
            double2* o = reinterpret_cast<double2*>(A.rw + 4 * (size_t)e);
            o[0] = make_double2(rw[0], rw[1]); o[1] = make_double2(rw[2], rw[3]);
        }
        return;
    }
    // ---- a single landmark with more than BA_TG edges (nl == 1)
    double h[9];
#pragma unroll
    for (int q = 0; q < 9; q++) h[q] = 0;
    for (int i = tid; i < ne; i += BA_TG) {
        EdgeLin L;
        edge_lin(A, P, e0 + i, cur, robust, delta, dsqr, tb, ts, L);
        double t9[9];
        edge_terms(L, t9);
#pragma unroll
        for (int q = 0; q < 9; q++) h[q] += t9[q];
    }
#pragma unroll
    for (int q = 0; q < 9; q++) h[q] = block_sum<BA_TG / 32>(h[q], red);
    if (tid == 0) {
        double* H = A.Hll + 6 * (size_t)l0;
#pragma unroll
        for (int q = 0; q < 6; q++) H[q] = h[q];
        A.bl[3 * (size_t)l0] = h[6]; A.bl[3 * (size_t)l0 + 1] = h[7]; A.bl[3 * (size_t)l0 + 2] = h[8];
    }
    double d9[9];
    landmark_dinv(h, lambda, d9);
    d9[6] = d9[0] * h[6] + d9[1] * h[7] + d9[2] * h[8]; d9[7] = d9[1] * h[6] + d9[3] * h[7] + d9[4] * h[8]; d9[8] = d9[2] * h[6] + d9[4] * h[7] + d9[5] * h[8];
    for (int i = tid; i < ne; i += BA_TG) {
        EdgeLin L;
        edge_lin(A, P, e0 + i, cur, robust, delta, dsqr, tb, ts, L);
        double yr[BA_YR], rw[4];
        edge_record(L, d9, yr, rw);
        double2* o = reinterpret_cast<double2*>(A.yr + BA_YR * (size_t)(e0 + i));
#pragma unroll
        for (int q = 0; q < BA_YR / 2; q++) o[q] = make_double2(yr[2 * q], yr[2 * q + 1]);
        double2* o2 = reinterpret_cast<double2*>(A.rw + 4 * (size_t)(e0 + i));
        o2[0] = make_double2(rw[0], rw[1]); o2[1] = make_double2(rw[2], rw[3]);
    }
}

// ------------------------------------------------------------------------------------------------ k_pairs
// warp per work item of a problem.
//   items [0, nChunksMax): a chunk of <= BA_CH (edge a, edge b) tuples of ONE pose pair (i, j) and ONE camera pair (ca, cb):
//     partial 6x6 of  sum Yt_a Bt_b^T = sum U_a (VD_a V_b^T) U_b^T  (the Schur product before the camera adjoints, which k_solve
//     applies once per camera pair).  Per tuple 80 bytes of yr[a] and 80 bytes of yr[b] are gathered; U_a and U_b are rebuilt from
//     x y w with the chunk's constants (intrinsics of ca / cb).  The tuples of a diagonal pair (i, i) are (e, e), one per edge of
//     pose i: there the 2x2 core is VD_e V_e^T - W_e I, i.e. the chunk sum is (Schur product) - Hpp_i in tJ space, so that the pose
//     block Hpp = sum tJ^T W tJ (block_solver.hpp:100-134 via constructQuadraticForm) needs no pass of its own.
//   items [nChunksMax, nChunksMax + K): free pose k, per camera c:  sum_e U_e r_e (bp) and sum_e U_e (VD_e bl) (Schur right-hand
//     side), both in tJ space
#define BA_STAGE_A (16 * 80)                     // edge a of 16 tuples: x y w W VD[6]
#define BA_STAGE_BYTES (BA_STAGE_A + 16 * 32)    // + edge b: x y w W  (diagonal pairs: r0 r1 (VD bl)0 (VD bl)1 of the same edge)
#define BA_NSTAGE 3
// Persistent warps: the grid holds as many CTAs as fit the device, every warp pulls items from a queue (which warp computes a chunk
// does not change its result).  While a chunk is being reduced, the record of the warp's next item is already in registers and the
// next chunk's tuple list is on its way to shared memory, so only the first item of a warp pays the start-up round trips.
// (A tensor-core form of the batch -- operand columns A_t = U'_a M^_t, B_t = U'_b exchanged through shared memory and summed with
// mma.sync.m8n8k4.f64, DMMA.884 -- was measured at the same chunk size: 14.1 instead of 13.4 ms per 15 launches.  The FP64 FMA pipe and the
// DMMA path have the same peak on B200, tools/microbench.cu: 62 / 64 FMA per clock and SM; the kernel is bound by LSU wavefronts -- one
// per gathered sector plus the shared-memory reads -- and the operand exchange adds three per tuple.)
__global__ void __launch_bounds__(128, 4) k_pairs(BABatch A, int n_items) {
    __shared__ __align__(16) unsigned char s_stage[4 * BA_NSTAGE * BA_STAGE_BYTES];   // per warp: BA_NSTAGE stages
    __shared__ int2 s_tup[4 * BA_CH];                                                  // per warp: the chunk's tuple list
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane & 1, slot = lane >> 1;
    unsigned char* stage0 = s_stage + (size_t)warp * BA_NSTAGE * BA_STAGE_BYTES;
    int2* tup = s_tup + warp * BA_CH;
    // piece g = k * 32 + lane of a batch (112 pieces of 16 bytes: 16 x 5 of edge a, then 16 x 2 of edge b); what does not depend on
    // the batch is fixed per lane here: tuple slot, which edge of the tuple, offset inside the 128-byte record.  Edge a brings three
    // sectors (x y w W | VD), edge b ONE (x y w W): V_b = W Jl_b is rebuilt from them and the rotation of (pose j, camera b), which
    // is a constant of the chunk -- four L2 sectors per tuple instead of six (the kernel is bound by the L2 gathers).
    int p_tl[4], p_src[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int g = k * 32 + lane, side = g >= 80;
        const int tl = side ? (g - 80) >> 1 : g / 5, part = side ? (g - 80) & 1 : g - 5 * (g / 5);
        p_tl[k] = g < 112 ? (tl | (side << 8)) : 0xffff;
        p_src[k] = side ? 2 * part : (part < 2 ? 2 * part : 6 + 2 * part);     // edge a: doubles 0..3 and 10..15 (VD); edge b: doubles 0..3
    }
    auto stage_tuples = [&](const int4& r) {                          // one cp.async group: the tuple list of chunk record r
        const int2* T = A.tuples + r.y;
        const unsigned tdst = (unsigned)__cvta_generic_to_shared(tup);
        for (int t = lane; t < r.z; t += 32) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tdst + 8u * t), "l"(T + t) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // dynamic item queue (the counter is zeroed by k_land, which precedes this kernel in every step)
    auto grab = [&]() { int v = 0; if (lane == 0) v = atomicAdd(A.pairs_counter, 1); return __shfl_sync(0xffffffffu, v, 0); };
    int it = grab();
    int4 rec0 = make_int4(0, 0, 0, 0), rec1 = rec0;
    if (it < n_items) { rec0 = A.item_rec[2 * (size_t)it]; rec1 = A.item_rec[2 * (size_t)it + 1]; }
    bool tup_ready = false;                                           // the tuple list of the current item is already (being) staged
    int nxt = it;
    for (; it < n_items; it = nxt) {
        nxt = grab();
        int4 nx0 = make_int4(0, 0, 0, 0), nx1 = nx0;
        if (nxt < n_items) { nx0 = A.item_rec[2 * (size_t)nxt]; nx1 = A.item_rec[2 * (size_t)nxt + 1]; }
        const int4 r0 = rec0, r1 = rec1;
        rec0 = nx0; rec1 = nx1;
        if (r0.z <= 0) continue;                                      // unused chunk slot
        // chunk item: a <= BA_CH-tuple chunk of one (pose pair, camera pair); the record holds {problem, first tuple, tuples, chunk}
        const int len = r0.z, ch = r0.w;
        const int done = A.state[r0.x].done;
        const bool diag = r1.z != 0;
        const double fxa = A.cam[BA_CAM_STRIDE * (size_t)r1.x], fya = A.cam[BA_CAM_STRIDE * (size_t)r1.x + 1];
        const double fxb = A.cam[BA_CAM_STRIDE * (size_t)r1.y], fyb = A.cam[BA_CAM_STRIDE * (size_t)r1.y + 1];
        // R(ext_cb * pose_j) at the accepted estimate: Jl_b = -1/z tmp R (types_six_dof_expmap.cpp:155-159)
        const double* Rb = A.tab[A.state[r0.x].cur] + BA_TAB * (size_t)r1.w;
        double Rm[9];
#pragma unroll
        for (int q = 0; q < 9; q++) Rm[q] = Rb[q];
        if (!tup_ready) stage_tuples(r0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        tup_ready = false;
        const bool next_chunk = nx0.z > 0;
        if (done) {
            if (next_chunk) { stage_tuples(nx0); tup_ready = true; }
            continue;
        }
        // the focal lengths of U = [fx u0 | fy u1] are folded into the 2x2 core; V_b = W_b Jl_b carries those of Jl_b itself
        const double f00 = fxa * fxb * fxb, f01 = fxa * fyb * fyb, f10 = fya * fxb * fxb, f11 = fya * fyb * fyb;
        // Batches of 16 tuples: the pieces of a batch (16 x 80 B per side) are copied global -> shared with cp.async (16 bytes per
        // request, whole sectors, no register write-back), BA_NSTAGE stages per warp.
        const int nbatch = (len + 15) >> 4;
        auto issue = [&](int bidx, int stg) {
            const int tl0 = bidx * 16;
            const unsigned dst0 = (unsigned)__cvta_generic_to_shared(stage0 + (size_t)stg * BA_STAGE_BYTES) + 16u * lane;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int tl = p_tl[k] & 0xff;
                if (p_tl[k] != 0xffff && tl0 + tl < len) {
                    const int2 ab = tup[tl0 + tl];
                    const double* src = A.yr + BA_YR * (size_t)((p_tl[k] >> 8) ? ab.y : ab.x) + p_src[k];
                    // diagonal pair (a == b): the second half carries {r, VD bl} of the same edge instead
                    if (diag && (p_tl[k] >> 8)) src = A.rw + 4 * (size_t)ab.x + p_src[k];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 512u * k), "l"(src) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        double acc[18];
#pragma unroll
        for (int i = 0; i < 18; i++) acc[i] = 0;
        double ub[3] = {0, 0, 0}, uu[3] = {0, 0, 0};      // diagonal pairs: rows 3h..3h+2 of sum U_e r_e (bp) and sum U_e (VD_e bl) (Schur rhs)
        // prefetch distance BA_NSTAGE - 1: every iteration commits exactly one (possibly empty) group
#pragma unroll
        for (int k = 0; k < BA_NSTAGE - 1; k++) {
            if (k < nbatch) issue(k, k); else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int bidx = 0; bidx < nbatch; bidx++) {
            if (bidx + BA_NSTAGE - 1 < nbatch) issue(bidx + BA_NSTAGE - 1, (bidx + BA_NSTAGE - 1) % BA_NSTAGE);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            // once the last batch has been issued nobody reads the tuple list any more: fetch the next chunk's (one more group in flight)
            if (next_chunk && !tup_ready && bidx + BA_NSTAGE - 1 >= nbatch - 1) { __syncwarp(); stage_tuples(nx0); tup_ready = true; }
            if (tup_ready) asm volatile("cp.async.wait_group %0;" ::"n"(BA_NSTAGE) : "memory");
            else asm volatile("cp.async.wait_group %0;" ::"n"(BA_NSTAGE - 1) : "memory");
            __syncwarp();
            // two lanes per tuple: lane parity h owns rows 3h..3h+2 of the 6x6 block
            if (bidx * 16 + slot < len) {
                const unsigned char* stg = stage0 + (size_t)(bidx % BA_NSTAGE) * BA_STAGE_BYTES;
                const double* ya = reinterpret_cast<const double*>(stg + 80 * slot);
                const double* eb = reinterpret_cast<const double*>(stg + BA_STAGE_A + 32 * slot);
                const double* ebx = diag ? ya : eb;               // x y w W of edge b
                // tJ_e^T = [fx u0 | fy u1] with  u0 = (xy, -(1 + x^2), y, -w, 0, xw),  u1 = (1 + y^2, -xy, -x, 0, -w, yw),  x = X/Z, y = Y/Z, w = 1/Z
                // (types_six_dof_expmap.cpp:136-153 with Z * (1/Z) = 1); the focal lengths are folded into M.
                double ua0[3], ua1[3];
                {
                    const double x = ya[0], y = ya[1], w = ya[2];
                    if (h == 0) {
                        const double xy = x * y;
                        ua0[0] = xy; ua0[1] = -fma(x, x, 1.0); ua0[2] = y;
                        ua1[0] = fma(y, y, 1.0); ua1[1] = -xy; ua1[2] = -x;
                    } else {
                        ua0[0] = -w; ua0[1] = 0; ua0[2] = x * w;
                        ua1[0] = 0; ua1[1] = -w; ua1[2] = y * w;
                    }
                }
                const double xb = ebx[0], yb = ebx[1], wb = ebx[2];
                const double sb = ebx[3] * wb;                     // V_b = W_b Jl_b = W_b w_b diag(f_b) [x_b R_2 - R_0 ; y_b R_2 - R_1]
                double m00 = 0, m01 = 0, m10 = 0, m11 = 0;        // M = diag(f_a) VD_a V_b^T diag(f_b) (2x2)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double v0 = fma(xb, Rm[6 + c], -Rm[c]), v1 = fma(yb, Rm[6 + c], -Rm[3 + c]), d0 = ya[4 + c], d1 = ya[7 + c];
                    m00 += d0 * v0; m01 += d0 * v1; m10 += d1 * v0; m11 += d1 * v1;
                }
                m00 *= sb * f00; m01 *= sb * f01; m10 *= sb * f10; m11 *= sb * f11;
                if (diag) {
                    const double Wf = ya[3];
                    m00 -= Wf * fxa * fxa; m11 -= Wf * fya * fya;   // U_e (VD_e V_e^T - W_e I) U_e^T: the pose block rides along
                    const double a0 = fxa * eb[0], a1 = fya * eb[1], c0 = fxa * eb[2], c1 = fya * eb[3];
#pragma unroll
                    for (int r = 0; r < 3; r++) { ub[r] += ua0[r] * a0 + ua1[r] * a1; uu[r] += ua0[r] * c0 + ua1[r] * c1; }
                }
                const double xyb = xb * yb, b01 = -fma(xb, xb, 1.0), b10 = fma(yb, yb, 1.0), xwb = xb * wb, ywb = yb * wb;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const double t0 = ua0[r] * m00 + ua1[r] * m10, t1 = ua0[r] * m01 + ua1[r] * m11;
                    double* a = acc + 6 * r;              // two chained FMAs per entry, the structural zeros of u0 / u1 skipped
                    a[0] = fma(t1, b10, fma(t0, xyb, a[0]));
                    a[1] = fma(t1, -xyb, fma(t0, b01, a[1]));
                    a[2] = fma(t1, -xb, fma(t0, yb, a[2]));
                    a[3] = fma(t0, -wb, a[3]);
                    a[4] = fma(t1, -wb, a[4]);
                    a[5] = fma(t1, ywb, fma(t0, xwb, a[5]));
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 18; i++) {
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        if (lane < 2) {
            double* out = A.partial + 36 * (size_t)ch + 18 * h;
#pragma unroll
            for (int i = 0; i < 18; i++) out[i] = acc[i];
        }
        if (diag) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
#pragma unroll
                for (int o = 16; o > 1; o >>= 1) { ub[i] += __shfl_xor_sync(0xffffffffu, ub[i], o); uu[i] += __shfl_xor_sync(0xffffffffu, uu[i], o); }
            }
            if (lane < 2) {
                double* out = A.prhs + 12 * (size_t)ch + 3 * h;
#pragma unroll
                for (int i = 0; i < 3; i++) { out[i] = ub[i]; out[6 + i] = uu[i]; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_solve
// Ends a round for problem p (called by one thread).  Returns nothing; sets done or arms the next round.
__device__ void round_over(const BABatch& A, BAState& S, bool stopped) {
    if (S.round == 0 && A.its2 >= 0 && !stopped) {
        S.round = 1; S.it = 0; S.mark = 1; S.round_start = 1; S.need_build = 1; S.lambda_pending = 1;
        *A.rs_flag = 1;                               // the next step's k_lin / k_build have work (k_land of this step has already cleared the flag)
    } else {
        S.done = 1;
    }
}

// Assembles, factorises (LDL^T) and solves the reduced camera system of one problem in `Hs` (packed lower triangle).  Inlined
// twice by k_solve -- once with the shared-memory matrix, once with a global-memory one -- so that each copy uses the
// loads / stores of its address space instead of generic ones.
__device__ __forceinline__ bool solve_reduced(const BABatch& A, const BAProb& P, const BAState& S, double* Hs, double* s_lcol, double* s_pan, int* s_ok_p,
                                              double lambda, double* s_wscr) {
    const int n = P.n, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // packed lower triangle, row i at i (i + 1) / 2: half the shared memory of a square matrix, so two CTAs (problems) share an SM
#define HS_AT(i, j) Hs[(size_t)(i) * ((i) + 1) / 2 + (j)]
    // ---- assemble the reduced camera system: diag blocks Hpp + lambda I, minus the Schur products.  The chunk partials hold
    //      N = sum tJ_a^T S tJ_b per (pose pair, camera pair); the block of the pair is sum over camera pairs of Adj_ca^T N Adj_cb
    for (int r = warp; r < n; r += BA_TS / 32) {
        for (int c = lane; c <= r; c += 32) {
            HS_AT(r, c) = r == c ? lambda : 0.0;     // Hpp arrives with the diagonal pairs' partials (k_pairs)
        }
    }
    __syncthreads();
    {
        // warp per pose pair: lanes own the 36 entries (lane, lane + 32) of N / tmp / the block; two 36-double scratch rows per warp
        const int nCh = S.nChunks;
        double* wN = s_wscr + 72 * warp;
        double* wT = wN + 36;
        for (int pid = warp; pid < P.nPairs; pid += BA_TS / 32) {
            if (A.pair_cnt[P.pair0 + pid] == 0) continue;
            double blk0 = 0, blk1 = 0;                     // entries lane and lane + 32
            // chunk ranges of all camera pairs first (lane q holds camera pair q), so that their reads are in flight together
            int my_nc = 0, my_first = 0;
            if (lane < P.CC) { my_nc = A.pc_nchunk[P.pc0 + pid * P.CC + lane]; my_first = A.pc_fchunk[P.pc0 + pid * P.CC + lane]; }
            // four camera pairs at a time: the first chunk partial of each is requested before any of them is used (one L2 round trip
            // per group instead of one per camera pair; a list longer than one chunk continues its chain below)
            for (int c0 = 0; c0 < P.CC; c0 += 4) {
                int ncs[4], fs[4];
                double a0[4], a1[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int combo = min(c0 + q, P.CC - 1);
                    ncs[q] = c0 + q < P.CC ? __shfl_sync(0xffffffffu, my_nc, combo) : 0;
                    fs[q] = __shfl_sync(0xffffffffu, my_first, combo);
                    a0[q] = 0; a1[q] = 0;
                    if (ncs[q] > 0 && fs[q] < nCh) {
                        const double* pp = A.partial + 36 * (size_t)(P.chunk0 + fs[q]);
                        a0[q] = pp[lane];
                        if (lane < 4) a1[q] = pp[32 + lane];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int nc = ncs[q], first = fs[q], combo = c0 + q;
                    if (nc == 0) continue;
                    double n0 = a0[q], n1 = a1[q];
                    for (int c = 1; c < nc && first + c < nCh; c++) {
                        const double* pp = A.partial + 36 * (size_t)(P.chunk0 + first + c);
                        n0 += pp[lane];
                        if (lane < 4) n1 += pp[32 + lane];
                    }
                    __syncwarp();
                    wN[lane] = n0;
                    if (lane < 4) wN[32 + lane] = n1;
                    __syncwarp();
                    const double* Aa = A.cam + BA_CAM_STRIDE * (size_t)(P.c0 + combo / P.nC) + BA_CAM_ADJ;
                    const double* Ab = A.cam + BA_CAM_STRIDE * (size_t)(P.c0 + combo % P.nC) + BA_CAM_ADJ;
                    for (int en = lane; en < 36; en += 32) {   // tmp = N Adj_cb
                        const int a6 = en / 6, j6 = en - 6 * a6;
                        double t = 0;
#pragma unroll
                        for (int k6 = 0; k6 < 6; k6++) t += wN[a6 * 6 + k6] * Ab[k6 * 6 + j6];
                        wT[en] = t;
                    }
                    __syncwarp();
                    for (int en = lane; en < 36; en += 32) {   // block += Adj_ca^T tmp
                        const int i6 = en / 6, j6 = en - 6 * i6;
                        double t = 0;
#pragma unroll
                        for (int k6 = 0; k6 < 6; k6++) t += Aa[k6 * 6 + i6] * wT[k6 * 6 + j6];
                        if (en < 32) blk0 += t; else blk1 += t;
                    }
                }
            }
            int i1, i2;
            pair_decode(pid, P.K, i1, i2);
            for (int en = lane; en < 36; en += 32) {
                const int r = en / 6, c = en - 6 * r;
                const double v = en < 32 ? blk0 : blk1;
                if (i1 != i2) HS_AT(6 * i2 + c, 6 * i1 + r) = -v;          // i1 < i2: the block lies above the diagonal, its transpose below
                else if (c <= r) HS_AT(6 * i1 + r, 6 * i1 + c) -= v;
            }
        }
    }
    // bp_k = sum_c Adj_c^T b_(k,c) and the Schur right-hand side bs_k = bp_k - sum_c Adj_c^T u_(k,c); b_(k,c) and u_(k,c) are the sums of
    // the right-hand-side partials of the chunks of the diagonal pair (k, k), camera pair (c, c)
    for (int i = tid; i < n; i += BA_TS) {
        const int k = i / 6, c6 = i - 6 * k;
        const int pidd = k * P.K - k * (k - 1) / 2;
        double add = 0, sub = 0;
        for (int cl = 0; cl < P.nC; cl++) {
            const double* Ad = A.cam + BA_CAM_STRIDE * (size_t)(P.c0 + cl) + BA_CAM_ADJ;
            const int pc = P.pc0 + pidd * P.CC + cl * P.nC + cl;
            const int first = A.pc_fchunk[pc], nc = A.pc_nchunk[pc];
            for (int a = 0; a < 6; a++) {
                double ub = 0, uu = 0;
                for (int c = 0; c < nc && first + c < S.nChunks; c++) {
                    const double* pr = A.prhs + 12 * (size_t)(P.chunk0 + first + c);
                    ub += pr[a]; uu += pr[6 + a];
                }
                add += Ad[a * 6 + c6] * ub; sub += Ad[a * 6 + c6] * uu;
            }
        }
        A.bp[6 * (size_t)P.k0 + i] = add;
        A.bs[6 * (size_t)P.k0 + i] = add - sub;
    }
    if (tid == 0) *s_ok_p = 1;
    __syncthreads();
    // ---- LDL^T in place (lower triangle: L below the diagonal, D on it), right-looking over block columns of 6 (one pose): three
    //      barriers per pose instead of two per scalar column.
    //        1. diagonal block: scalar LDL^T of the 6x6 by one warp
    //        2. panel: thread per row r below the block: W_r = L_r D from W_r L_jj^T = A_r (forward substitution), L_r = W_r D^-1
    //        3. trailing update: H(r, k) -= W_r . L_k for the rows / columns below the block (warp per row, lanes along the row)
    for (int j0 = 0; j0 < n; j0 += 6) {
        if (warp == 0) {
            for (int c = 0; c < 6; c++) {
                const double dc = HS_AT(j0 + c, j0 + c);
                if (dc == 0.0 || !isfinite(dc)) { if (lane == 0) *s_ok_p = 0; break; }
                if (lane > c && lane < 6) HS_AT(j0 + lane, j0 + c) /= dc;
                __syncwarp();
                if (lane < 36) {                 // H(r, k) -= L_rc d_c L_kc for c < k <= r < 6
                    const int r = lane / 6, k = lane - 6 * r;
                    if (k > c && k <= r) HS_AT(j0 + r, j0 + k) -= HS_AT(j0 + r, j0 + c) * dc * HS_AT(j0 + k, j0 + c);
                }
                if (lane == 0 && c < 5) {        // lanes 32..35 of the 36 entries: (5, 2) .. (5, 5)
                    for (int k = 2; k < 6; k++) if (k > c) HS_AT(j0 + 5, j0 + k) -= HS_AT(j0 + 5, j0 + c) * dc * HS_AT(j0 + k, j0 + c);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (*s_ok_p == 0) break;
        {
            double Lj[15], dj[6];                // L_jj (strict lower, row-major) and d of the block: the same for every row
#pragma unroll
            for (int c = 0, q = 0; c < 6; c++) {
                dj[c] = HS_AT(j0 + c, j0 + c);
#pragma unroll
                for (int k = 0; k < c; k++) Lj[q++] = HS_AT(j0 + c, j0 + k);
            }
            for (int r = j0 + 6 + tid; r < n; r += BA_TS) {
                double* row = &HS_AT(r, j0);
                double w[6];
#pragma unroll
                for (int c = 0, q = 0; c < 6; c++) {
                    double a = row[c];
#pragma unroll
                    for (int k = 0; k < c; k++) a -= w[k] * Lj[q++];
                    w[c] = a;
                }
#pragma unroll
                for (int c = 0; c < 6; c++) { row[c] = w[c] / dj[c]; s_pan[6 * r + c] = w[c]; }
            }
        }
        __syncthreads();
        for (int r = j0 + 6 + warp; r < n; r += BA_TS / 32) {
            const double w0 = s_pan[6 * r], w1 = s_pan[6 * r + 1], w2 = s_pan[6 * r + 2], w3 = s_pan[6 * r + 3], w4 = s_pan[6 * r + 4], w5 = s_pan[6 * r + 5];
            double* row = &HS_AT(r, 0);
            for (int k = j0 + 6 + lane; k <= r; k += 32) {
                const double* Lk = &HS_AT(k, j0);
                row[k] -= w0 * Lk[0] + w1 * Lk[1] + w2 * Lk[2] + w3 * Lk[3] + w4 * Lk[4] + w5 * Lk[5];
            }
        }
        __syncthreads();
    }
    __syncthreads();
    const bool ok = *s_ok_p != 0;
    double* x = A.xp + 6 * (size_t)P.k0;
    if (ok && warp == 0) {
        const double* bs = A.bs + 6 * (size_t)P.k0;
        double* xs = s_lcol;                   // the solve runs on the shared copy
        for (int i = lane; i < n; i += 32) xs[i] = bs[i];
        __syncwarp();
        // forward: L y = b, block column by block column (the 6 unknowns of a block are solved redundantly by every lane)
        for (int j0 = 0; j0 < n; j0 += 6) {
            double y[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double a = xs[j0 + c];
#pragma unroll
                for (int k = 0; k < c; k++) a -= HS_AT(j0 + c, j0 + k) * y[k];
                y[c] = a;
            }
            __syncwarp();
            if (lane < 6) xs[j0 + lane] = y[lane == 0 ? 0 : lane == 1 ? 1 : lane == 2 ? 2 : lane == 3 ? 3 : lane == 4 ? 4 : 5];
            for (int i = j0 + 6 + lane; i < n; i += 32) {
                const double* Li = &HS_AT(i, j0);
                xs[i] -= Li[0] * y[0] + Li[1] * y[1] + Li[2] * y[2] + Li[3] * y[3] + Li[4] * y[4] + Li[5] * y[5];
            }
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) xs[i] /= HS_AT(i, i);
        __syncwarp();
        // backward: L^T x = y
        for (int j0 = n - 6; j0 >= 0; j0 -= 6) {
            double y[6];
#pragma unroll
            for (int c = 5; c >= 0; c--) {
                double a = xs[j0 + c];
#pragma unroll
                for (int k = c + 1; k < 6; k++) a -= HS_AT(j0 + k, j0 + c) * y[k];
                y[c] = a;
            }
            __syncwarp();
            if (lane < 6) xs[j0 + lane] = y[lane == 0 ? 0 : lane == 1 ? 1 : lane == 2 ? 2 : lane == 3 ? 3 : lane == 4 ? 4 : 5];
            for (int i = lane; i < j0; i += 32) {
                double a = xs[i];
#pragma unroll
                for (int c = 0; c < 6; c++) a -= HS_AT(j0 + c, i) * y[c];
                xs[i] = a;
            }
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) x[i] = xs[i];
    }
    return ok;
}

__global__ void __launch_bounds__(BA_TS, 2) k_solve(BABatch A, int hs_smem_doubles, int max_n, int hs_smem_n) {
    extern __shared__ double sm_hs[];
    __shared__ double red[BA_TS / 32];
    __shared__ double s_wscr[(BA_TS / 32) * 72];   // per warp: N and N Adj of the pair being assembled
    __shared__ int s_go, s_ok;
    double* s_lcol = sm_hs + hs_smem_doubles;   // n doubles after the matrix
    const int p = blockIdx.x, tid = threadIdx.x;
    BAState& S = A.state[p];
    if (S.done) return;
    const BAProb& P = A.prob[p];
    const int n = P.n;
    // ---- start of a round / of an iteration (thread 0)
    if (tid == 0) {
        int go = 1;
        if (S.need_build) {
            if (S.round_start) {
                double chi0 = 0, act = 0;
                for (int b = 0; b < P.nbE; b++) { chi0 += A.partE[2 * (size_t)(P.blkE0 + b)]; act += A.partE[2 * (size_t)(P.blkE0 + b) + 1]; }
                if (S.round == 0) S.initial_chi2 = chi0;
                const int its = S.round == 0 ? A.its1 : A.its2;
                if (act == 0.0 || S.it >= its) {   // SparseOptimizer::optimize returns at once (sparse_optimizer.cpp:356-359) / optimize(0)
                    if (act > 0.0) { S.currentChi = chi0; S.last = S.cur; }
                    S.mark = 0; S.round_start = 0;
                    round_over(A, S, false);
                    S.skip = 1;
                    go = 0;
                } else {
                    S.currentChi = chi0;
                    S.last = S.cur;
                    S.lambda = 1e-5 * __longlong_as_double((long long)S.maxdiag_bits);   // computeLambdaInit
                    S.ni = 2; S.nBad = 0;
                    S.lambda_pending = 0;
                }
            }
            if (go) { S.iniChi = S.currentChi; S.qmax = 0; S.rho = 0; }
        }
        s_go = go;
    }
    __syncthreads();
    if (!s_go) return;
    const double lambda = S.lambda;
    bool ok;
    // s_lcol: n doubles (any n); the panel W [n][6] follows it for matrices in shared memory, else it lies behind the packed triangle in global memory
    if (n <= hs_smem_n) ok = solve_reduced(A, P, S, sm_hs, s_lcol, s_lcol + max_n, &s_ok, lambda, s_wscr);
    else ok = solve_reduced(A, P, S, A.Hs + P.hs_off, s_lcol, A.Hs + P.hs_off + (((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1), &s_ok, lambda, s_wscr);
    double* x = A.xp + 6 * (size_t)P.k0;
    if (!ok) for (int i = tid; i < n; i += BA_TS) x[i] = 0.0;   // g2o applies a stale x; the trial is rejected either way
    __syncthreads();
    // y_(k,c) = Adj_c x_k : what the back-substitution needs of the pose increments (Jp x = tJ (Adj_c x))
    for (int idx = tid; idx < P.K * P.nC * 6; idx += BA_TS) {
        const int i6 = idx % 6, kc = idx / 6, cl = kc % P.nC, k = kc / P.nC;
        const double* Ad = A.cam + BA_CAM_STRIDE * (size_t)(P.c0 + cl) + BA_CAM_ADJ + 6 * i6;
        const double* xk = x + 6 * k;
        A.ut_y[6 * (size_t)(P.ut0 + kc) + i6] = Ad[0] * xk[0] + Ad[1] * xk[1] + Ad[2] * xk[2] + Ad[3] * xk[3] + Ad[4] * xk[4] + Ad[5] * xk[5];
    }
    // ---- trial poses: exp(x) * pose for free poses, copy for fixed ones;  pose part of computeScale()
    const int cur = S.cur;
    for (int i = tid; i < P.nP; i += BA_TS) {
        const int kg = A.pose_free[P.p0 + i];
        const double* src = A.pose[cur] + 7 * (size_t)(P.p0 + i);
        double* dst = A.pose[cur ^ 1] + 7 * (size_t)(P.p0 + i);
        double np7[7];
        if (kg >= 0) se3_oplus(A.xp + 6 * (size_t)kg, src, np7);
        else for (int q = 0; q < 7; q++) np7[q] = src[q];
        for (int q = 0; q < 7; q++) dst[q] = np7[q];
        for (int cl = 0; cl < P.nC; cl++) {           // the trial estimate's projection table
            double o[BA_TAB];
            tab_entry(A.cam + BA_CAM_STRIDE * (size_t)(P.c0 + cl), np7, o);
            double2* d = reinterpret_cast<double2*>(A.tab[cur ^ 1] + BA_TAB * (size_t)(P.rt0 + i * P.nC + cl));
#pragma unroll
            for (int k = 0; k < BA_TAB / 2; k++) d[k] = make_double2(o[2 * k], o[2 * k + 1]);
        }
    }
    double sc = 0;
    for (int j = tid; j < n; j += BA_TS) { const double xj = x[j]; sc += xj * (lambda * xj + A.bp[6 * (size_t)P.k0 + j]); }
    sc = block_sum<BA_TS / 32>(sc, red);
    if (tid == 0) { S.poseScale = sc; S.solve_ok = ok ? 1 : 0; }
}

// ------------------------------------------------------------------------------------------------ k_back
// LM decision of one problem after a trial (optimization_algorithm_levenberg.cpp:86-164); one thread.  Returns true when the trial
// estimate was accepted (S.cur flipped).
__device__ bool lm_decide(const BABatch& A, const BAProb& P, BAState& S) {
    double tempChi = 0, scale = 0;
    const volatile double* PL = A.partL;
    for (int q = 0; q < P.nbL; q++) { tempChi += PL[2 * (size_t)(P.blkL0 + q)]; scale += PL[2 * (size_t)(P.blkL0 + q) + 1]; }
    if (!S.solve_ok) tempChi = 1.7976931348623157e308;
    scale = (S.poseScale + scale) + 1e-3;
    const double rho = (S.currentChi - tempChi) / scale;
    S.rho = rho;
    S.trials++;
    bool accepted;
    if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow(2 * rho - 1, 3.0);
        alpha = fmin(alpha, 2. / 3.);
        S.lambda *= fmax(1. / 3., alpha);
        S.ni = 2;
        S.currentChi = tempChi;
        accepted = true;
    } else {
        S.lambda *= S.ni;
        S.ni *= 2;
        accepted = false;
    }
    S.qmax++;
    S.last = S.cur ^ 1;
    if (accepted) S.cur ^= 1;           // pop() is a no-op: the rejected estimate stays in the other buffer
    const bool stopped = *A.stop_dev != 0;
    const bool again = rho < 0 && S.qmax < 10 && !stopped;
    S.mark = 0; S.round_start = 0;
    if (again) { S.need_build = 0; return accepted; }
    S.iterations++;
    int res = 0;
    if (S.qmax == 10 || rho == 0) res = 1;
    else {
        if ((S.iniChi - S.currentChi) * 1e3 < S.iniChi) S.nBad++; else S.nBad = 0;
        if (S.nBad >= 3) res = 1;
    }
    S.it++;
    S.need_build = 1;
    const int its = S.round == 0 ? A.its1 : A.its2;
    if (stopped) S.stopped = 1;
    if (res || S.it >= its || stopped) round_over(A, S, stopped);
    return accepted;
}

// thread per landmark (a thread-per-edge form over the landmark groups of k_land, with the records and the projection table staged in
// shared memory, was measured 45 % slower: these kernels are bound by LSU wavefronts, not by the strided loads):
//   x_l = D (bl - sum_e B_e^T x_p),  B_e^T x_p = V_e^T tJ_e (Adj_c x_p)  from the first 80 bytes of the edge records (block_solver.hpp:461-481),
//   trial point, then the trial error and robust chi2 of every edge of the landmark through the trial estimate's projection table;
// the last CTA of the problem takes the LM decision.
__global__ void __launch_bounds__(BA_TL, 8) k_back(BABatch A) {
    __shared__ double red[BA_TL / 32];
    __shared__ int s_last;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int p = A.blkL_prob[b];
    const int l = A.blkL_l0[b] + tid;
    // the landmark's edge range does not depend on the problem's state: requested before the state is read
    asm volatile("prefetch.global.L1 [%0];" ::"l"(A.pt_off + l));
    BAState& S = A.state[p];
    if (S.done && !S.skip) return;
    const BAProb& P = A.prob[p];
    const bool skip = S.skip != 0;
    double chi = 0, sc = 0;
    if (!skip && l < P.l0 + P.nL) {
        const int cur = S.cur;
        const double lambda = S.lambda;
        const bool robust = S.round == 0 && A.delta > 0;
        const double delta = A.delta, dsqr = delta * delta;
        double x0 = 0, x1 = 0, x2 = 0;
        const double* bl = A.bl + 3 * (size_t)l;
        const int e_begin = A.pt_off[l], e_end = A.pt_off[l + 1];
        const double* __restrict__ tb = A.tab[cur ^ 1] + BA_TAB * (size_t)P.rt0;
        const int* __restrict__ ecam_ = A.e_cam;
        const int* __restrict__ epose_ = A.e_pose;
        if (S.solve_ok) {
            double c0 = bl[0], c1 = bl[1], c2 = bl[2];
#pragma unroll 2
            for (int e = e_begin; e < e_end; e++) {
                const int kg = A.e_kf[e];
                if (kg < 0) continue;
                const int cg = ecam_[e];
                const double* c = A.cam + BA_CAM_STRIDE * (size_t)cg;
                double r[10], y[6];
                const double2* rp = reinterpret_cast<const double2*>(A.yr + BA_YR * (size_t)e);
#pragma unroll
                for (int q = 0; q < 5; q++) { const double2 v = rp[q]; r[2 * q] = v.x; r[2 * q + 1] = v.y; }
                load6(A.ut_y + 6 * (size_t)(P.ut0 + (kg - P.k0) * P.nC + (cg - P.c0)), y);
                // tJ_e y with tJ_e^T = [fx u0 | fy u1],  u0 = (xy, -(1 + x^2), y, -w, 0, xw),  u1 = (1 + y^2, -xy, -x, 0, -w, yw)
                const double x = r[0], yy = r[1], w = r[2], xy = x * yy;
                const double s0 = c[0] * (xy * y[0] - fma(x, x, 1.0) * y[1] + yy * y[2] - w * y[3] + x * w * y[5]);
                const double s1 = c[1] * (fma(yy, yy, 1.0) * y[0] - xy * y[1] - x * y[2] - w * y[4] + yy * w * y[5]);
                c0 -= r[4] * s0 + r[7] * s1; c1 -= r[5] * s0 + r[8] * s1; c2 -= r[6] * s0 + r[9] * s1;
            }
            double d[6];
            landmark_dinv(A.Hll + 6 * (size_t)l, lambda, d);
            x0 = d[0] * c0 + d[1] * c1 + d[2] * c2; x1 = d[1] * c0 + d[3] * c1 + d[4] * c2; x2 = d[2] * c0 + d[4] * c1 + d[5] * c2;
        }
        const double* po = A.pt[cur] + 3 * (size_t)l;
        const double pn[3] = {po[0] + x0, po[1] + x1, po[2] + x2};
        double* pw = A.pt[cur ^ 1] + 3 * (size_t)l;
        pw[0] = pn[0]; pw[1] = pn[1]; pw[2] = pn[2];
        sc = x0 * (lambda * x0 + bl[0]) + x1 * (lambda * x1 + bl[1]) + x2 * (lambda * x2 + bl[2]);
        // software pipeline: the ids / observation / weight of edge e + 1 are fetched before edge e is evaluated (the error store of
        // edge e would otherwise fence the loads of the next iteration: the pointers of BABatch are not restrict-qualified)
        const unsigned char* __restrict__ lvl_ = A.level;
        const double2* __restrict__ obs_ = reinterpret_cast<const double2*>(A.e_obs);
        const double* __restrict__ info_ = A.e_info;
        double2* __restrict__ errw_ = reinterpret_cast<double2*>(A.err[cur ^ 1]);
        int n_lv = 0, n_cam = 0, n_pose = 0;
        double2 n_obs = make_double2(0.0, 0.0);
        double n_info = 0;
        if (e_begin < e_end) { n_lv = lvl_[e_begin]; n_cam = ecam_[e_begin]; n_pose = epose_[e_begin]; n_obs = obs_[e_begin]; n_info = info_[e_begin]; }
        for (int e = e_begin; e < e_end; e++) {
            const int lv = n_lv, cg = n_cam, pg = n_pose;
            const double2 ob = n_obs;
            const double w = n_info;
            if (e + 1 < e_end) { n_lv = lvl_[e + 1]; n_cam = ecam_[e + 1]; n_pose = epose_[e + 1]; n_obs = obs_[e + 1]; n_info = info_[e + 1]; }
            if (lv) continue;
            double T[12], pc[3], er[2];
            load_tab(tb + BA_TAB * ((pg - P.p0) * P.nC + (cg - P.c0)), T);
            tab_project(T, pn, pc);
            tab_error(pc, A.cam + BA_CAM_STRIDE * (size_t)cg, ob.x, ob.y, er);
            errw_[e] = make_double2(er[0], er[1]);
            const double c2 = (er[0] * er[0] + er[1] * er[1]) * w;
            chi += robust ? huber_rho0(c2, delta, dsqr) : c2;
        }
    }
    const double cs = block_sum<BA_TL / 32>(chi, red);
    const double ss = block_sum<BA_TL / 32>(sc, red);
    if (tid == 0) {
        A.partL[2 * (size_t)b] = cs; A.partL[2 * (size_t)b + 1] = ss;
        __threadfence();
        const unsigned t = atomicAdd(&S.ticket, 1u);
        s_last = (t == (unsigned)P.nbL - 1);
    }
    __syncthreads();
    if (!s_last || tid != 0) return;
    // ---- the last CTA of the problem: LM decision (tab[cur ^ 1] already belongs to the trial estimate, so an accepted trial needs nothing more)
    __threadfence();
    S.ticket = 0;
    if (skip) { S.skip = 0; return; }
    lm_decide(A, P, S);
}

// ------------------------------------------------------------------------------------------------ final
__global__ void __launch_bounds__(BA_TE) k_final(BABatch A) {
    __shared__ double red[BA_TE / 32];
    const int b = blockIdx.x;
    const int p = A.blkE_prob[b];
    const BAState& S = A.state[p];
    const BAProb& P = A.prob[p];
    const int lb = b - P.blkE0, tid = threadIdx.x;
    const int cur = S.cur;
    const int e = P.e0 + lb * BA_TE + tid;
    double nout = 0;
    if (e < P.e0 + P.nE) {
        const double l0 = A.err[S.last][2 * e], l1 = A.err[S.last][2 * e + 1];
        double pc[3];
        tab_project(A.tab[cur] + BA_TAB * (size_t)(P.rt0 + (A.e_pose[e] - P.p0) * P.nC + (A.e_cam[e] - P.c0)), A.pt[cur] + 3 * (size_t)A.e_pt[e], pc);
        const bool out = (l0 * l0 + l1 * l1) * A.e_info[e] > A.chi2_th || !(pc[2] > 0.0);
        A.outlier[e] = out;
        nout = out;
    }
    nout = block_sum<BA_TE / 32>(nout, red);
    if (tid == 0) A.partE[2 * (size_t)b] = nout;
    for (int i = lb * BA_TE + tid; i < P.nP; i += P.nbE * BA_TE) {
        double R[9];
        const double* s = A.pose[cur] + 7 * (size_t)(P.p0 + i);
        q_to_matrix(s, R);
        double* o = A.poses_out + 12 * (size_t)(P.p0 + i);
        for (int r = 0; r < 3; r++) { o[r * 4] = R[r * 3]; o[r * 4 + 1] = R[r * 3 + 1]; o[r * 4 + 2] = R[r * 3 + 2]; o[r * 4 + 3] = s[4 + r]; }
    }
    for (int i = lb * BA_TE + tid; i < 3 * P.nL; i += P.nbE * BA_TE) A.points_out[3 * (size_t)P.l0 + i] = A.pt[cur][3 * (size_t)P.l0 + i];
}
__global__ void k_stats(BABatch A) {   // one CTA
    __shared__ int s_active;
    if (threadIdx.x == 0) s_active = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < A.nProb; p += blockDim.x) {
        const BAState& S = A.state[p];
        const BAProb& P = A.prob[p];
        double tot = 0;
        if (P.nE > 0) for (int b = 0; b < P.nbE; b++) tot += A.partE[2 * (size_t)(P.blkE0 + b)];
        orbba_stats_t st;
        st.initial_chi2 = S.initial_chi2; st.final_chi2 = S.currentChi; st.final_lambda = S.lambda;
        st.iterations = S.iterations; st.trials = S.trials; st.outliers = (int)tot;
        st.status = S.aborted ? ORB_E_ABORTED : ORB_OK;
        A.stats[p] = st;
        if (!S.done) atomicAdd(&s_active, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) *A.n_active = s_active;   // mapped host memory, read by finish()
}

// ------------------------------------------------------------------------------------------------ compact input -> device arrays
// orbba_upload_f32: the staged data are the caller's 16-byte edge records, CV_32F points and the per-level weights (what the
// reference holds: cv::KeyPoint::pt is float, MapPoint::GetWorldPos() CV_32F, mvInvLevelSigma2 float); the FP64 arrays of the kernels
// are produced here, so the host -> device copy carries 16 instead of 36 bytes per edge.
struct BACompact {
    const orbba_edge16_t* e16;        // [Etot] stored (landmark-grouped) order
    const float* pts32;               // [Ltot][3]
    const float* isig;                // per problem: n_levels weights at isig0[p]
    const int* isig0;
    int *e_pose, *e_pt, *e_cam;
    double *e_obs, *e_info, *pt0;
};
__global__ void k_expand(BABatch A, BACompact C) {
    const int b = blockIdx.x;
    const int p = A.blkE_prob[b];
    const BAProb P = A.prob[p];
    const int lb = b - P.blkE0, tid = threadIdx.x;
    const int e = P.e0 + lb * BA_TE + tid;
    if (e < P.e0 + P.nE) {
        const orbba_edge16_t r = C.e16[e];
        C.e_pose[e] = P.p0 + r.pose; C.e_pt[e] = P.l0 + (int)r.point; C.e_cam[e] = P.c0 + r.cam;
        C.e_obs[2 * (size_t)e] = (double)r.u; C.e_obs[2 * (size_t)e + 1] = (double)r.v;
        C.e_info[e] = (double)C.isig[C.isig0[p] + r.octave];
    }
    for (int i = lb * BA_TE + tid; i < 3 * P.nL; i += P.nbE * BA_TE) C.pt0[3 * (size_t)P.l0 + i] = (double)C.pts32[3 * (size_t)P.l0 + i];
}

// ================================================================================================ host side
// One view over the two input forms of orbba_upload / orbba_upload_f32
struct ProbView {
    int nP = 0, nL = 0, nE = 0, nC = 0, nLev = 0;
    const uint8_t* pose_fixed = nullptr;
    const double *poses = nullptr, *points = nullptr, *obs = nullptr, *info = nullptr, *cam_K = nullptr, *cam_ext = nullptr, *cam_adj = nullptr;
    const int32_t *ep = nullptr, *el = nullptr, *ec = nullptr;
    const float *poses32 = nullptr, *points32 = nullptr, *isig = nullptr;
    const orbba_edge16_t* e16 = nullptr;
    bool compact = false;
    int pose(int e) const { return compact ? (int)e16[e].pose : ep[e]; }
    int point(int e) const { return compact ? (int)e16[e].point : el[e]; }
    int cam(int e) const { return compact ? (int)e16[e].cam : ec[e]; }
    double pose_el(int i, int k) const { return compact ? (double)poses32[12 * i + k] : poses[12 * i + k]; }
    bool has_nulls() const {
        if ((nP && !pose_fixed) || !cam_K || !cam_ext || !cam_adj) return true;
        if (compact) return (nP && !poses32) || (nL && !points32) || (nE && (!e16 || !isig));
        return (nP && !poses) || (nL && !points) || (nE && (!ep || !el || !ec || !obs || !info));
    }
};
static ProbView view_of(const orbba_problem_t& Q) {
    ProbView V;
    V.nP = Q.n_poses; V.nL = Q.n_points; V.nE = Q.n_edges; V.nC = Q.n_cams;
    V.pose_fixed = Q.pose_fixed; V.poses = Q.poses; V.points = Q.points; V.obs = Q.edge_obs; V.info = Q.edge_inv_sigma2;
    V.cam_K = Q.cam_K; V.cam_ext = Q.cam_ext; V.cam_adj = Q.cam_adj; V.ep = Q.edge_pose; V.el = Q.edge_point; V.ec = Q.edge_cam;
    return V;
}
static ProbView view_of(const orbba_problem_f32_t& Q) {
    ProbView V;
    V.compact = true;
    V.nP = Q.n_poses; V.nL = Q.n_points; V.nE = Q.n_edges; V.nC = Q.n_cams; V.nLev = Q.n_levels;
    V.pose_fixed = Q.pose_fixed; V.poses32 = Q.poses; V.points32 = Q.points; V.e16 = Q.edges; V.isig = Q.inv_sigma2;
    V.cam_K = Q.cam_K; V.cam_ext = Q.cam_ext; V.cam_adj = Q.cam_adj;
    return V;
}

struct orbba {
    int device = 0, max_problems = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // optional: uploads (host->device + index kernels) go here, runs wait for them
    cudaEvent_t up_ev = nullptr, done_ev = nullptr;   // upload finished / the enqueued LM steps finished
    bool up_pending = false;
    cudaStream_t dl_stream = nullptr;        // device->host result copies (do not queue behind later work on the compute stream)
    int n = 0;
    BABatch A;                               // device pointers of the uploaded batch
    std::vector<BAProb> probs;
    std::vector<std::vector<int>> perm;      // per problem: caller edge index of every stored edge (empty = identity)
    uint8_t* d_arena = nullptr; size_t arena_cap = 0;
    uint8_t* h_stage = nullptr; size_t stage_cap = 0;   // pinned staging of the static arrays
    int *d_blkP_prob = nullptr, *d_blkP_first = nullptr, *d_blkI_first = nullptr, *d_pose_prob = nullptr;
    int nbE = 0, nbL = 0, nbP = 0, nbI = 0, nbG = 0, Ktot = 0, max_n = 0;
    int hs_smem_n = 0;                 // reduced systems up to this size are factorised in shared memory (0: a very large problem in the batch took the space)
    int pairs_grid = 148 * 4;          // resident CTAs of the persistent k_pairs (set from the occupancy calculator at create)
    long long Etot = 0, Ltot = 0, Ptot = 0;
    int* h_flags = nullptr;                  // pinned + mapped: [0] stop flag, [1] active problems after the last step
    int* d_flags = nullptr;
    bool pending = false;                    // steps were enqueued and not yet checked for completion
    int its1 = 0, its2 = 0;
    long long launches = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool profile = false;
    double prof_ms = 0; int prof_calls = 0; bool prof_pending = false;
    std::vector<cudaEvent_t> kev;            // per-kernel timing: BA_KEV_STEPS x 7 events
    int kev_steps = 0;
};
#define BA_KEV_STEPS 96

static void orbba_free(orbba* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    cudaFree(b->d_arena);
    if (b->h_stage) cudaFreeHost(b->h_stage);
    if (b->h_flags) cudaFreeHost(b->h_flags);
    if (b->ev[0]) { cudaEventDestroy(b->ev[0]); cudaEventDestroy(b->ev[1]); }
    if (b->up_ev) cudaEventDestroy(b->up_ev);
    if (b->done_ev) cudaEventDestroy(b->done_ev);
    if (b->dl_stream) cudaStreamDestroy(b->dl_stream);
    for (cudaEvent_t e : b->kev) cudaEventDestroy(e);
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    delete b;
}

namespace {
// runs fn(0..n-1) on up to 16 host threads (problems are independent); inline for a single item
template <class F>
void parallel_for(int n, F fn) {
    // host threads for validation / flattening: ORB_HOST_THREADS (e.g. cores / ranks when several processes share a node), else up to 16
    static const int cap = []() { const char* e = getenv("ORB_HOST_THREADS"); const int v = e ? atoi(e) : 0; return v > 0 ? std::min(v, 64) : 16; }();
    const int nt = std::min({n, cap, (int)std::max(1u, std::thread::hardware_concurrency())});
    if (nt <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; t++) ts.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < n;) fn(i); });
    for (std::thread& t : ts) t.join();
}
struct Layout {   // bump allocator over one buffer; offsets are 256-byte aligned
    size_t cur = 0;
    size_t add(size_t bytes) { const size_t off = (cur + 255) & ~(size_t)255; cur = off + bytes; return off; }
};
}  // namespace

static int launch_steps(orbba* b, int steps) {
    const BABatch& A = b->A;
    cudaStream_t st = b->stream;
    const int hs_n = std::min(b->max_n, b->hs_smem_n);   // problems above the limit keep their matrix in global memory
    const int hs_doubles = (hs_n * (hs_n + 1) / 2 + 1) & ~1;
    const size_t smem = ((size_t)hs_doubles + std::max(b->max_n, 6) + 6 * (size_t)std::max(hs_n, 6)) * sizeof(double);   // matrix + solution column + panel
    for (int s = 0; s < steps; s++) {
        cudaEvent_t* kv = (b->profile && b->kev_steps < BA_KEV_STEPS && !b->kev.empty()) ? &b->kev[(size_t)7 * b->kev_steps] : nullptr;
        if (kv) cudaEventRecord(kv[0], st);
        const int cap = 148 * 16;                      // one wave
        k_lin<<<std::min(b->nbE, cap), BA_TE, 0, st>>>(A, b->nbE);
        if (kv) cudaEventRecord(kv[1], st);
        k_build<0><<<std::min(b->nbL, cap), BA_TL, 0, st>>>(A, b->nbL, b->d_pose_prob);
        if (b->Ktot > 0) k_build<1><<<std::min(b->Ktot, cap), BA_TL, 0, st>>>(A, b->Ktot, b->d_pose_prob);
        if (kv) cudaEventRecord(kv[2], st);
        if (b->nbG > 0) k_land<<<b->nbG, BA_TG, 0, st>>>(A);
        if (kv) cudaEventRecord(kv[3], st);
        if (b->nbI > 0) k_pairs<<<std::min(b->nbI, b->pairs_grid), 128, 0, st>>>(A, b->nbI * 4);
        if (kv) cudaEventRecord(kv[4], st);
        k_solve<<<b->n, BA_TS, smem, st>>>(A, hs_doubles, std::max(b->max_n, 6), b->hs_smem_n);
        if (kv) cudaEventRecord(kv[5], st);
        k_back<<<b->nbL, BA_TL, 0, st>>>(A);
        if (kv) { cudaEventRecord(kv[6], st); b->kev_steps++; }
        b->launches += 4 + (b->nbG > 0) + (b->nbI > 0) + (b->Ktot > 0);
    }
    b->h_flags[1] = 0;
    k_final<<<b->nbE, BA_TE, 0, st>>>(A);
    k_stats<<<1, 256, 0, st>>>(A);
    b->launches += 2;
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaEventRecord(b->done_ev, st));
    return ORB_OK;
}

// waits for the enqueued steps; problems that needed more LM trials than were enqueued get further steps
static int finish(orbba* b) {
    if (!b->pending) return ORB_OK;
    ORB_CUDA(cudaEventSynchronize(b->done_ev));      // only this handle's steps: later work on the same stream is not waited for
    int guard = 0;
    while (b->h_flags[1] > 0 && guard++ < 64) {
        int rc = launch_steps(b, 4);
        if (rc != ORB_OK) return rc;
        ORB_CUDA(cudaEventSynchronize(b->done_ev));
    }
    b->pending = false;
    return ORB_OK;
}

// accessors for orb_pose.cu (PoseOptimization shares the handle's device and stream)
cudaStream_t orbba_stream_of(orbba* b) { return b->stream; }
int orbba_device_of(orbba* b) { return b->device; }
void orbba_count_launches(orbba* b, int n) { b->launches += n; }

extern "C" {

int orbba_create(orbba_t** out, int device, int max_problems) {
    if (!out) ORB_FAIL(ORB_E_INVALID, "orbba_create: out is NULL");
    *out = nullptr;
    if (max_problems < 1 || max_problems > 65535) ORB_FAIL(ORB_E_INVALID, "orbba_create: max_problems out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_create: device %d not present", device);
    cudaDeviceProp prop;
    ORB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_create: device %d is sm_%d%d, the kernels are built for sm_100a only", device, prop.major, prop.minor);
    ORB_CUDA(cudaSetDevice(device));
    orbba* b = new (std::nothrow) orbba();
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_create: out of host memory");
    b->device = device; b->max_problems = max_problems;
    memset(&b->A, 0, sizeof(b->A));
    cudaError_t ce = cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaHostAlloc((void**)&b->h_flags, 4 * sizeof(int), cudaHostAllocMapped);
    if (ce == cudaSuccess) { memset(b->h_flags, 0, 4 * sizeof(int)); ce = cudaHostGetDevicePointer((void**)&b->d_flags, b->h_flags, 0); }
    if (ce == cudaSuccess) ce = cudaEventCreate(&b->ev[0]);
    if (ce == cudaSuccess) ce = cudaEventCreate(&b->ev[1]);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&b->up_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&b->done_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&b->dl_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
    if (ce == cudaSuccess) {
        int per_sm = 0;
        ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pairs, 128, 0);
        if (ce == cudaSuccess) b->pairs_grid = std::max(1, per_sm) * prop.multiProcessorCount;
    }
    if (ce != cudaSuccess) { int rc = orbhost::check_cuda(ce, "orbba_create", __FILE__, __LINE__); orbba_free(b); return rc; }
    b->stream = b->own_stream;
    *out = b;
    return ORB_OK;
}

void orbba_destroy(orbba_t* b) { orbba_free(b); }

int orbba_set_stream(orbba_t* b, void* s) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_set_stream: NULL handle");
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    b->stream = s ? (cudaStream_t)s : b->own_stream;
    return ORB_OK;
}
int orbba_set_copy_stream(orbba_t* b, void* s) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_set_copy_stream: NULL handle");
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    b->copy_stream = (cudaStream_t)s;
    return ORB_OK;
}
int orbba_synchronize(orbba_t* b) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_synchronize: NULL handle");
    ORB_CUDA(cudaSetDevice(b->device));
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaStreamSynchronize(b->stream));
    return ORB_OK;
}
long long orbba_launch_count(const orbba_t* b) { return b ? b->launches : 0; }

}  // extern "C"

// Validates, concatenates and uploads a batch of problems, then builds the (pose pair -> edge tuples) index on the device.
static int upload_views(orbba* b, const std::vector<ProbView>& problems, int n, bool compact) {
    if (n < 0 || n > b->max_problems) ORB_FAIL(ORB_E_INVALID, "orbba_upload: n=%d exceeds max_problems=%d", n, b->max_problems);
    ORB_CUDA(cudaSetDevice(b->device));
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    b->n = 0;
    b->probs.assign(n, BAProb());
    b->perm.assign(n, std::vector<int>());
    if (n == 0) return ORB_OK;
    // ---- pass 1a: per-problem analysis (validation, free-pose numbering, landmark grouping, tuple count) on all host cores
    std::vector<std::vector<int>> pose_free_local(n);
    std::vector<std::vector<int>> groups(n);   // k_land blocks per problem: (first landmark, landmarks) pairs
    std::vector<long long> tups(n, 0);
    std::vector<int> Ks(n, 0);
    std::vector<std::string> errs(n);
    parallel_for(n, [&](int p) {
        const ProbView& Q = problems[p];
        const int nP = Q.nP, nL = Q.nL, nE = Q.nE, nC = Q.nC;
        char msg[256];
        if (nP < 0 || nL < 0 || nE < 0 || nC < 1) { snprintf(msg, sizeof(msg), "orbba_upload: problem %d has negative sizes", p); errs[p] = msg; return; }
        if (nC * nC > BA_MAXCC) { snprintf(msg, sizeof(msg), "orbba_upload: problem %d has a rig of %d cameras (at most 4)", p, nC); errs[p] = msg; return; }
        if (Q.has_nulls()) { snprintf(msg, sizeof(msg), "orbba_upload: problem %d has a NULL array", p); errs[p] = msg; return; }
        if (compact && (nP > 65535 || Q.nLev < 1 || Q.nLev > 255)) { snprintf(msg, sizeof(msg), "orbba_upload_f32: problem %d: at most 65535 poses, 1..255 levels", p); errs[p] = msg; return; }
        std::vector<int>& pf = pose_free_local[p];
        pf.assign(nP, -1);
        int K = 0;
        for (int i = 0; i < nP; i++) if (!Q.pose_fixed[i]) pf[i] = K++;
        Ks[p] = K;
        bool grouped = true;
        for (int e = 0; e < nE; e++) {
            if (Q.pose(e) < 0 || Q.pose(e) >= nP || Q.point(e) < 0 || Q.point(e) >= nL || Q.cam(e) < 0 || Q.cam(e) >= nC || (compact && Q.e16[e].octave >= Q.nLev)) {
                snprintf(msg, sizeof(msg), "orbba_upload: problem %d edge %d indexes out of range", p, e); errs[p] = msg; return;
            }
            if (e > 0 && Q.point(e) < Q.point(e - 1)) grouped = false;
        }
        if (!grouped) {   // stable counting sort by landmark (the reference adds edges landmark by landmark, src/Optimizer.cc:530-575)
            std::vector<int> cnt(nL + 1, 0);
            for (int e = 0; e < nE; e++) cnt[Q.point(e) + 1]++;
            for (int l = 0; l < nL; l++) cnt[l + 1] += cnt[l];
            b->perm[p].resize(nE);
            for (int e = 0; e < nE; e++) b->perm[p][cnt[Q.point(e)]++] = e;
        }
        // free-pose observations per landmark; one observation per (landmark, keyframe) as in MapPoint::mObservations
        std::vector<int> stamp(nP, -1);
        long long tup = 0;
        int run = 0, prev = -1;
        const std::vector<int>& pm = b->perm[p];
        for (int s = 0; s < nE; s++) {
            const int e = pm.empty() ? s : pm[s];
            const int l = Q.point(e), ps = Q.pose(e);
            if (stamp[ps] == l) {
                snprintf(msg, sizeof(msg), "orbba_upload: problem %d observes landmark %d twice from pose %d (MapPoint::mObservations holds one per keyframe)", p, l, ps);
                errs[p] = msg; return;
            }
            stamp[ps] = l;
            if (l != prev) { tup += (long long)run * (run + 1) / 2; run = 0; prev = l; }
            if (pf[ps] >= 0) run++;
        }
        tup += (long long)run * (run + 1) / 2;
        tups[p] = tup;
        {   // k_land blocks: greedy packing of whole landmarks into groups of <= BA_TG edges (a larger landmark stands alone)
            std::vector<int> cntl(nL, 0);
            for (int e = 0; e < nE; e++) cntl[Q.point(e)]++;
            std::vector<int>& G = groups[p];
            int first = 0, ne = 0;
            for (int l = 0; l < nL; l++) {
                const bool full = l > first && (ne + cntl[l] > BA_TG || l - first >= BA_TG);
                if (full) { G.push_back(first); G.push_back(l - first); first = l; ne = 0; }
                ne += cntl[l];
            }
            if (nL > first) { G.push_back(first); G.push_back(nL - first); }
        }
        if (tup > 0x7fffffffLL || (long long)K * (K + 1) / 2 > 4000000LL) { snprintf(msg, sizeof(msg), "orbba_upload: problem %d is too large for the pair index (%d free poses)", p, K); errs[p] = msg; return; }
        if (6 * K > 3000) { snprintf(msg, sizeof(msg), "orbba_upload: problem %d has %d free poses; above 500 use the distributed global BA entry points", p, K); errs[p] = msg; }
    });
    for (int p = 0; p < n; p++) if (!errs[p].empty()) ORB_FAIL(ORB_E_INVALID, "%s", errs[p].c_str());
    // ---- pass 1b: offsets
    long long Etot = 0, Ltot = 0, Ptot = 0, Ctot = 0, Ktot = 0, pairTot = 0, tupTot = 0, chunkTot = 0, eofTot = 0, hsTot = 0, itemTot = 0;
    long long rtTot = 0, pcTot = 0, utTot = 0, bmTot = 0;
    int nbE = 0, nbL = 0, nbP = 0, nbI = 0, nbG = 0, max_n = 0;
    for (int p = 0; p < n; p++) max_n = std::max(max_n, 6 * Ks[p]);
    // shared memory of k_solve: packed matrix of the largest in-shared problem + solution column (any n) + panel; a very large problem in the
    // batch sends every matrix to global memory
    int hs_smem_n = BA_HS_SMEM_N;
    if (((size_t)hs_smem_n * (hs_smem_n + 1) / 2 + 2 + max_n + 6 * (size_t)hs_smem_n) * 8 > 200 * 1024) hs_smem_n = 0;
    std::vector<int> bP0(n), bI0(n);
    for (int p = 0; p < n; p++) {
        const ProbView& Q = problems[p];
        const int nP = Q.nP, nL = Q.nL, nE = Q.nE, nC = Q.nC, K = Ks[p];
        const long long tup = tups[p];
        BAProb& P = b->probs[p];
        P.e0 = (int)Etot; P.nE = nE; P.l0 = (int)Ltot; P.nL = nL; P.p0 = (int)Ptot; P.nP = nP; P.c0 = (int)Ctot; P.nC = nC;
        P.k0 = (int)Ktot; P.K = K; P.n = 6 * K;
        P.pair0 = (int)pairTot; P.nPairs = K * (K + 1) / 2;
        P.CC = nC * nC;
        P.rt0 = (int)rtTot; P.pc0 = (int)pcTot; P.ut0 = (int)utTot;
        rtTot += (long long)nP * nC; pcTot += (long long)P.nPairs * P.CC; utTot += (long long)K * nC;
        P.chunk0 = (int)chunkTot; P.nChunksMax = P.nPairs * P.CC + (int)(tup / BA_CH);
        P.item0 = (int)itemTot; P.nItems = P.nChunksMax;
        P.blkE0 = nbE; P.nbE = std::max(1, (nE + BA_TE - 1) / BA_TE);
        P.blkL0 = nbL; P.nbL = std::max(1, (nL + BA_TL - 1) / BA_TL);
        P.blkG0 = nbG; P.nbG = (int)groups[p].size() / 2;
        P.tup0 = tupTot; P.eof0 = eofTot; P.hs_off = hsTot;
        P.bmW = (nL + 31) / 32; P.bm0 = bmTot; bmTot += (long long)K * nC * P.bmW;
        bP0[p] = nbP; bI0[p] = nbI; P.blkI0 = nbI;
        Etot += nE; Ltot += nL; Ptot += nP; Ctot += nC; Ktot += K; pairTot += P.nPairs; tupTot += tup; chunkTot += P.nChunksMax;
        itemTot += P.nItems; eofTot += (long long)nL * K;
        if (P.n > hs_smem_n) hsTot += (long long)std::max(P.n, 14) * (std::max(P.n, 14) | 1);
        nbE += P.nbE; nbL += P.nbL; nbG += P.nbG; nbP += (P.nPairs + 3) / 4; nbI += (P.nItems + 3) / 4;
        max_n = std::max(max_n, P.n);
        if (Etot > 0x1fffffffLL || eofTot > 0x7fffffffLL * 4 || pcTot > 0x7fffffffLL || chunkTot > 0x7fffffffLL || tupTot > 0x7fffffffLL) ORB_FAIL(ORB_E_INVALID, "orbba_upload: batch too large");
    }
    // ---- layout: static (staged from the host) then device-only
    Layout L;
    const size_t o_prob = L.add(sizeof(BAProb) * n);
    // the per-edge / per-point FP64 inputs are staged from the host (orbba_upload) or produced on the device from the staged compact
    // records (orbba_upload_f32): 16 instead of 36 bytes per edge, 12 instead of 24 per point over PCIe
    size_t o_epose = 0, o_ept = 0, o_ecam = 0, o_eobs = 0, o_einfo = 0, o_pt0 = 0, o_e16 = 0, o_pts32 = 0, o_isig = 0, o_isig0 = 0;
    long long levTot = 0;
    for (int p = 0; p < n; p++) levTot += problems[p].nLev;
    if (!compact) {
        o_epose = L.add(4 * Etot); o_ept = L.add(4 * Etot); o_ecam = L.add(4 * Etot); o_eobs = L.add(16 * Etot); o_einfo = L.add(8 * Etot); o_pt0 = L.add(24 * Ltot);
    } else {
        o_e16 = L.add(16 * Etot); o_pts32 = L.add(12 * Ltot); o_isig = L.add(4 * (size_t)std::max<long long>(levTot, 1)); o_isig0 = L.add(4 * (size_t)n);
    }
    const size_t o_cam = L.add(8 * BA_CAM_STRIDE * Ctot);
    const size_t o_ptoff = L.add(4 * (Ltot + 1)), o_pfree = L.add(4 * Ptot), o_pose0 = L.add(56 * Ptot);
    const size_t o_blkE = L.add(4 * (size_t)nbE), o_blkL = L.add(4 * (size_t)nbL), o_blkLl = L.add(4 * (size_t)nbL), o_item = L.add(4 * (size_t)std::max(nbI, 1));
    const size_t o_blkPp = L.add(4 * (size_t)std::max(nbP, 1)), o_blkPf = L.add(4 * (size_t)std::max(nbP, 1)), o_blkIf = L.add(4 * (size_t)std::max(nbI, 1));
    const size_t o_poseprob = L.add(4 * (size_t)std::max<long long>(Ktot, 1)), o_freepose = L.add(4 * (size_t)std::max<long long>(Ktot, 1));
    const size_t o_blkGp = L.add(4 * (size_t)std::max(nbG, 1)), o_blkGl = L.add(4 * (size_t)std::max(nbG, 1)), o_blkGn = L.add(4 * (size_t)std::max(nbG, 1)), o_blkGd = L.add(16 * (size_t)std::max(nbG, 1));
    const size_t static_bytes = L.add(0);
    if (compact) {
        o_epose = L.add(4 * Etot); o_ept = L.add(4 * Etot); o_ecam = L.add(4 * Etot); o_eobs = L.add(16 * Etot); o_einfo = L.add(8 * Etot); o_pt0 = L.add(24 * Ltot);
    }
    const size_t o_state = L.add(sizeof(BAState) * n);
    const size_t o_bm = L.add(4 * (size_t)std::max<long long>(bmTot, 1));
    const size_t o_ekf = L.add(4 * (size_t)std::max<long long>(Etot, 1));
    const size_t o_eof = L.add(4 * (size_t)eofTot), o_pcnt = L.add(4 * (size_t)pairTot), o_poff = L.add(4 * (size_t)pairTot);
    const size_t o_pccnt = L.add(4 * (size_t)pcTot), o_pcoff = L.add(4 * (size_t)pcTot), o_pcfch = L.add(4 * (size_t)pcTot), o_pcnch = L.add(4 * (size_t)pcTot);
    const size_t o_cpair = L.add(4 * (size_t)chunkTot), o_cstart = L.add(4 * (size_t)chunkTot), o_clen = L.add(4 * (size_t)chunkTot);
    const size_t o_irec = L.add(32 * 4 * (size_t)std::max(nbI, 1)), o_pcount = L.add(256);   // pcount: [0] k_pairs queue head, [16] rs_flag, [32] stop_dev
    const size_t o_tup = L.add(8 * (size_t)tupTot);
    const size_t o_pose_a = L.add(56 * Ptot), o_pose_b = L.add(56 * Ptot), o_pt_a = L.add(24 * Ltot), o_pt_b = L.add(24 * Ltot);
    const size_t o_err_a = L.add(16 * Etot), o_err_b = L.add(16 * Etot), o_level = L.add(Etot);
    const size_t o_yr = L.add(8 * BA_YR * (size_t)Etot), o_rw = L.add(32 * (size_t)Etot);
    const size_t o_er = L.add(64 * (size_t)Etot), o_tab0 = L.add(8 * BA_TAB * (size_t)std::max<long long>(rtTot, 1)), o_tab1 = L.add(8 * BA_TAB * (size_t)std::max<long long>(rtTot, 1));
    const size_t o_utu = L.add(48 * (size_t)std::max<long long>(utTot, 1)), o_utb = L.add(48 * (size_t)std::max<long long>(utTot, 1)), o_uty = L.add(48 * (size_t)std::max<long long>(utTot, 1));
    const size_t o_Hll = L.add(48 * Ltot), o_bl = L.add(24 * Ltot);
    const size_t o_Hpp = L.add(288 * Ktot), o_bp = L.add(48 * Ktot), o_bs = L.add(48 * Ktot), o_xp = L.add(48 * Ktot);
    const size_t o_prhs = L.add(96 * (size_t)std::max<long long>(chunkTot, 1));
    const size_t o_partial = L.add(288 * (size_t)chunkTot), o_partE = L.add(16 * (size_t)nbE), o_partL = L.add(16 * (size_t)std::max(std::max(nbL, nbG), 1));
    const size_t o_Hs = L.add(8 * (size_t)hsTot);
    const size_t o_poses_out = L.add(96 * Ptot), o_points_out = L.add(24 * Ltot), o_outlier = L.add(Etot), o_stats = L.add(sizeof(orbba_stats_t) * n);
    const size_t total = L.add(0) + 256;
    if (total > b->arena_cap) {
        cudaFree(b->d_arena); b->d_arena = nullptr; b->arena_cap = 0;
        ORB_CUDA(cudaMalloc((void**)&b->d_arena, total));
        b->arena_cap = total;
    }
    if (static_bytes > b->stage_cap) {
        if (b->h_stage) cudaFreeHost(b->h_stage);
        b->h_stage = nullptr; b->stage_cap = 0;
        ORB_CUDA(cudaHostAlloc((void**)&b->h_stage, static_bytes + 256, cudaHostAllocDefault));
        b->stage_cap = static_bytes;
    }
    uint8_t* H = b->h_stage;
    uint8_t* D = b->d_arena;
    // ---- pass 2: fill the staging buffer
    memcpy(H + o_prob, b->probs.data(), sizeof(BAProb) * n);
    int *h_epose = (int*)(H + o_epose), *h_ept = (int*)(H + o_ept), *h_ecam = (int*)(H + o_ecam), *h_ptoff = (int*)(H + o_ptoff), *h_pfree = (int*)(H + o_pfree);
    double *h_eobs = (double*)(H + o_eobs), *h_einfo = (double*)(H + o_einfo), *h_cam = (double*)(H + o_cam), *h_pose0 = (double*)(H + o_pose0), *h_pt0 = (double*)(H + o_pt0);
    int *h_blkE = (int*)(H + o_blkE), *h_blkL = (int*)(H + o_blkL), *h_item = (int*)(H + o_item), *h_blkPp = (int*)(H + o_blkPp), *h_blkPf = (int*)(H + o_blkPf),
        *h_blkIf = (int*)(H + o_blkIf), *h_poseprob = (int*)(H + o_poseprob), *h_freepose = (int*)(H + o_freepose);
    int *h_blkGp = (int*)(H + o_blkGp), *h_blkGl = (int*)(H + o_blkGl), *h_blkGn = (int*)(H + o_blkGn);
    int4* h_blkGd = (int4*)(H + o_blkGd);
    orbba_edge16_t* h_e16 = (orbba_edge16_t*)(H + o_e16);
    float *h_pts32 = (float*)(H + o_pts32), *h_isig = (float*)(H + o_isig);
    int* h_isig0 = (int*)(H + o_isig0);
    std::vector<int> lev0(n, 0);
    for (int p = 1; p < n; p++) lev0[p] = lev0[p - 1] + problems[p - 1].nLev;
    parallel_for(n, [&](int p) {
        const int bP = bP0[p], bI = bI0[p];
        const ProbView& Q = problems[p];
        const BAProb& P = b->probs[p];
        const std::vector<int>& pm = b->perm[p];
        if (!compact) {
            for (int s = 0; s < P.nE; s++) {
                const int e = pm.empty() ? s : pm[s];
                const size_t g = (size_t)P.e0 + s;
                h_epose[g] = P.p0 + Q.ep[e]; h_ept[g] = P.l0 + Q.el[e]; h_ecam[g] = P.c0 + Q.ec[e];
                h_eobs[2 * g] = Q.obs[2 * e]; h_eobs[2 * g + 1] = Q.obs[2 * e + 1]; h_einfo[g] = Q.info[e];
            }
        } else {
            if (pm.empty()) { if (P.nE) memcpy(h_e16 + P.e0, Q.e16, sizeof(orbba_edge16_t) * (size_t)P.nE); }
            else for (int s = 0; s < P.nE; s++) h_e16[(size_t)P.e0 + s] = Q.e16[pm[s]];
            if (P.nL) memcpy(h_pts32 + 3 * (size_t)P.l0, Q.points32, sizeof(float) * 3 * (size_t)P.nL);
            for (int l = 0; l < Q.nLev; l++) h_isig[lev0[p] + l] = Q.isig[l];
            h_isig0[p] = lev0[p];
        }
        {   // CSR by landmark over the grouped edge order
            int s = 0;
            for (int l = 0; l < P.nL; l++) {
                h_ptoff[P.l0 + l] = P.e0 + s;
                while (s < P.nE && Q.point(pm.empty() ? s : pm[s]) == l) s++;
            }
        }
        for (int i = 0; i < P.nP; i++) h_pfree[P.p0 + i] = pose_free_local[p][i] >= 0 ? P.k0 + pose_free_local[p][i] : -1;
        for (int k = 0; k < P.K; k++) h_poseprob[P.k0 + k] = p;
        for (int i = 0; i < P.nP; i++) if (pose_free_local[p][i] >= 0) h_freepose[P.k0 + pose_free_local[p][i]] = P.p0 + i;
        for (int c = 0; c < P.nC; c++) {
            double* Dc = h_cam + (size_t)BA_CAM_STRIDE * (P.c0 + c);
            for (int i = 0; i < 4; i++) Dc[i] = Q.cam_K[4 * c + i];
            const double* T = Q.cam_ext + 12 * c;
            const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
            q_from_matrix(R, Dc + 4);
            if (Dc[7] < 0) for (int i = 4; i < 8; i++) Dc[i] = -Dc[i];
            const double nn = sqrt(Dc[4] * Dc[4] + Dc[5] * Dc[5] + Dc[6] * Dc[6] + Dc[7] * Dc[7]);
            for (int i = 4; i < 8; i++) Dc[i] /= nn;
            Dc[8] = T[3]; Dc[9] = T[7]; Dc[10] = T[11];
            for (int i = 0; i < 36; i++) Dc[BA_CAM_ADJ + i] = Q.cam_adj[36 * c + i];
        }
        for (int i = 0; i < P.nP; i++) {   // Converter::toSE3Quat + SE3Quat(R, t)
            double T[12];
            for (int k = 0; k < 12; k++) T[k] = Q.pose_el(i, k);
            double* Dp = h_pose0 + 7 * (size_t)(P.p0 + i);
            const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
            q_from_matrix(R, Dp);
            if (Dp[3] < 0) for (int k = 0; k < 4; k++) Dp[k] = -Dp[k];
            const double nn = sqrt(Dp[0] * Dp[0] + Dp[1] * Dp[1] + Dp[2] * Dp[2] + Dp[3] * Dp[3]);
            for (int k = 0; k < 4; k++) Dp[k] /= nn;
            Dp[4] = T[3]; Dp[5] = T[7]; Dp[6] = T[11];
        }
        if (!compact && P.nL) memcpy(h_pt0 + 3 * (size_t)P.l0, Q.points, sizeof(double) * 3 * P.nL);
        for (int q = 0; q < P.nbE; q++) h_blkE[P.blkE0 + q] = p;
        for (int q = 0; q < P.nbL; q++) { h_blkL[P.blkL0 + q] = p; ((int*)(H + o_blkLl))[P.blkL0 + q] = P.l0 + q * BA_TL; }
        for (int q = 0; q < P.nbG; q++) { h_blkGp[P.blkG0 + q] = p; h_blkGl[P.blkG0 + q] = P.l0 + groups[p][2 * q]; h_blkGn[P.blkG0 + q] = groups[p][2 * q + 1]; }
        for (int q = 0; q < P.nbG; q++) {
            const int gl0 = groups[p][2 * q], gnl = groups[p][2 * q + 1];
            const int ge0 = h_ptoff[P.l0 + gl0], ge1 = gl0 + gnl >= P.nL ? P.e0 + P.nE : h_ptoff[P.l0 + gl0 + gnl];
            h_blkGd[P.blkG0 + q] = make_int4(p, P.l0 + gl0, ge0, gnl | ((ge1 - ge0) << 8));
        }
        const int bp4 = (P.nPairs + 3) / 4, bi4 = (P.nItems + 3) / 4;
        for (int q = 0; q < bp4; q++) { h_blkPp[bP + q] = p; h_blkPf[bP + q] = bP; }
        for (int q = 0; q < bi4; q++) { h_item[bI + q] = p; h_blkIf[bI + q] = bI; }
    });
    h_ptoff[Ltot] = (int)Etot;
    // ---- device pointers
    BABatch& A = b->A;
    memset(&A, 0, sizeof(A));
    A.nProb = n;
    A.prob = (const BAProb*)(D + o_prob); A.state = (BAState*)(D + o_state);
    A.e_pose = (const int*)(D + o_epose); A.e_pt = (const int*)(D + o_ept); A.e_cam = (const int*)(D + o_ecam);
    A.e_obs = (const double*)(D + o_eobs); A.e_info = (const double*)(D + o_einfo); A.cam = (const double*)(D + o_cam);
    A.pt_off = (const int*)(D + o_ptoff); A.pose_free = (const int*)(D + o_pfree);
    A.pose0 = (const double*)(D + o_pose0); A.pt0 = (const double*)(D + o_pt0);
    A.blkE_prob = (const int*)(D + o_blkE); A.blkL_prob = (const int*)(D + o_blkL); A.blkL_l0 = (const int*)(D + o_blkLl); A.item_prob = (const int*)(D + o_item);
    A.blkG_prob = (const int*)(D + o_blkGp); A.blkG_l0 = (const int*)(D + o_blkGl); A.blkG_nl = (const int*)(D + o_blkGn); A.blkG_desc = (const int4*)(D + o_blkGd);
    b->d_blkP_prob = (int*)(D + o_blkPp); b->d_blkP_first = (int*)(D + o_blkPf); b->d_blkI_first = (int*)(D + o_blkIf); b->d_pose_prob = (int*)(D + o_poseprob);
    A.bm = (unsigned*)(D + o_bm);
    A.e_kf = (int*)(D + o_ekf);
    A.edge_of = (int*)(D + o_eof); A.pair_cnt = (int*)(D + o_pcnt); A.pair_off = (int*)(D + o_poff);
    A.pc_cnt = (int*)(D + o_pccnt); A.pc_off = (int*)(D + o_pcoff); A.pc_fchunk = (int*)(D + o_pcfch); A.pc_nchunk = (int*)(D + o_pcnch);
    A.free_pose = (const int*)(D + o_freepose);
    A.chunk_pair = (int*)(D + o_cpair); A.chunk_start = (int*)(D + o_cstart); A.chunk_len = (int*)(D + o_clen);
    A.item_rec = (int4*)(D + o_irec); A.pairs_counter = (int*)(D + o_pcount); A.rs_flag = (int*)(D + o_pcount) + 16; A.stop_dev = (int*)(D + o_pcount) + 32;
    A.tuples = (int2*)(D + o_tup);
    A.pose[0] = (double*)(D + o_pose_a); A.pose[1] = (double*)(D + o_pose_b); A.pt[0] = (double*)(D + o_pt_a); A.pt[1] = (double*)(D + o_pt_b);
    A.err[0] = (double*)(D + o_err_a); A.err[1] = (double*)(D + o_err_b); A.level = D + o_level;
    A.yr = (double*)(D + o_yr); A.rw = (double*)(D + o_rw);
    A.er = (double*)(D + o_er); A.tab[0] = (double*)(D + o_tab0); A.tab[1] = (double*)(D + o_tab1); A.ut_u = (double*)(D + o_utu); A.ut_b = (double*)(D + o_utb); A.ut_y = (double*)(D + o_uty);
    A.Hll = (double*)(D + o_Hll); A.bl = (double*)(D + o_bl);
    A.Hpp = (double*)(D + o_Hpp); A.bp = (double*)(D + o_bp); A.bs = (double*)(D + o_bs); A.xp = (double*)(D + o_xp);
    A.partial = (double*)(D + o_partial); A.prhs = (double*)(D + o_prhs); A.partE = (double*)(D + o_partE); A.partL = (double*)(D + o_partL); A.Hs = (double*)(D + o_Hs);
    A.poses_out = (double*)(D + o_poses_out); A.points_out = (double*)(D + o_points_out); A.outlier = D + o_outlier;
    A.stats = (orbba_stats_t*)(D + o_stats);
    A.stop = b->d_flags; A.n_active = b->d_flags + 1;
    b->nbE = nbE; b->nbL = nbL; b->nbP = nbP; b->nbI = nbI; b->nbG = nbG; b->Ktot = (int)Ktot; b->max_n = max_n; b->hs_smem_n = hs_smem_n;
    b->Etot = Etot; b->Ltot = Ltot; b->Ptot = Ptot;
    // ---- upload + index construction on the device (on the copy stream when one is set: overlaps with a run of another handle)
    cudaStream_t st = b->copy_stream ? b->copy_stream : b->stream;
    ORB_CUDA(cudaMemcpyAsync(D, H, static_bytes, cudaMemcpyHostToDevice, st));
    if (compact) {
        BACompact C;
        C.e16 = (const orbba_edge16_t*)(D + o_e16); C.pts32 = (const float*)(D + o_pts32); C.isig = (const float*)(D + o_isig); C.isig0 = (const int*)(D + o_isig0);
        C.e_pose = (int*)(D + o_epose); C.e_pt = (int*)(D + o_ept); C.e_cam = (int*)(D + o_ecam);
        C.e_obs = (double*)(D + o_eobs); C.e_info = (double*)(D + o_einfo); C.pt0 = (double*)(D + o_pt0);
        k_expand<<<nbE, BA_TE, 0, st>>>(A, C);
        b->launches++;
    }
    ORB_CUDA(cudaMemsetAsync(D + o_state, 0, sizeof(BAState) * n, st));
    ORB_CUDA(cudaMemsetAsync(D + o_irec, 0, 32 * 4 * (size_t)std::max(nbI, 1), st));
    if (eofTot) ORB_CUDA(cudaMemsetAsync(D + o_eof, 0xff, 4 * (size_t)eofTot, st));
    if (bmTot) ORB_CUDA(cudaMemsetAsync(D + o_bm, 0, 4 * (size_t)bmTot, st));
    k_edge_of<<<nbE, BA_TE, 0, st>>>(A);
    if (nbP > 0) {
        k_pair_count<<<nbP, 128, 0, st>>>(A, b->d_blkP_prob, b->d_blkP_first);
        k_pair_scan<<<n, 256, 0, st>>>(A);
        k_pair_fill<<<nbP, 128, 0, st>>>(A, b->d_blkP_prob, b->d_blkP_first);
        b->launches += 3;
    }
    b->launches += 1;
    ORB_CUDA(cudaGetLastError());
    if (b->copy_stream) { ORB_CUDA(cudaEventRecord(b->up_ev, st)); b->up_pending = true; }
    b->n = n;
    return ORB_OK;
}

extern "C" {

int orbba_upload(orbba_t* b, const orbba_problem_t* problems, int n) {
    if (!b || (!problems && n > 0)) ORB_FAIL(ORB_E_INVALID, "orbba_upload: bad argument");
    std::vector<ProbView> V((size_t)std::max(n, 0));
    for (int p = 0; p < n; p++) V[p] = view_of(problems[p]);
    return upload_views(b, V, n, false);
}
// The same batch from the reference's own storage types: CV_32F poses / points, 16-byte edge records, per-level weights.
int orbba_upload_f32(orbba_t* b, const orbba_problem_f32_t* problems, int n) {
    if (!b || (!problems && n > 0)) ORB_FAIL(ORB_E_INVALID, "orbba_upload_f32: bad argument");
    std::vector<ProbView> V((size_t)std::max(n, 0));
    for (int p = 0; p < n; p++) V[p] = view_of(problems[p]);
    return upload_views(b, V, n, true);
}

// Runs the uploaded batch from its uploaded initial estimates (asynchronous on the handle's stream).
// its2 < 0: single round, no outlier pass (Optimizer::BundleAdjustment); huber_delta <= 0: no robust kernel.
int orbba_run(orbba_t* b, int its1, int its2, double huber_delta, double chi2_th) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_run: NULL handle");
    if (b->n == 0) return ORB_OK;
    if (its1 < 0) ORB_FAIL(ORB_E_INVALID, "orbba_run: its1 < 0");
    ORB_CUDA(cudaSetDevice(b->device));
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    b->A.its1 = its1; b->A.its2 = its2; b->A.delta = huber_delta; b->A.chi2_th = chi2_th;
    if (b->up_pending) { ORB_CUDA(cudaStreamWaitEvent(b->stream, b->up_ev, 0)); b->up_pending = false; }
    if (b->profile) ORB_CUDA(cudaEventRecord(b->ev[0], b->stream));
    k_reset<<<b->nbE, BA_TE, 0, b->stream>>>(b->A, b->h_flags[0] != 0);
    b->launches++;
    // one step per LM trial; accepted-first-try iterations need its1 + its2 steps, plus one step per round switch
    rc = launch_steps(b, its1 + std::max(its2, 0) + 3);
    if (rc != ORB_OK) return rc;
    if (b->profile) { ORB_CUDA(cudaEventRecord(b->ev[1], b->stream)); b->prof_pending = true; }
    b->pending = true;
    return ORB_OK;
}

int orbba_profile(orbba_t* b, int enable) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_profile: NULL handle");
    ORB_CUDA(cudaSetDevice(b->device));
    if (enable && b->kev.empty()) {
        b->kev.resize((size_t)BA_KEV_STEPS * 7);
        for (cudaEvent_t& e : b->kev) ORB_CUDA(cudaEventCreate(&e));
    }
    b->kev_steps = 0;
    b->profile = enable != 0; b->prof_ms = 0; b->prof_calls = 0; b->prof_pending = false;
    return ORB_OK;
}
int orbba_stage_ms(orbba_t* b, double* ms1, int* calls) {   // only the LAST run is kept per synchronisation: call after each run
    if (!b || !ms1) ORB_FAIL(ORB_E_INVALID, "orbba_stage_ms: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaStreamSynchronize(b->stream));
    if (b->prof_pending) {
        float ms = 0;
        ORB_CUDA(cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]));
        b->prof_ms += ms; b->prof_calls++; b->prof_pending = false;
    }
    *ms1 = b->prof_ms;
    if (calls) *calls = b->prof_calls;
    b->prof_ms = 0; b->prof_calls = 0;
    return ORB_OK;
}

// Device time per kernel of the LM step {k_lin, k_build, k_trial_lm, k_pairs, k_solve, k_back}, summed over the steps recorded
// since the last call (profiling must be enabled); *steps = number of steps summed.
int orbba_kernel_ms(orbba_t* b, double* ms6, int* steps) {
    if (!b || !ms6) ORB_FAIL(ORB_E_INVALID, "orbba_kernel_ms: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaStreamSynchronize(b->stream));
    for (int k = 0; k < 6; k++) ms6[k] = 0;
    for (int s = 0; s < b->kev_steps; s++)
        for (int k = 0; k < 6; k++) {
            float ms = 0;
            ORB_CUDA(cudaEventElapsedTime(&ms, b->kev[(size_t)7 * s + k], b->kev[(size_t)7 * s + k + 1]));
            ms6[k] += ms;
        }
    if (steps) *steps = b->kev_steps;
    b->kev_steps = 0;
    return ORB_OK;
}

// Results of every problem of the last run, concatenated in upload order: poses_out [sum n_poses][12], points_out
// [sum n_points][3], edge_outlier [sum n_edges] (caller's edge order), stats [n].  Any pointer may be NULL.  Synchronises.
int orbba_download_batch(orbba_t* b, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_download_batch: NULL handle");
    if (b->n == 0) return ORB_OK;
    ORB_CUDA(cudaSetDevice(b->device));
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    cudaStream_t st = b->dl_stream;
    if (poses_out && b->Ptot) ORB_CUDA(cudaMemcpyAsync(poses_out, b->A.poses_out, sizeof(double) * 12 * b->Ptot, cudaMemcpyDeviceToHost, st));
    if (points_out && b->Ltot) ORB_CUDA(cudaMemcpyAsync(points_out, b->A.points_out, sizeof(double) * 3 * b->Ltot, cudaMemcpyDeviceToHost, st));
    if (edge_outlier && b->Etot) ORB_CUDA(cudaMemcpyAsync(edge_outlier, b->A.outlier, (size_t)b->Etot, cudaMemcpyDeviceToHost, st));
    if (stats) ORB_CUDA(cudaMemcpyAsync(stats, b->A.stats, sizeof(orbba_stats_t) * b->n, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    if (edge_outlier)
        for (int p = 0; p < b->n; p++) {
            const std::vector<int>& pm = b->perm[p];
            if (pm.empty()) continue;
            const BAProb& P = b->probs[p];
            std::vector<uint8_t> tmp(edge_outlier + P.e0, edge_outlier + P.e0 + P.nE);
            for (int s2 = 0; s2 < P.nE; s2++) edge_outlier[P.e0 + pm[s2]] = tmp[s2];
        }
    return ORB_OK;
}

// Copies results of problem `p` of the last run to the host (synchronises the stream).  Any pointer may be NULL.
int orbba_download(orbba_t* b, int p, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats) {
    if (!b || p < 0 || p >= b->n) ORB_FAIL(ORB_E_INVALID, "orbba_download: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
    int rc = finish(b);
    if (rc != ORB_OK) return rc;
    const BAProb& P = b->probs[p];
    cudaStream_t st = b->dl_stream;
    std::vector<uint8_t> tmp;
    if (poses_out && P.nP) ORB_CUDA(cudaMemcpyAsync(poses_out, b->A.poses_out + 12 * (size_t)P.p0, sizeof(double) * 12 * P.nP, cudaMemcpyDeviceToHost, st));
    if (points_out && P.nL) ORB_CUDA(cudaMemcpyAsync(points_out, b->A.points_out + 3 * (size_t)P.l0, sizeof(double) * 3 * P.nL, cudaMemcpyDeviceToHost, st));
    if (edge_outlier && P.nE) {
        if (b->perm[p].empty()) ORB_CUDA(cudaMemcpyAsync(edge_outlier, b->A.outlier + P.e0, P.nE, cudaMemcpyDeviceToHost, st));
        else { tmp.resize(P.nE); ORB_CUDA(cudaMemcpyAsync(tmp.data(), b->A.outlier + P.e0, P.nE, cudaMemcpyDeviceToHost, st)); }
    }
    if (stats) ORB_CUDA(cudaMemcpyAsync(stats, b->A.stats + p, sizeof(orbba_stats_t), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    if (!tmp.empty()) for (int s = 0; s < P.nE; s++) edge_outlier[b->perm[p][s]] = tmp[s];
    return ORB_OK;
}

// Optimizer::LocalBundleAdjustment for ONE problem, host buffers in and out, synchronous.  `stop` (may be NULL) is
// polled while the kernels run and forwarded to the device, which reads it after every LM trial
// (g2o: SparseOptimizer::terminate(), sparse_optimizer.cpp:376, optimization_algorithm_levenberg.cpp:149).
static int local_run(orbba* b, int its1, int its2, double huber_delta, double chi2_th, const volatile uint8_t* stop, double* poses_out, double* points_out,
                     uint8_t* edge_outlier, orbba_stats_t* stats);

int orbba_local(orbba_t* b, const orbba_problem_t* problem, int its1, int its2, double huber_delta, double chi2_th,
                const volatile uint8_t* stop, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats) {
    if (!b || !problem) ORB_FAIL(ORB_E_INVALID, "orbba_local: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
    b->h_flags[0] = (stop && *stop) ? 1 : 0;
    int rc = orbba_upload(b, problem, 1);
    if (rc != ORB_OK) return rc;
    return local_run(b, its1, its2, huber_delta, chi2_th, stop, poses_out, points_out, edge_outlier, stats);
}
// the same from the compact form (what the reference holds: CV_32F poses / points, float key points, per-level weights)
int orbba_local_f32(orbba_t* b, const orbba_problem_f32_t* problem, int its1, int its2, double huber_delta, double chi2_th,
                    const volatile uint8_t* stop, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats) {
    if (!b || !problem) ORB_FAIL(ORB_E_INVALID, "orbba_local_f32: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
    b->h_flags[0] = (stop && *stop) ? 1 : 0;
    int rc = orbba_upload_f32(b, problem, 1);
    if (rc != ORB_OK) return rc;
    return local_run(b, its1, its2, huber_delta, chi2_th, stop, poses_out, points_out, edge_outlier, stats);
}
static int local_run(orbba* b, int its1, int its2, double huber_delta, double chi2_th, const volatile uint8_t* stop, double* poses_out, double* points_out,
                     uint8_t* edge_outlier, orbba_stats_t* stats) {
    int rc;
    rc = orbba_run(b, its1, its2, huber_delta, chi2_th);
    if (rc != ORB_OK) return rc;
    if (stop) {
        cudaEvent_t done = b->ev[1];
        if (!b->profile) ORB_CUDA(cudaEventRecord(done, b->stream));
        for (;;) {
            const cudaError_t q = cudaEventQuery(done);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) return orbhost::check_cuda(q, "cudaEventQuery", __FILE__, __LINE__);
            if (*stop) b->h_flags[0] = 1;
        }
    }
    orbba_stats_t st;
    rc = orbba_download(b, 0, poses_out, points_out, edge_outlier, &st);
    b->h_flags[0] = 0;
    if (rc != ORB_OK) return rc;
    if (stats) *stats = st;
    return st.status;
}

// Optimizer::BundleAdjustment / GlobalBundleAdjustemnt (src/Optimizer.cc:62-248) on one GPU: a single optimize(iterations),
// optional Huber kernel, no outlier pass.
int orbba_global(orbba_t* b, const orbba_problem_t* problem, int iterations, double huber_delta, const volatile uint8_t* stop,
                 double* poses_out, double* points_out, orbba_stats_t* stats) {
    return orbba_local(b, problem, iterations, -1, huber_delta, 1e300, stop, poses_out, points_out, nullptr, stats);
}

}  // extern "C"
