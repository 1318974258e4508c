// orb_pose.cu -- sm_100a Optimizer::PoseOptimization behind orbba_pose_optimization.
//
// Reference path (file:line under /root/reference): Optimizer::PoseOptimization src/Optimizer.cc:250-405 -- one VertexSE3Expmap,
// one unary EdgeSE3ProjectXYZOnlyPose per keypoint that holds a map point (Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:200-255),
// Huber(sqrt(5.991)), BlockSolver_6_3 over LinearSolverDense, four rounds of optimize(10) that each restart from Frame::mTcw, with an
// inlier / outlier re-classification (chi2 > 5.991) after every round and the robust kernel dropped after the third.
// It runs two to three times per frame between the matcher calls (src/Tracking.cc:1321,1427,1488).
//
// B200 formulation: a batch of frames (one per tracked sequence), ONE persistent CTA per frame runs all four rounds -- every LM
// iteration and trial, the 6x6 solve, the exp-map update, the lambda policy and the re-classification -- without returning to the
// host.  FP64, fixed-order block reductions (bit-reproducible).  No CPU fallback.
#include <string.h>

#include <vector>

#include "orb_common.h"
#include "orb_ba_core.cuh"

#define PO_T 256
#define PO_W (PO_T / 32)
#define PO_CAMS 4

struct POFrame { int e0, nE, c0, nC; };
struct POArgs {
    int n;
    const POFrame* fr;
    const double *pose0, *Xw, *obs, *info, *cams;
    const int* cam;
    double* err;
    unsigned char *level, *outlier;
    double* pose_out;
    int *inliers, *counts;
};

struct orbba;
cudaStream_t orbba_stream_of(orbba*);
int orbba_device_of(orbba*);
void orbba_count_launches(orbba*, int n);

__device__ __forceinline__ double po_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ double po_block_sum(double v, double* red) {
    v = po_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < PO_W; w++) s += red[w];
    return s;
}

struct POShared {
    double pose[7], init[7], trial[7];
    double cam[PO_CAMS * BA_CAM_STRIDE];
    double red[PO_W], red27[PO_W * 27];
    double H[36], b[6], x[6];
    double lambda, ni, currentChi, iniChi, rho;
    int nBad, qmax, ok, go, accepted, iterations, trials;
};

// computeActiveErrors + activeRobustChi2 at `pose` (errors of level-0 edges are stored: they are what e->chi2() reads later)
__device__ double po_errors(const POArgs& A, const POFrame& F, const POShared& S, const double* pose, const double* cams, bool robust, double delta,
                            double* red) {
    const double dsqr = delta * delta;
    double chi = 0;
    for (int e = F.e0 + threadIdx.x; e < F.e0 + F.nE; e += PO_T) {
        if (A.level[e]) continue;
        const double* c = cams + BA_CAM_STRIDE * (A.cam[e] - F.c0);
        double pc[3], er[2];
        edge_project(pose, A.Xw + 3 * (size_t)e, c, pc);
        edge_error(pc, c, A.obs + 2 * (size_t)e, er);
        A.err[2 * (size_t)e] = er[0]; A.err[2 * (size_t)e + 1] = er[1];
        const double c2 = (er[0] * er[0] + er[1] * er[1]) * A.info[e];
        chi += robust ? huber_rho0(c2, delta, dsqr) : c2;
    }
    return po_block_sum(chi, red);
}

__global__ void __launch_bounds__(PO_T) k_pose_opt(POArgs A) {
    __shared__ POShared S;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const POFrame F = A.fr[f];
    const bool cam_sm = F.nC <= PO_CAMS;
    if (cam_sm) for (int i = tid; i < F.nC * BA_CAM_STRIDE; i += PO_T) S.cam[i] = A.cams[(size_t)BA_CAM_STRIDE * F.c0 + i];
    if (tid < 7) { S.init[tid] = A.pose0[7 * (size_t)f + tid]; S.pose[tid] = S.init[tid]; }
    if (tid == 0) { S.iterations = 0; S.trials = 0; }
    for (int e = F.e0 + tid; e < F.e0 + F.nE; e += PO_T) { A.level[e] = 0; A.outlier[e] = 0; A.err[2 * (size_t)e] = 0; A.err[2 * (size_t)e + 1] = 0; }
    __syncthreads();
    const double* cams = cam_sm ? S.cam : A.cams + (size_t)BA_CAM_STRIDE * F.c0;
    const double delta = (double)sqrtf(5.991f);            // const float deltaMono = sqrt(5.991)
    int nBad = 0;
    if (F.nE >= 3) {
        for (int round = 0; round < 4; round++) {
            const bool robust = round < 3;                  // setRobustKernel(0) after the third round (it == 2)
            if (tid < 7) S.pose[tid] = S.init[tid];         // every round restarts from Frame::mTcw
            double nact = 0;
            for (int e = F.e0 + tid; e < F.e0 + F.nE; e += PO_T) nact += A.level[e] == 0;
            nact = po_block_sum(nact, S.red);
            if (nact > 0) {
                bool ok_iter = true;
                for (int it = 0; it < 10 && ok_iter; it++) {
                    const double chi0 = po_errors(A, F, S, S.pose, cams, robust, delta, S.red);
                    // ---- buildSystem: H (21 unique) and b (6) over the active edges
                    double h[27];
#pragma unroll
                    for (int i = 0; i < 27; i++) h[i] = 0;
                    const double dsqr = delta * delta;
                    for (int e = F.e0 + tid; e < F.e0 + F.nE; e += PO_T) {
                        if (A.level[e]) continue;
                        const double* c = cams + BA_CAM_STRIDE * (A.cam[e] - F.c0);
                        double pc[3];
                        edge_project(S.pose, A.Xw + 3 * (size_t)e, c, pc);
                        const double X = pc[0], Y = pc[1], Z = pc[2], iz = -1. / Z;
                        const double t00 = iz * c[0], t02 = iz * (-X / Z * c[0]), t11 = iz * c[1], t12 = iz * (-Y / Z * c[1]);
                        const double tJ[12] = {t02 * Y, t00 * Z - t02 * X, -t00 * Y, t00, 0, t02, -t11 * Z + t12 * Y, -t12 * X, t11 * X, 0, t11, t12};
                        const double* Ad = c + BA_CAM_ADJ;
                        double Jp[12];
#pragma unroll
                        for (int i = 0; i < 2; i++)
#pragma unroll
                            for (int j = 0; j < 6; j++) {
                                double s = 0;
#pragma unroll
                                for (int k = 0; k < 6; k++) s += tJ[i * 6 + k] * Ad[k * 6 + j];
                                Jp[i * 6 + j] = s;
                            }
                        const double w = A.info[e], e0 = A.err[2 * (size_t)e], e1 = A.err[2 * (size_t)e + 1];
                        double wr = 1.0;
                        if (robust) { const double c2 = (e0 * e0 + e1 * e1) * w; if (c2 > dsqr) wr = delta / sqrt(c2); }
                        const double W = wr * w, r0 = -w * e0 * wr, r1 = -w * e1 * wr;
                        int u = 0;
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            h[21 + i] += Jp[i] * r0 + Jp[6 + i] * r1;
#pragma unroll
                            for (int j = i; j < 6; j++) h[u++] += (Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j]) * W;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 27; i++) { const double v = po_warp_sum(h[i]); if (lane == 0) S.red27[warp * 27 + i] = v; }
                    __syncthreads();
                    if (tid < 27) {
                        double s = 0;
                        for (int w = 0; w < PO_W; w++) s += S.red27[w * 27 + tid];
                        if (tid >= 21) S.b[tid - 21] = s;
                        else {
                            int i = 0, rem = tid;
                            while (rem >= 6 - i) { rem -= 6 - i; i++; }
                            const int j = i + rem;
                            S.H[i * 6 + j] = s; S.H[j * 6 + i] = s;
                        }
                    }
                    __syncthreads();
                    if (tid == 0) {
                        if (it == 0) {                      // computeLambdaInit
                            double md = 0;
                            for (int j = 0; j < 6; j++) md = fmax(md, fabs(S.H[j * 7]));
                            S.lambda = 1e-5 * md; S.ni = 2; S.nBad = 0;
                        }
                        S.currentChi = chi0; S.iniChi = chi0; S.qmax = 0; S.rho = 0;
                    }
                    __syncthreads();
                    bool again = true;
                    while (again) {
                        if (tid == 0) {                     // (Hpp + lambda I) x = b : LinearSolverDense, LDL^T without pivoting
                            double M[36], x[6];
                            for (int i = 0; i < 36; i++) M[i] = S.H[i] + ((i % 7 == 0) ? S.lambda : 0.0);
                            bool ok = true;
                            for (int j = 0; j < 6 && ok; j++) {
                                const double dj = M[j * 6 + j];
                                if (dj == 0.0 || !isfinite(dj)) { ok = false; break; }
                                for (int i = j + 1; i < 6; i++) M[i * 6 + j] /= dj;
                                for (int i = j + 1; i < 6; i++)
                                    for (int k = j + 1; k <= i; k++) M[i * 6 + k] -= M[i * 6 + j] * dj * M[k * 6 + j];
                            }
                            for (int i = 0; i < 6; i++) x[i] = S.b[i];
                            if (ok) {
                                for (int j = 0; j < 6; j++) for (int i = j + 1; i < 6; i++) x[i] -= M[i * 6 + j] * x[j];
                                for (int i = 0; i < 6; i++) x[i] /= M[i * 6 + i];
                                for (int j = 5; j >= 0; j--) for (int i = 0; i < j; i++) x[i] -= M[j * 6 + i] * x[j];
                            } else {
                                for (int i = 0; i < 6; i++) x[i] = 0;
                            }
                            for (int i = 0; i < 6; i++) S.x[i] = x[i];
                            S.ok = ok;
                            se3_oplus(x, S.pose, S.trial);
                        }
                        __syncthreads();
                        double tempChi = po_errors(A, F, S, S.trial, cams, robust, delta, S.red);
                        if (tid == 0) {
                            if (!S.ok) tempChi = 1.7976931348623157e308;
                            double scale = 0;
                            for (int j = 0; j < 6; j++) scale += S.x[j] * (S.lambda * S.x[j] + S.b[j]);
                            scale += 1e-3;
                            const double rho = (S.currentChi - tempChi) / scale;
                            S.rho = rho;
                            S.trials++;
                            if (rho > 0 && isfinite(tempChi)) {
                                double alpha = 1. - pow(2 * rho - 1, 3.0);
                                alpha = fmin(alpha, 2. / 3.);
                                S.lambda *= fmax(1. / 3., alpha);
                                S.ni = 2;
                                S.currentChi = tempChi;
                                for (int q = 0; q < 7; q++) S.pose[q] = S.trial[q];
                            } else {
                                S.lambda *= S.ni;
                                S.ni *= 2;
                            }
                            S.qmax++;
                            S.go = rho < 0 && S.qmax < 10;
                        }
                        __syncthreads();
                        again = S.go != 0;
                    }
                    if (tid == 0) {
                        S.iterations++;
                        int res = 0;
                        if (S.qmax == 10 || S.rho == 0) res = 1;
                        else {
                            if ((S.iniChi - S.currentChi) * 1e3 < S.iniChi) S.nBad++; else S.nBad = 0;
                            if (S.nBad >= 3) res = 1;
                        }
                        S.go = res == 0;
                    }
                    __syncthreads();
                    ok_iter = S.go != 0;
                    __syncthreads();
                }
            }
            // ---- classification (src/Optimizer.cc:366-391): outliers are re-evaluated at the optimised pose, inliers keep the error
            //      of the last computeActiveErrors; chi2 is compared in float like the reference
            double bad = 0;
            for (int e = F.e0 + tid; e < F.e0 + F.nE; e += PO_T) {
                double e0 = A.err[2 * (size_t)e], e1 = A.err[2 * (size_t)e + 1];
                if (A.outlier[e]) {
                    const double* c = cams + BA_CAM_STRIDE * (A.cam[e] - F.c0);
                    double pc[3], er[2];
                    edge_project(S.pose, A.Xw + 3 * (size_t)e, c, pc);
                    edge_error(pc, c, A.obs + 2 * (size_t)e, er);
                    e0 = er[0]; e1 = er[1];
                    A.err[2 * (size_t)e] = e0; A.err[2 * (size_t)e + 1] = e1;
                }
                const float chi2 = (float)((e0 * e0 + e1 * e1) * A.info[e]);
                const bool out = chi2 > 5.991f;
                A.outlier[e] = out; A.level[e] = out;
                bad += out;
            }
            nBad = (int)po_block_sum(bad, S.red);
            __syncthreads();
            if (F.nE < 10) break;                           // optimizer.edges().size() < 10
        }
    }
    if (tid == 0) {
        double R[9];
        q_to_matrix(S.pose, R);
        double* o = A.pose_out + 12 * (size_t)f;
        for (int r = 0; r < 3; r++) { o[r * 4] = R[r * 3]; o[r * 4 + 1] = R[r * 3 + 1]; o[r * 4 + 2] = R[r * 3 + 2]; o[r * 4 + 3] = S.pose[4 + r]; }
        A.inliers[f] = F.nE >= 3 ? F.nE - nBad : 0;
        A.counts[2 * f] = S.iterations; A.counts[2 * f + 1] = S.trials;
    }
}

namespace {
struct PoArena {
    uint8_t* d = nullptr; size_t cap = 0; int device = -1;
    void release() {
        if (d && device >= 0) {
            int prev = -1;
            cudaGetDevice(&prev);
            if (cudaSetDevice(device) == cudaSuccess) cudaFree(d);
            if (prev >= 0) cudaSetDevice(prev);
            cudaGetLastError();
        }
        d = nullptr; cap = 0;
    }
    ~PoArena() { release(); }  // thread exit
};
thread_local PoArena g_po;
}  // namespace

extern "C" {

int orbba_pose_optimization(orbba_t* h, const orbpo_frame_t* frames, int n, double* poses_out, uint8_t* outlier, int32_t* n_inliers, int32_t* lm_counts) {
    if (!h) ORB_FAIL(ORB_E_INVALID, "orbba_pose_optimization: NULL handle");
    if (n == 0) return ORB_OK;
    if (!frames || n < 0 || !poses_out || !outlier || !n_inliers) ORB_FAIL(ORB_E_INVALID, "orbba_pose_optimization: bad argument");
    long long Etot = 0, Ctot = 0;
    std::vector<POFrame> fr((size_t)n);
    for (int f = 0; f < n; f++) {
        const orbpo_frame_t& Q = frames[f];
        if (!Q.pose || Q.n_obs < 0 || Q.n_cams < 1 || !Q.cam_K || !Q.cam_ext || !Q.cam_adj || (Q.n_obs && (!Q.Xw || !Q.obs || !Q.inv_sigma2 || !Q.cam)))
            ORB_FAIL(ORB_E_INVALID, "orbba_pose_optimization: frame %d has a NULL field or negative size", f);
        for (int e = 0; e < Q.n_obs; e++)
            if (Q.cam[e] < 0 || Q.cam[e] >= Q.n_cams) ORB_FAIL(ORB_E_INVALID, "orbba_pose_optimization: frame %d observation %d names camera %d", f, e, Q.cam[e]);
        fr[f].e0 = (int)Etot; fr[f].nE = Q.n_obs; fr[f].c0 = (int)Ctot; fr[f].nC = Q.n_cams;
        Etot += Q.n_obs; Ctot += Q.n_cams;
        if (Etot > 0x7fffffffLL) ORB_FAIL(ORB_E_INVALID, "orbba_pose_optimization: batch too large");
    }
    const int device = orbba_device_of(h);
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = orbba_stream_of(h);
    size_t cur = 0;
    auto add = [&](size_t b) { const size_t o = (cur + 255) & ~(size_t)255; cur = o + b; return o; };
    const size_t o_fr = add(sizeof(POFrame) * n), o_pose = add(56 * (size_t)n), o_Xw = add(24 * (size_t)Etot), o_obs = add(16 * (size_t)Etot), o_info = add(8 * (size_t)Etot),
                 o_cam = add(4 * (size_t)Etot), o_cams = add(8 * BA_CAM_STRIDE * (size_t)Ctot);
    const size_t staged = add(0);
    const size_t o_err = add(16 * (size_t)Etot), o_level = add((size_t)Etot), o_out = add((size_t)Etot), o_pout = add(96 * (size_t)n), o_inl = add(4 * (size_t)n), o_cnt = add(8 * (size_t)n);
    const size_t total = add(0) + 256;
    PoArena& G = g_po;
    if (G.device != device || total > G.cap) {
        G.release();
        G.device = device;
        ORB_CUDA(cudaMalloc((void**)&G.d, total + total / 2));
        G.cap = total + total / 2;
    }
    std::vector<uint8_t> H(staged);
    memcpy(H.data() + o_fr, fr.data(), sizeof(POFrame) * n);
    double* hp = (double*)(H.data() + o_pose);
    double* hc = (double*)(H.data() + o_cams);
    for (int f = 0; f < n; f++) {
        const orbpo_frame_t& Q = frames[f];
        const double* T = Q.pose;
        double* Dp = hp + 7 * (size_t)f;
        const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
        q_from_matrix(R, Dp);                               // Converter::toSE3Quat + SE3Quat(R, t)
        if (Dp[3] < 0) for (int k = 0; k < 4; k++) Dp[k] = -Dp[k];
        const double nn = sqrt(Dp[0] * Dp[0] + Dp[1] * Dp[1] + Dp[2] * Dp[2] + Dp[3] * Dp[3]);
        for (int k = 0; k < 4; k++) Dp[k] /= nn;
        Dp[4] = T[3]; Dp[5] = T[7]; Dp[6] = T[11];
        const size_t e0 = (size_t)fr[f].e0;
        if (Q.n_obs) {
            memcpy(H.data() + o_Xw + 24 * e0, Q.Xw, 24 * (size_t)Q.n_obs);
            memcpy(H.data() + o_obs + 16 * e0, Q.obs, 16 * (size_t)Q.n_obs);
            memcpy(H.data() + o_info + 8 * e0, Q.inv_sigma2, 8 * (size_t)Q.n_obs);
            int* ci = (int*)(H.data() + o_cam) + e0;
            for (int e = 0; e < Q.n_obs; e++) ci[e] = fr[f].c0 + Q.cam[e];
        }
        for (int c = 0; c < Q.n_cams; c++) {
            double* Dc = hc + (size_t)BA_CAM_STRIDE * (fr[f].c0 + c);
            memset(Dc, 0, sizeof(double) * BA_CAM_STRIDE);
            for (int i = 0; i < 4; i++) Dc[i] = Q.cam_K[4 * c + i];
            const double* E = Q.cam_ext + 12 * c;
            const double Rc[9] = {E[0], E[1], E[2], E[4], E[5], E[6], E[8], E[9], E[10]};
            q_from_matrix(Rc, Dc + 4);
            if (Dc[7] < 0) for (int i = 4; i < 8; i++) Dc[i] = -Dc[i];
            const double nc = sqrt(Dc[4] * Dc[4] + Dc[5] * Dc[5] + Dc[6] * Dc[6] + Dc[7] * Dc[7]);
            for (int i = 4; i < 8; i++) Dc[i] /= nc;
            Dc[8] = E[3]; Dc[9] = E[7]; Dc[10] = E[11];
            for (int i = 0; i < 36; i++) Dc[BA_CAM_ADJ + i] = Q.cam_adj[36 * c + i];
        }
    }
    uint8_t* D = G.d;
    ORB_CUDA(cudaMemcpyAsync(D, H.data(), staged, cudaMemcpyHostToDevice, st));
    POArgs A;
    A.n = n; A.fr = (const POFrame*)(D + o_fr); A.pose0 = (const double*)(D + o_pose); A.Xw = (const double*)(D + o_Xw); A.obs = (const double*)(D + o_obs);
    A.info = (const double*)(D + o_info); A.cams = (const double*)(D + o_cams); A.cam = (const int*)(D + o_cam);
    A.err = (double*)(D + o_err); A.level = D + o_level; A.outlier = D + o_out; A.pose_out = (double*)(D + o_pout); A.inliers = (int*)(D + o_inl); A.counts = (int*)(D + o_cnt);
    k_pose_opt<<<n, PO_T, 0, st>>>(A);
    ORB_CUDA(cudaGetLastError());
    orbba_count_launches(h, 1);
    ORB_CUDA(cudaMemcpyAsync(poses_out, D + o_pout, 96 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (Etot) ORB_CUDA(cudaMemcpyAsync(outlier, D + o_out, (size_t)Etot, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(n_inliers, D + o_inl, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (lm_counts) ORB_CUDA(cudaMemcpyAsync(lm_counts, D + o_cnt, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

}  // extern "C"
