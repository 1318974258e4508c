// orb_gba.cu -- sm_100a GlobalBundleAdjustemnt for maps that do not fit the batched LocalBA path (thousands of key frames),
// on one GPU or landmark-partitioned over the GPUs of a node with one exchange step per LM trial.
//
// Reference path (file:line under /root/reference): Optimizer::GlobalBundleAdjustemnt / BundleAdjustment src/Optimizer.cc:62-248
// (all key frames and map points, one optimize(nIterations), Huber kernel when bRobust, no outlier pass) over the same g2o
// machinery as LocalBA (see orb_ba.cu for the per-block citations).
//
// Partitioning: key-frame poses are replicated, every rank owns the landmarks `point_id mod world == rank`
// together with their edges.  Landmark blocks are independent given the poses (Schur structure, block_solver.hpp:381-432), so a
// rank builds, for its landmarks only, Hll / bl, its share of Hpp / bp and its share of the reduced camera system
//     Hs = Hpp + lambda I - sum_l B_l Dinv_l B_l^T ,   bs = bp - sum_l B_l Dinv_l bl
// then ONE NCCL all-reduce (sum, FP64, over NVLink) of [Hs | bs] makes the system identical on every rank; the dense Cholesky
// (cuSOLVER potrf / potrs: a plain library factorisation, the only tensor-core-eligible part of the path) and the LM decision
// are replicated, the back-substitution and the trial errors are local again, and one all-reduce of two scalars closes the trial.
// NCCL and cuSOLVER are loaded with dlopen so that the library still loads on a box without them.
#include <dlfcn.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "orb_common.h"
#include "orb_ba_core.cuh"

#define G_T 128

struct GArgs {
    int nP, nL, nE, K, n, rank0_adds_bp;
    const int *e_pose, *e_pt, *e_cam, *pose_free, *pt_off;
    const double *e_obs, *e_info, *cam;
    double *pose[2], *pt[2], *err[2];
    double *rec, *B, *Y, *v, *Hll, *bl, *Hpp, *bp, *bs, *x, *Hs;
    double *part;            // per-block partial sums: [nb][2]
    double *red;             // [0] chi  [1] landmark scale  [2] max diag (landmarks)  [3] active edges  [4] pose scale  [5] max diag (poses)
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

// linearizeOplus + the per-edge part of constructQuadraticForm at estimate `cur` (errors already in err[cur])
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
    // pose block: Hpp_k += Jp^T W Jp, bp_k += Jp^T r  (this rank's share; summed over ranks by the all-reduce)
    const int k = A.pose_free[pi];
    if (k >= 0) {
        double* H = A.Hpp + 36 * (size_t)k;
        double* b = A.bp + 6 * (size_t)k;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            atomicAdd(&b[i], Jp[i] * R[7] + Jp[6 + i] * R[8]);
#pragma unroll
            for (int j = 0; j < 6; j++) atomicAdd(&H[i * 6 + j], (Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j]) * W);
        }
    }
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
        atomicMax(reinterpret_cast<unsigned long long*>(A.red + 2), (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like integers
    }
}
// max |diag Hpp| after the all-reduce (one CTA)
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

// thread per (edge, row): Y_e = B_e Dinv, v_e = Y_e bl ; bs_k -= v_e
__global__ void __launch_bounds__(256) g_trial(GArgs A, double lambda) {
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    if (q >= 6LL * A.nE) return;
    const int e = (int)(q / 6), r = (int)(q - 6LL * e);
    const int k = A.pose_free[A.e_pose[e]];
    if (k < 0) return;
    const int l = A.e_pt[e];
    double d[6];
    g_dinv(A.Hll + 6 * (size_t)l, lambda, d);
    const double* Br = A.B + 18 * (size_t)e + 3 * r;
    const double x0 = Br[0], x1 = Br[1], x2 = Br[2];
    const double y0 = x0 * d[0] + x1 * d[1] + x2 * d[2], y1 = x0 * d[1] + x1 * d[3] + x2 * d[4], y2 = x0 * d[2] + x1 * d[4] + x2 * d[5];
    double* Yr = A.Y + 18 * (size_t)e + 3 * r;
    Yr[0] = y0; Yr[1] = y1; Yr[2] = y2;
    const double* bl = A.bl + 3 * (size_t)l;
    atomicAdd(&A.bs[6 * (size_t)k + r], -(y0 * bl[0] + y1 * bl[1] + y2 * bl[2]));
}

// warp per landmark: Hs(block ki <= kj) -= Y_i B_j^T for every pair of its free-pose edges.  Hs is row-major n x n with the
// upper block triangle filled (= column-major lower triangle for cuSOLVER).
__global__ void __launch_bounds__(G_T) g_schur(GArgs A) {
    const int l = blockIdx.x * (G_T / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (l >= A.nL) return;
    const int e0 = A.pt_off[l], e1 = A.pt_off[l + 1];
    const size_t n = (size_t)A.n;
    for (int a = e0; a < e1; a++) {
        const int ka = A.pose_free[A.e_pose[a]];
        if (ka < 0) continue;
        for (int c = a; c < e1; c++) {
            const int kc = A.pose_free[A.e_pose[c]];
            if (kc < 0) continue;
            const int ei = ka <= kc ? a : c, ej = ka <= kc ? c : a;
            const int ki = min(ka, kc), kj = max(ka, kc);
            const double* Yi = A.Y + 18 * (size_t)ei;
            const double* Bj = A.B + 18 * (size_t)ej;
            for (int en = lane; en < 36; en += 32) {
                const int r = en / 6, cc = en - r * 6;
                const double s = Yi[r * 3] * Bj[cc * 3] + Yi[r * 3 + 1] * Bj[cc * 3 + 1] + Yi[r * 3 + 2] * Bj[cc * 3 + 2];
                atomicAdd(&A.Hs[(size_t)(6 * ki + r) * n + 6 * kj + cc], -s);
            }
        }
    }
}

// after the all-reduce: diagonal blocks += Hpp + lambda I
__global__ void __launch_bounds__(256) g_add_diag(GArgs A, double lambda) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 36 * A.K) return;
    const int k = i / 36, r = (i - 36 * k) / 6, c = i - 36 * k - 6 * r;
    A.Hs[(size_t)(6 * k + r) * A.n + 6 * k + c] += A.Hpp[i] + (r == c ? lambda : 0.0);
}

// trial poses: exp(x) * pose for free poses (x = bs after potrs), copy for fixed; pose part of computeScale() -> red[4]
__global__ void __launch_bounds__(256) g_pose_update(GArgs A, int cur, double lambda, int ok) {
    __shared__ double sm[256];
    double sc = 0;
    for (int i = threadIdx.x; i < A.nP; i += 256) {
        const int k = A.pose_free[i];
        const double* src = A.pose[cur] + 7 * (size_t)i;
        double* dst = A.pose[cur ^ 1] + 7 * (size_t)i;
        if (k >= 0 && ok) {
            const double* x = A.x + 6 * (size_t)k;
            se3_oplus(x, src, dst);
            for (int j = 0; j < 6; j++) sc += x[j] * (lambda * x[j] + A.bp[6 * (size_t)k + j]);
        } else {
            for (int q = 0; q < 7; q++) dst[q] = src[q];
        }
    }
    sm[threadIdx.x] = sc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < 256; i++) s += sm[i]; A.red[4] = s; }
}

// thread per landmark: increment, trial point, trial errors and chi2 of its edges, landmark part of computeScale()
__global__ void __launch_bounds__(G_T) g_back(GArgs A, int cur, double lambda, int ok, int robust, double delta) {
    __shared__ double red[G_T / 32];
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

// ================================================================================================ dynamic libraries
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
struct SolverApi {
    void* lib = nullptr;
    int (*Create)(void**) = nullptr;
    int (*Destroy)(void*) = nullptr;
    int (*SetStream)(void*, cudaStream_t) = nullptr;
    int (*PotrfBufferSize)(void*, int, int, double*, int, int*) = nullptr;
    int (*Potrf)(void*, int, int, double*, int, double*, int, int*) = nullptr;
    int (*Potrs)(void*, int, int, int, const double*, int, double*, int, int*) = nullptr;
};
NcclApi g_nccl;
SolverApi g_solver;

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
bool load_solver() {
    if (g_solver.lib) return true;
    for (const char* name : {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so"}) {
        void* h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (!h) continue;
        g_solver.Create = (int (*)(void**))dlsym(h, "cusolverDnCreate");
        g_solver.Destroy = (int (*)(void*))dlsym(h, "cusolverDnDestroy");
        g_solver.SetStream = (int (*)(void*, cudaStream_t))dlsym(h, "cusolverDnSetStream");
        g_solver.PotrfBufferSize = (int (*)(void*, int, int, double*, int, int*))dlsym(h, "cusolverDnDpotrf_bufferSize");
        g_solver.Potrf = (int (*)(void*, int, int, double*, int, double*, int, int*))dlsym(h, "cusolverDnDpotrf");
        g_solver.Potrs = (int (*)(void*, int, int, int, const double*, int, double*, int, int*))dlsym(h, "cusolverDnDpotrs");
        if (g_solver.Create && g_solver.Potrf && g_solver.Potrs && g_solver.PotrfBufferSize) { g_solver.lib = h; return true; }
        dlclose(h);
    }
    return false;
}
const int NCCL_F64 = 8, NCCL_SUM = 0, NCCL_MAX = 2, FILL_LOWER = 0;
}  // namespace

struct orbgba {
    int device = 0, rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    void* comm = nullptr;
    void* solver = nullptr;
    uint8_t* arena = nullptr; size_t arena_cap = 0;
    double* work = nullptr; int lwork = 0;
    int* d_info = nullptr;
    double* h_red = nullptr;        // pinned
    long long launches = 0;
    double allreduce_ms = 0, solve_ms = 0;     // accumulated over the last optimize call (events)
    size_t allreduce_bytes = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

static void gba_free(orbgba* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g->comm);
    if (g->solver && g_solver.Destroy) g_solver.Destroy(g->solver);
    cudaFree(g->arena); cudaFree(g->work); cudaFree(g->d_info);
    if (g->h_red) cudaFreeHost(g->h_red);
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
    if (!load_solver()) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_dist_create: libcusolver not found (dense reduced-camera solve)");
    if (world > 1 && !load_nccl()) ORB_FAIL(ORB_E_NO_DEVICE, "orbba_dist_create: libnccl.so.2 not found");
    ORB_CUDA(cudaSetDevice(device));
    orbgba* g = new (std::nothrow) orbgba();
    if (!g) ORB_FAIL(ORB_E_INVALID, "orbba_dist_create: out of host memory");
    g->device = device; g->rank = rank; g->world = world;
    cudaError_t ce = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&g->d_info, 16);
    if (ce == cudaSuccess) ce = cudaHostAlloc((void**)&g->h_red, 64 * sizeof(double), cudaHostAllocDefault);
    for (int i = 0; i < 4 && ce == cudaSuccess; i++) ce = cudaEventCreate(&g->ev[i]);
    if (ce != cudaSuccess) { int rc = orbhost::check_cuda(ce, "orbba_dist_create", __FILE__, __LINE__); gba_free(g); return rc; }
    if (g_solver.Create(&g->solver) != 0 || g_solver.SetStream(g->solver, g->stream) != 0) { gba_free(g); ORB_FAIL(ORB_E_CUDA, "orbba_dist_create: cusolverDnCreate failed"); }
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

// Optimizer::BundleAdjustment on this rank's shard: ALL poses (replicated, identical on every rank), this rank's landmarks
// (points [n_points][3]) and their edges (edge_point indexes the local landmark array).  Collective: every rank of the
// communicator must call it with the same poses / iterations / huber_delta.
int orbba_dist_optimize(orbgba_t* g, const orbba_problem_t* Q, int iterations, double huber_delta, const volatile uint8_t* stop,
                        double* poses_out, double* points_out, orbba_stats_t* stats) {
    if (!g || !Q) ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: bad argument");
    const int nP = Q->n_poses, nL = Q->n_points, nE = Q->n_edges, nC = Q->n_cams;
    if (nP < 1 || nL < 0 || nE < 0 || nC < 1 || iterations < 0) ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: bad sizes");
    if (!Q->poses || !Q->pose_fixed || (nL && !Q->points) || (nE && (!Q->edge_pose || !Q->edge_point || !Q->edge_cam || !Q->edge_obs || !Q->edge_inv_sigma2)) ||
        !Q->cam_K || !Q->cam_ext || !Q->cam_adj)
        ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: NULL array");
    ORB_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = g->stream;
    // ---- host-side preparation: free-pose numbering, edges grouped by landmark (stable), CSR
    std::vector<int> pose_free(nP, -1);
    int K = 0;
    for (int i = 0; i < nP; i++) if (!Q->pose_fixed[i]) pose_free[i] = K++;
    const long long n = 6LL * K;
    if (n > 46000) ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: %d free poses exceed the dense reduced-camera solver", K);
    std::vector<int> cnt(nL + 1, 0), perm(nE);
    for (int e = 0; e < nE; e++) {
        if (Q->edge_pose[e] < 0 || Q->edge_pose[e] >= nP || Q->edge_point[e] < 0 || Q->edge_point[e] >= nL || Q->edge_cam[e] < 0 || Q->edge_cam[e] >= nC)
            ORB_FAIL(ORB_E_INVALID, "orbba_dist_optimize: edge %d indexes out of range", e);
        cnt[Q->edge_point[e] + 1]++;
    }
    for (int l = 0; l < nL; l++) cnt[l + 1] += cnt[l];
    std::vector<int> pt_off(cnt);
    for (int e = 0; e < nE; e++) perm[cnt[Q->edge_point[e]]++] = e;
    // ---- layout
    size_t cur_off = 0;
    auto add = [&](size_t b) { const size_t o = (cur_off + 255) & ~(size_t)255; cur_off = o + b; return o; };
    const size_t o_epose = add(4 * (size_t)nE), o_ept = add(4 * (size_t)nE), o_ecam = add(4 * (size_t)nE), o_pfree = add(4 * (size_t)nP), o_ptoff = add(4 * (size_t)(nL + 1));
    const size_t o_eobs = add(16 * (size_t)nE), o_einfo = add(8 * (size_t)nE), o_cam = add(8 * BA_CAM_STRIDE * (size_t)nC);
    const size_t o_pose0 = add(56 * (size_t)nP), o_pt0 = add(24 * (size_t)nL);
    const size_t staged = add(0);
    const size_t o_pose1 = add(56 * (size_t)nP), o_pt1 = add(24 * (size_t)nL), o_err0 = add(16 * (size_t)nE), o_err1 = add(16 * (size_t)nE);
    const size_t o_rec = add(8 * BA_REC * (size_t)nE), o_B = add(144 * (size_t)nE), o_Y = add(144 * (size_t)nE);
    const size_t o_Hll = add(48 * (size_t)nL), o_bl = add(24 * (size_t)nL);
    const size_t o_Hpp = add(8 * (size_t)(36 + 6) * std::max(K, 1));                 // Hpp | bp contiguous: one all-reduce
    const size_t o_Hs = add(8 * (size_t)(n * n + n + 8));                            // Hs | bs contiguous: one all-reduce
    const size_t o_x = add(8 * (size_t)std::max<long long>(n, 1));
    const int nbE = std::max(1, (nE + G_T - 1) / G_T), nbL = std::max(1, (nL + G_T - 1) / G_T);
    const size_t o_part = add(16 * (size_t)std::max(nbE, nbL)), o_red = add(64 * 8);
    const size_t o_pout = add(96 * (size_t)nP), o_lout = add(24 * (size_t)std::max(nL, 1));
    const size_t total = add(0) + 256;
    if (total > g->arena_cap) {
        cudaFree(g->arena); g->arena = nullptr; g->arena_cap = 0;
        ORB_CUDA(cudaMalloc((void**)&g->arena, total));
        g->arena_cap = total;
    }
    std::vector<uint8_t> H(staged);
    int *h_epose = (int*)(H.data() + o_epose), *h_ept = (int*)(H.data() + o_ept), *h_ecam = (int*)(H.data() + o_ecam);
    double *h_eobs = (double*)(H.data() + o_eobs), *h_einfo = (double*)(H.data() + o_einfo), *h_cam = (double*)(H.data() + o_cam), *h_pose0 = (double*)(H.data() + o_pose0);
    for (int s = 0; s < nE; s++) {
        const int e = perm[s];
        h_epose[s] = Q->edge_pose[e]; h_ept[s] = Q->edge_point[e]; h_ecam[s] = Q->edge_cam[e];
        h_eobs[2 * s] = Q->edge_obs[2 * e]; h_eobs[2 * s + 1] = Q->edge_obs[2 * e + 1]; h_einfo[s] = Q->edge_inv_sigma2[e];
    }
    memcpy(H.data() + o_pfree, pose_free.data(), 4 * (size_t)nP);
    memcpy(H.data() + o_ptoff, pt_off.data(), 4 * (size_t)(nL + 1));
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
    uint8_t* D = g->arena;
    ORB_CUDA(cudaMemcpyAsync(D, H.data(), staged, cudaMemcpyHostToDevice, st));
    GArgs A;
    memset(&A, 0, sizeof(A));
    A.nP = nP; A.nL = nL; A.nE = nE; A.K = K; A.n = (int)n; A.rank0_adds_bp = g->rank == 0;
    A.e_pose = (const int*)(D + o_epose); A.e_pt = (const int*)(D + o_ept); A.e_cam = (const int*)(D + o_ecam); A.pose_free = (const int*)(D + o_pfree);
    A.pt_off = (const int*)(D + o_ptoff); A.e_obs = (const double*)(D + o_eobs); A.e_info = (const double*)(D + o_einfo); A.cam = (const double*)(D + o_cam);
    A.pose[0] = (double*)(D + o_pose0); A.pose[1] = (double*)(D + o_pose1); A.pt[0] = (double*)(D + o_pt0); A.pt[1] = (double*)(D + o_pt1);
    A.err[0] = (double*)(D + o_err0); A.err[1] = (double*)(D + o_err1);
    A.rec = (double*)(D + o_rec); A.B = (double*)(D + o_B); A.Y = (double*)(D + o_Y); A.v = nullptr;
    A.Hll = (double*)(D + o_Hll); A.bl = (double*)(D + o_bl); A.Hpp = (double*)(D + o_Hpp); A.bp = A.Hpp + 36 * (size_t)std::max(K, 1);
    A.Hs = (double*)(D + o_Hs); A.bs = A.Hs + n * n; A.x = (double*)(D + o_x);
    A.part = (double*)(D + o_part); A.red = (double*)(D + o_red);
    double* d_pout = (double*)(D + o_pout);
    double* d_lout = (double*)(D + o_lout);
    // cuSOLVER workspace
    if (n > 0) {
        int lw = 0;
        if (g_solver.PotrfBufferSize(g->solver, FILL_LOWER, (int)n, A.Hs, (int)n, &lw) != 0) ORB_FAIL(ORB_E_CUDA, "cusolverDnDpotrf_bufferSize failed");
        if (lw > g->lwork) { cudaFree(g->work); g->work = nullptr; g->lwork = 0; ORB_CUDA(cudaMalloc((void**)&g->work, sizeof(double) * (size_t)lw)); g->lwork = lw; }
    }
    g->allreduce_ms = 0; g->solve_ms = 0; g->allreduce_bytes = 0;
    const bool robust = huber_delta > 0;
    auto read_red = [&](int count) -> int {   // device red[] -> host
        ORB_CUDA(cudaMemcpyAsync(g->h_red, A.red, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        return ORB_OK;
    };
    int rc;
    orbba_stats_t S;
    memset(&S, 0, sizeof(S));
    int cur = 0;
    const bool stopped0 = stop && *stop;
    // ---- initial errors: chi2 and the global number of edges
    ORB_CUDA(cudaMemsetAsync(A.red, 0, 64 * 8, st));
    g_errors<<<nbE, G_T, 0, st>>>(A, cur, robust, huber_delta);
    g_reduce<<<1, 256, 0, st>>>(A, nbE, 0, -1);
    g->launches += 2;
    g->h_red[0] = 0;
    {
        // red[3] = local edge count, summed over ranks
        const double ne = (double)nE;
        ORB_CUDA(cudaMemcpyAsync(A.red + 3, &ne, 8, cudaMemcpyHostToDevice, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    if ((rc = all_reduce(g, A.red, 4, NCCL_SUM)) != ORB_OK) return rc;
    if ((rc = read_red(4)) != ORB_OK) return rc;
    double currentChi = g->h_red[0];
    const double totalEdges = g->h_red[3];
    S.initial_chi2 = currentChi;
    double lambda = 0, ni = 2, rho = 0;
    int nBad = 0;
    bool ok_iter = !stopped0 && totalEdges > 0;
    for (int it = 0; it < iterations && ok_iter; it++) {
        if (stop && *stop) break;
        // ---- buildSystem
        ORB_CUDA(cudaMemsetAsync(A.Hpp, 0, 8 * (size_t)(36 + 6) * std::max(K, 1), st));
        ORB_CUDA(cudaMemsetAsync(A.red + 2, 0, 8, st));
        g_lin<<<nbE, G_T, 0, st>>>(A, cur, robust, huber_delta);
        g_build_lm<<<nbL, G_T, 0, st>>>(A);
        g->launches += 2;
        ORB_CUDA(cudaEventRecord(g->ev[0], st));
        if ((rc = all_reduce(g, A.Hpp, (size_t)(36 + 6) * K, NCCL_SUM)) != ORB_OK) return rc;
        if (it == 0) {
            if ((rc = all_reduce(g, A.red + 2, 1, NCCL_MAX)) != ORB_OK) return rc;   // non-negative doubles: max is order-free
            g_pose_maxdiag<<<1, 256, 0, st>>>(A);
            g->launches++;
            if ((rc = read_red(6)) != ORB_OK) return rc;
            lambda = 1e-5 * std::max(g->h_red[2], g->h_red[5]);     // computeLambdaInit
            ni = 2; nBad = 0;
        }
        const double iniChi = currentChi;
        int qmax = 0;
        rho = 0;
        bool again = true;
        while (again) {
            // ---- setLambda + Schur complement (this rank's landmarks)
            ORB_CUDA(cudaMemsetAsync(A.Hs, 0, 8 * (size_t)(n * n + n), st));
            if (g->rank == 0 && n > 0) ORB_CUDA(cudaMemcpyAsync(A.bs, A.bp, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
            if (nE > 0) {
                g_trial<<<(unsigned)((6LL * nE + 255) / 256), 256, 0, st>>>(A, lambda);
                g_schur<<<(nL + G_T / 32 - 1) / (G_T / 32), G_T, 0, st>>>(A);
                g->launches += 2;
            }
            // ---- the exchange step: [Hs | bs] summed over ranks
            ORB_CUDA(cudaEventRecord(g->ev[0], st));
            if ((rc = all_reduce(g, A.Hs, (size_t)(n * n + n), NCCL_SUM)) != ORB_OK) return rc;
            ORB_CUDA(cudaEventRecord(g->ev[1], st));
            int info = 0;
            if (n > 0) {
                g_add_diag<<<(36 * K + 255) / 256, 256, 0, st>>>(A, lambda);
                g->launches++;
                if (g_solver.Potrf(g->solver, FILL_LOWER, (int)n, A.Hs, (int)n, g->work, g->lwork, g->d_info) != 0) ORB_FAIL(ORB_E_CUDA, "cusolverDnDpotrf failed");
                ORB_CUDA(cudaMemcpyAsync(&info, g->d_info, 4, cudaMemcpyDeviceToHost, st));
                ORB_CUDA(cudaStreamSynchronize(st));
                if (info == 0) {
                    if (g_solver.Potrs(g->solver, FILL_LOWER, (int)n, 1, A.Hs, (int)n, A.bs, (int)n, g->d_info) != 0) ORB_FAIL(ORB_E_CUDA, "cusolverDnDpotrs failed");
                    ORB_CUDA(cudaMemcpyAsync(A.x, A.bs, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
                }
            }
            ORB_CUDA(cudaEventRecord(g->ev[2], st));
            const int ok = info == 0;
            // ---- update + trial errors (local), then the two scalars
            g_pose_update<<<1, 256, 0, st>>>(A, cur, lambda, ok);
            g_back<<<nbL, G_T, 0, st>>>(A, cur, lambda, ok, robust, huber_delta);
            g_reduce<<<1, 256, 0, st>>>(A, nbL, 0, 1);
            g->launches += 3;
            if ((rc = all_reduce(g, A.red, 2, NCCL_SUM)) != ORB_OK) return rc;
            if ((rc = read_red(5)) != ORB_OK) return rc;
            {
                float ms = 0;
                cudaEventElapsedTime(&ms, g->ev[0], g->ev[1]); g->allreduce_ms += ms;
                cudaEventElapsedTime(&ms, g->ev[1], g->ev[2]); g->solve_ms += ms;
            }
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
            const bool stopped = stop && *stop;
            again = rho < 0 && qmax < 10 && !stopped;
        }
        S.iterations++;
        if (qmax == 10 || rho == 0) ok_iter = false;
        else {
            if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
            if (nBad >= 3) ok_iter = false;
        }
    }
    g_outputs<<<std::max(1, (std::max(nP, 3 * nL / 8 + 1) + 255) / 256), 256, 0, st>>>(A, cur, d_pout, d_lout);
    g->launches++;
    ORB_CUDA(cudaGetLastError());
    if (poses_out) ORB_CUDA(cudaMemcpyAsync(poses_out, d_pout, 96 * (size_t)nP, cudaMemcpyDeviceToHost, st));
    if (points_out && nL) ORB_CUDA(cudaMemcpyAsync(points_out, d_lout, 24 * (size_t)nL, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    S.final_chi2 = currentChi; S.final_lambda = lambda; S.outliers = 0; S.status = stopped0 ? ORB_E_ABORTED : ORB_OK;
    if (stats) *stats = S;
    return S.status;
}

}  // extern "C"
