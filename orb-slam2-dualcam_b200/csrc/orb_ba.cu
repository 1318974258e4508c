// orb_ba.cu -- sm_100a Levenberg-Marquardt bundle adjustment behind the orbba_* C-ABI.
//
// Reference path (file:line under /root/reference):
//   Optimizer::LocalBundleAdjustment src/Optimizer.cc:407-696, Optimizer::BundleAdjustment :70-248
//   EdgeSE3ProjectXYZ::{computeError,linearizeOplus,isDepthPositive} Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:109-169
//   BaseBinaryEdge::constructQuadraticForm .../core/base_binary_edge.hpp:55-120, RobustKernelHuber .../core/robust_kernel_impl.cpp:78-91
//   BlockSolver<6,3>::{buildSystem,setLambda,solve} .../core/block_solver.hpp:354-604 (Schur complement)
//   OptimizationAlgorithmLevenberg::solve .../core/optimization_algorithm_levenberg.cpp:61-189, SE3Quat .../types/se3quat.h
//
// B200 formulation: ONE persistent CTA runs the whole optimisation of one problem -- both LM rounds, every trial, the
// lambda policy and the outlier re-classification -- without returning to the host; a batch of independent problems
// (one per keyframe / per sequence) fills the 148 SMs.  Everything is FP64.  No atomics: every sum has one owner
//   per edge      residual, 2x6 / 2x3 Jacobians, Huber weight                       (thread per edge)
//   per landmark  Hll, bl, (Hll + lambda I)^-1, back-substitution                   (thread per landmark, CSR by landmark)
//   per pose      Hpp, bp, Schur right-hand side                                    (warp per pose, CSR by pose)
//   per pose pair 6x6 block of the reduced camera matrix: Hpp - sum_l B_il Dinv_l B_jl^T  (warp per pair, list of (edge,edge) tuples)
// so results are bit-reproducible run to run.  The reduced camera system (6K x 6K, K free poses) lives in shared
// memory (K <= 26) and is factorised in place by LDL^T.
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "orb_common.h"

#define BA_T 512
#define BA_WARPS (BA_T / 32)
#define BA_JSTRIDE 21          // per-edge record: Jp[12] Jl[6] W r0 r1 (SoA: J[k * nE + e])
#define BA_CAM_STRIDE 47       // fx fy cx cy | ext quat xyzw | ext t | adj[36]

struct BAProb {
    int nP, nL, nE, nC, K, n, nPairs, nTuples;
    const int *e_pose, *e_pt, *e_cam, *pose_free;
    const double *e_obs, *e_info, *cam;
    const int *pt_off, *pt_edges, *pose_off, *pose_edges, *pair_off, *pair_ij, *tuples;
    const double *pose0, *pt0;
    double *pose, *pose_bak, *pt, *pt_bak, *err, *J, *Hll, *bl, *Dinv, *db, *Hpp, *bp, *bs, *x, *Hs;
    unsigned char* level;
    double *poses_out, *points_out;
    unsigned char* outlier;
    orbba_stats_t* stats;
};

// ------------------------------------------------------------------------------------------------ SE3 (unit quaternion xyzw + t)
__device__ __forceinline__ void q_rotate(const double* q, const double* v, double* o) {
    double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
    ux += ux; uy += uy; uz += uz;
    o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ void se3_map(const double* s, const double* p, double* o) {
    q_rotate(s, p, o);
    o[0] += s[4]; o[1] += s[5]; o[2] += s[6];
}
__device__ __forceinline__ void q_normalize(double* q) {
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
__device__ __forceinline__ void q_mul(const double* a, const double* b, double* r) {
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
__host__ __device__ inline void q_from_matrix(const double* m, double* q) {   // Eigen::Quaterniond(Matrix3d)
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}
__host__ __device__ inline void q_to_matrix(const double* q, double* R) {    // Eigen toRotationMatrix
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// pose <- exp(u) * pose  (VertexSE3Expmap::oplusImpl, SE3Quat::exp, SE3Quat::operator*)
__device__ void se3_oplus(const double* u, double* s) {
    const double ox = u[0], oy = u[1], oz = u[2];
    const double theta = sqrt(ox * ox + oy * oy + oz * oz);
    const double Om[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
    double Om2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Om2[i * 3 + j] = Om[i * 3] * Om[j] + Om[i * 3 + 1] * Om[3 + j] + Om[i * 3 + 2] * Om[6 + j];
    double R[9], V[9];
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + Om[i] + Om2[i]; V[i] = R[i]; }
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3.0);
        for (int i = 0; i < 9; i++) {
            const double id = (i % 4 == 0 ? 1.0 : 0.0);
            R[i] = id + a * Om[i] + b * Om2[i];
            V[i] = id + b * Om[i] + c * Om2[i];
        }
    }
    double e[7];
    q_from_matrix(R, e);
    q_normalize(e);
    for (int i = 0; i < 3; i++) e[4 + i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
    double rt[3], rq[4];
    q_rotate(e, s + 4, rt);
    q_mul(e, s, rq);
    q_normalize(rq);
    s[0] = rq[0]; s[1] = rq[1]; s[2] = rq[2]; s[3] = rq[3];
    s[4] = e[4] + rt[0]; s[5] = e[5] + rt[1]; s[6] = e[6] + rt[2];
}

// ------------------------------------------------------------------------------------------------ block reductions (fixed order)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < BA_WARPS; w++) s += red[w];
    return s;
}
__device__ double block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < BA_WARPS; w++) s = fmax(s, red[w]);
    return s;
}

__device__ __forceinline__ void project_edge(const BAProb& P, int e, double* pc) {
    double pr[3];
    se3_map(P.pose + 7 * P.e_pose[e], P.pt + 3 * P.e_pt[e], pr);
    se3_map(P.cam + BA_CAM_STRIDE * P.e_cam[e] + 4, pr, pc);
}
__device__ __forceinline__ double huber_rho0(double e, double delta, double dsqr) { return e <= dsqr ? e : 2 * sqrt(e) * delta - dsqr; }

// computeActiveErrors + activeRobustChi2
__device__ double compute_errors(const BAProb& P, bool robust, double delta, double dsqr, double* red) {
    double local = 0;
    for (int e = threadIdx.x; e < P.nE; e += BA_T) {
        if (P.level[e]) continue;
        double pc[3];
        project_edge(P, e, pc);
        const double* c = P.cam + BA_CAM_STRIDE * P.e_cam[e];
        const double e0 = P.e_obs[2 * e] - (pc[0] / pc[2] * c[0] + c[2]);
        const double e1 = P.e_obs[2 * e + 1] - (pc[1] / pc[2] * c[1] + c[3]);
        P.err[2 * e] = e0; P.err[2 * e + 1] = e1;
        const double c2 = (e0 * e0 + e1 * e1) * P.e_info[e];
        local += robust ? huber_rho0(c2, delta, dsqr) : c2;
    }
    return block_sum(local, red);
}

// BlockSolver::buildSystem: linearizeOplus + constructQuadraticForm
__device__ void build_system(const BAProb& P, bool robust, double delta, double dsqr) {
    const int E = P.nE;
    for (int e = threadIdx.x; e < E; e += BA_T) {
        if (P.level[e]) continue;
        const double* c = P.cam + BA_CAM_STRIDE * P.e_cam[e];
        const double* ps = P.pose + 7 * P.e_pose[e];
        double pc[3];
        project_edge(P, e, pc);
        const double X = pc[0], Y = pc[1], Z = pc[2], iz = -1. / Z;
        const double t00 = iz * c[0], t02 = iz * (-X / Z * c[0]), t11 = iz * c[1], t12 = iz * (-Y / Z * c[1]);
        double* J = P.J + e;
        if (P.pose_free[P.e_pose[e]] >= 0) {
            // (-1/z * tmp) * J3, J3 = [-skew(p) | I]
            const double tJ[12] = {t02 * Y, t00 * Z - t02 * X, -t00 * Y, t00, 0, t02,
                                   -t11 * Z + t12 * Y, -t12 * X, t11 * X, 0, t11, t12};
            const double* A = c + 11;
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    double s = 0;
#pragma unroll
                    for (int k = 0; k < 6; k++) s += tJ[i * 6 + k] * A[k * 6 + j];
                    J[(size_t)(i * 6 + j) * E] = s;
                }
        }
        double q[4], R[9];
        q_mul(c + 4, ps, q);
        q_normalize(q);
        q_to_matrix(q, R);
#pragma unroll
        for (int j = 0; j < 3; j++) {
            J[(size_t)(12 + j) * E] = t00 * R[j] + t02 * R[6 + j];
            J[(size_t)(15 + j) * E] = t11 * R[3 + j] + t12 * R[6 + j];
        }
        const double w = P.e_info[e], e0 = P.err[2 * e], e1 = P.err[2 * e + 1];
        double wr = 1.0;
        if (robust) {
            const double c2 = (e0 * e0 + e1 * e1) * w;
            if (c2 > dsqr) wr = delta / sqrt(c2);
        }
        J[(size_t)18 * E] = wr * w;
        J[(size_t)19 * E] = -w * e0 * wr;
        J[(size_t)20 * E] = -w * e1 * wr;
    }
    __syncthreads();
    // landmarks
    for (int l = threadIdx.x; l < P.nL; l += BA_T) {
        double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0, b0 = 0, b1 = 0, b2 = 0;
        for (int k = P.pt_off[l]; k < P.pt_off[l + 1]; k++) {
            const int e = P.pt_edges[k];
            if (P.level[e]) continue;
            const double* J = P.J + e;
            const double a0 = J[(size_t)12 * E], a1 = J[(size_t)13 * E], a2 = J[(size_t)14 * E];
            const double c0 = J[(size_t)15 * E], c1 = J[(size_t)16 * E], c2 = J[(size_t)17 * E];
            const double W = J[(size_t)18 * E], r0 = J[(size_t)19 * E], r1 = J[(size_t)20 * E];
            h00 += (a0 * a0 + c0 * c0) * W; h01 += (a0 * a1 + c0 * c1) * W; h02 += (a0 * a2 + c0 * c2) * W;
            h11 += (a1 * a1 + c1 * c1) * W; h12 += (a1 * a2 + c1 * c2) * W; h22 += (a2 * a2 + c2 * c2) * W;
            b0 += a0 * r0 + c0 * r1; b1 += a1 * r0 + c1 * r1; b2 += a2 * r0 + c2 * r1;
        }
        double* H = P.Hll + 9 * (size_t)l;
        H[0] = h00; H[1] = h01; H[2] = h02; H[3] = h01; H[4] = h11; H[5] = h12; H[6] = h02; H[7] = h12; H[8] = h22;
        P.bl[3 * l] = b0; P.bl[3 * l + 1] = b1; P.bl[3 * l + 2] = b2;
    }
    // poses: warp per free pose
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < P.K; k += BA_WARPS) {
        double h[21], b[6];
#pragma unroll
        for (int i = 0; i < 21; i++) h[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = 0;
        for (int t = P.pose_off[k] + lane; t < P.pose_off[k + 1]; t += 32) {
            const int e = P.pose_edges[t];
            if (P.level[e]) continue;
            const double* J = P.J + e;
            double a[6], c[6];
#pragma unroll
            for (int i = 0; i < 6; i++) { a[i] = J[(size_t)i * E]; c[i] = J[(size_t)(6 + i) * E]; }
            const double W = J[(size_t)18 * E], r0 = J[(size_t)19 * E], r1 = J[(size_t)20 * E];
            int u = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                b[i] += a[i] * r0 + c[i] * r1;
#pragma unroll
                for (int j = i; j < 6; j++) h[u++] += (a[i] * a[j] + c[i] * c[j]) * W;
            }
        }
#pragma unroll
        for (int i = 0; i < 21; i++) h[i] = warp_sum(h[i]);
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = warp_sum(b[i]);
        if (lane == 0) {
            double* H = P.Hpp + 36 * (size_t)k;
            int u = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                P.bp[6 * k + i] = b[i];
#pragma unroll
                for (int j = i; j < 6; j++) { H[i * 6 + j] = h[u]; H[j * 6 + i] = h[u]; u++; }
            }
        }
    }
    __syncthreads();
}

// BlockSolver::setLambda + solve: Schur complement, LDL^T of the reduced camera system, landmark back-substitution.
// Returns false when the factorisation meets a zero pivot (g2o: linear solver failure -> the trial is rejected).
__device__ bool solve_system(const BAProb& P, double lambda, double* Hs, int* s_flag) {
    const int E = P.nE, n = P.n, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int l = tid; l < P.nL; l += BA_T) {
        const double* H = P.Hll + 9 * (size_t)l;
        const double m0 = H[0] + lambda, m1 = H[1], m2 = H[2], m3 = H[3], m4 = H[4] + lambda, m5 = H[5], m6 = H[6], m7 = H[7], m8 = H[8] + lambda;
        const double c00 = m4 * m8 - m5 * m7, c01 = m5 * m6 - m3 * m8, c02 = m3 * m7 - m4 * m6;
        const double id = 1.0 / (m0 * c00 + m1 * c01 + m2 * c02);
        double* D = P.Dinv + 9 * (size_t)l;
        D[0] = c00 * id; D[1] = (m2 * m7 - m1 * m8) * id; D[2] = (m1 * m5 - m2 * m4) * id;
        D[3] = c01 * id; D[4] = (m0 * m8 - m2 * m6) * id; D[5] = (m2 * m3 - m0 * m5) * id;
        D[6] = c02 * id; D[7] = (m1 * m6 - m0 * m7) * id; D[8] = (m0 * m4 - m1 * m3) * id;
        const double b0 = P.bl[3 * l], b1 = P.bl[3 * l + 1], b2 = P.bl[3 * l + 2];
        P.db[3 * l] = D[0] * b0 + D[1] * b1 + D[2] * b2;
        P.db[3 * l + 1] = D[3] * b0 + D[4] * b1 + D[5] * b2;
        P.db[3 * l + 2] = D[6] * b0 + D[7] * b1 + D[8] * b2;
    }
    for (int i = tid; i < n * n; i += BA_T) {
        const int r = i / n, c = i - r * n;
        const int kr = r / 6, kc = c / 6;
        Hs[i] = kr == kc ? P.Hpp[36 * (size_t)kr + (r - 6 * kr) * 6 + (c - 6 * kc)] + (r == c ? lambda : 0.0) : 0.0;
    }
    if (tid == 0) *s_flag = 1;
    __syncthreads();
    // reduced camera matrix: one warp per pose pair
    for (int s = warp; s < P.nPairs; s += BA_WARPS) {
        double acc[36];
#pragma unroll
        for (int i = 0; i < 36; i++) acc[i] = 0;
        for (int t = P.pair_off[s] + lane; t < P.pair_off[s + 1]; t += 32) {
            const int a1 = P.tuples[2 * t], a2 = P.tuples[2 * t + 1];
            if (P.level[a1] | P.level[a2]) continue;
            const double* D = P.Dinv + 9 * (size_t)P.e_pt[a1];
            const double* J1 = P.J + a1;
            const double* J2 = P.J + a2;
            const double u0 = J1[(size_t)12 * E], u1 = J1[(size_t)13 * E], u2 = J1[(size_t)14 * E];
            const double v0 = J1[(size_t)15 * E], v1 = J1[(size_t)16 * E], v2 = J1[(size_t)17 * E];
            const double ww = J1[(size_t)18 * E] * J2[(size_t)18 * E];
            // G = Jl1 * Dinv (2x3), S = G * Jl2^T * (W1 W2) (2x2)
            const double g00 = u0 * D[0] + u1 * D[3] + u2 * D[6], g01 = u0 * D[1] + u1 * D[4] + u2 * D[7], g02 = u0 * D[2] + u1 * D[5] + u2 * D[8];
            const double g10 = v0 * D[0] + v1 * D[3] + v2 * D[6], g11 = v0 * D[1] + v1 * D[4] + v2 * D[7], g12 = v0 * D[2] + v1 * D[5] + v2 * D[8];
            const double p0 = J2[(size_t)12 * E], p1 = J2[(size_t)13 * E], p2 = J2[(size_t)14 * E];
            const double q0 = J2[(size_t)15 * E], q1 = J2[(size_t)16 * E], q2 = J2[(size_t)17 * E];
            const double s00 = (g00 * p0 + g01 * p1 + g02 * p2) * ww, s01 = (g00 * q0 + g01 * q1 + g02 * q2) * ww;
            const double s10 = (g10 * p0 + g11 * p1 + g12 * p2) * ww, s11 = (g10 * q0 + g11 * q1 + g12 * q2) * ww;
            double m0[6], m1[6];   // rows of S * Jp2 (2x6)
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const double x0 = J2[(size_t)j * E], x1 = J2[(size_t)(6 + j) * E];
                m0[j] = s00 * x0 + s01 * x1;
                m1[j] = s10 * x0 + s11 * x1;
            }
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const double y0 = J1[(size_t)i * E], y1 = J1[(size_t)(6 + i) * E];
#pragma unroll
                for (int j = 0; j < 6; j++) acc[i * 6 + j] += y0 * m0[j] + y1 * m1[j];
            }
        }
#pragma unroll
        for (int i = 0; i < 36; i++) acc[i] = warp_sum(acc[i]);
        const int i1 = P.pair_ij[2 * s], i2 = P.pair_ij[2 * s + 1];
        // lanes write the entries they are responsible for: entry id = lane and lane + 32
#pragma unroll
        for (int i = 0; i < 36; i++) {
            if ((i & 31) == lane) {
                const int r = i / 6, c = i - r * 6;
                const size_t a = (size_t)(6 * i1 + r) * n + 6 * i2 + c;
                const double v = Hs[a] - acc[i];
                Hs[a] = v;
                if (i1 != i2) Hs[(size_t)(6 * i2 + c) * n + 6 * i1 + r] = v;
            }
        }
    }
    // Schur right-hand side: bs = bp - sum_e B_e db_l
    for (int k = warp; k < P.K; k += BA_WARPS) {
        double cf[6];
#pragma unroll
        for (int i = 0; i < 6; i++) cf[i] = 0;
        for (int t = P.pose_off[k] + lane; t < P.pose_off[k + 1]; t += 32) {
            const int e = P.pose_edges[t];
            if (P.level[e]) continue;
            const double* J = P.J + e;
            const double* d = P.db + 3 * (size_t)P.e_pt[e];
            const double W = J[(size_t)18 * E];
            const double s0 = (J[(size_t)12 * E] * d[0] + J[(size_t)13 * E] * d[1] + J[(size_t)14 * E] * d[2]) * W;
            const double s1 = (J[(size_t)15 * E] * d[0] + J[(size_t)16 * E] * d[1] + J[(size_t)17 * E] * d[2]) * W;
#pragma unroll
            for (int i = 0; i < 6; i++) cf[i] += J[(size_t)i * E] * s0 + J[(size_t)(6 + i) * E] * s1;
        }
#pragma unroll
        for (int i = 0; i < 6; i++) cf[i] = warp_sum(cf[i]);
        if (lane < 6) {
            double v = cf[0];
#pragma unroll
            for (int i = 1; i < 6; i++) if (lane == i) v = cf[i];
            P.bs[6 * k + lane] = P.bp[6 * k + lane] - v;
        }
    }
    __syncthreads();
    // LDL^T in place (lower triangle: L below the diagonal, D on it); right-looking
    for (int j = 0; j < n; j++) {
        const double dj = Hs[(size_t)j * n + j];
        if (dj == 0.0 || !isfinite(dj)) { if (tid == 0) *s_flag = 0; break; }
        for (int i = j + 1 + tid; i < n; i += BA_T) Hs[(size_t)i * n + j] /= dj;
        __syncthreads();
        const int m = n - j - 1;
        for (int idx = tid; idx < m * m; idx += BA_T) {
            const int a = idx / m, b = idx - a * m;
            if (b > a) continue;
            const int i = j + 1 + a, k = j + 1 + b;
            Hs[(size_t)i * n + k] -= Hs[(size_t)i * n + j] * dj * Hs[(size_t)k * n + j];
        }
        __syncthreads();
    }
    __syncthreads();
    const bool ok = *s_flag != 0;
    if (ok && warp == 0) {
        // forward, diagonal, backward substitution by one warp (x kept in global memory, n <= a few hundred)
        double* x = P.x;
        for (int i = lane; i < n; i += 32) x[i] = P.bs[i];
        __syncwarp();
        for (int j = 0; j < n; j++) {
            const double xj = x[j];
            for (int i = j + 1 + lane; i < n; i += 32) x[i] -= Hs[(size_t)i * n + j] * xj;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) x[i] /= Hs[(size_t)i * n + i];
        __syncwarp();
        for (int j = n - 1; j >= 0; j--) {
            const double xj = x[j];
            for (int i = lane; i < j; i += 32) x[i] -= Hs[(size_t)j * n + i] * xj;
            __syncwarp();
        }
    }
    __syncthreads();
    if (!ok) {
        for (int i = tid; i < n + 3 * P.nL; i += BA_T) P.x[i] = 0.0;   // g2o applies the stale x; the trial is rejected either way
        __syncthreads();
        return false;
    }
    // landmarks: xl = Dinv (bl - B^T xp)
    for (int l = tid; l < P.nL; l += BA_T) {
        double c0 = P.bl[3 * l], c1 = P.bl[3 * l + 1], c2 = P.bl[3 * l + 2];
        for (int k = P.pt_off[l]; k < P.pt_off[l + 1]; k++) {
            const int e = P.pt_edges[k];
            if (P.level[e]) continue;
            const int pi = P.pose_free[P.e_pose[e]];
            if (pi < 0) continue;
            const double* J = P.J + e;
            const double* xp = P.x + 6 * pi;
            double s0 = 0, s1 = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) { s0 += J[(size_t)i * E] * xp[i]; s1 += J[(size_t)(6 + i) * E] * xp[i]; }
            const double W = J[(size_t)18 * E];
            s0 *= W; s1 *= W;
            c0 -= J[(size_t)12 * E] * s0 + J[(size_t)15 * E] * s1;
            c1 -= J[(size_t)13 * E] * s0 + J[(size_t)16 * E] * s1;
            c2 -= J[(size_t)14 * E] * s0 + J[(size_t)17 * E] * s1;
        }
        const double* D = P.Dinv + 9 * (size_t)l;
        P.x[n + 3 * l] = D[0] * c0 + D[1] * c1 + D[2] * c2;
        P.x[n + 3 * l + 1] = D[3] * c0 + D[4] * c1 + D[5] * c2;
        P.x[n + 3 * l + 2] = D[6] * c0 + D[7] * c1 + D[8] * c2;
    }
    __syncthreads();
    return true;
}

// the stop flag lives in mapped host memory and may change at any time: one thread reads it, everybody uses that value
__device__ bool read_stop(const volatile int* stop, int* s_tmp) {
    __syncthreads();
    if (threadIdx.x == 0) *s_tmp = stop ? *stop : 0;
    __syncthreads();
    return *s_tmp != 0;
}

struct LMState {
    double lambda, ni, currentChi, tempChi, rho, iniChi;
    int nBad, qmax, result, iterations, trials, stopped;
};

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg::solve inside
__device__ void optimize(const BAProb& P, int iterations, bool robust, double delta, double* Hs, double* red, int* s_flag, int* s_stop,
                         LMState* S, const volatile int* stop) {
    const double dsqr = delta * delta;
    const int tid = threadIdx.x, n = P.n, nx = P.n + 3 * P.nL;
    {   // SparseOptimizer::optimize returns at once when nothing is active (sparse_optimizer.cpp:356-359)
        double na = 0;
        for (int e = tid; e < P.nE; e += BA_T) na += P.level[e] == 0;
        if (block_sum(na, red) == 0.0) return;
    }
    bool ok = true;
    for (int it = 0; it < iterations && ok; it++) {
        if (read_stop(stop, s_stop)) { if (tid == 0) S->stopped = 1; break; }
        const double chi0 = compute_errors(P, robust, delta, dsqr, red);
        build_system(P, robust, delta, dsqr);
        if (it == 0) {
            double md = 0;
            for (int i = tid; i < 6 * P.K; i += BA_T) md = fmax(md, fabs(P.Hpp[36 * (size_t)(i / 6) + (i % 6) * 7]));
            for (int i = tid; i < 3 * P.nL; i += BA_T) md = fmax(md, fabs(P.Hll[9 * (size_t)(i / 3) + (i % 3) * 4]));
            md = block_max(md, red);
            if (tid == 0) { S->lambda = 1e-5 * md; S->ni = 2; S->nBad = 0; }
        }
        if (tid == 0) { S->currentChi = chi0; S->iniChi = chi0; S->qmax = 0; S->rho = 0; }
        __syncthreads();
        bool again = true;
        while (again) {
            for (int i = tid; i < 7 * P.nP; i += BA_T) P.pose_bak[i] = P.pose[i];     // push()
            for (int i = tid; i < 3 * P.nL; i += BA_T) P.pt_bak[i] = P.pt[i];
            const double lambda = S->lambda;
            const bool ok2 = solve_system(P, lambda, Hs, s_flag);
            for (int i = tid; i < P.nP; i += BA_T) {                                   // update()
                const int k = P.pose_free[i];
                if (k >= 0) se3_oplus(P.x + 6 * k, P.pose + 7 * i);
            }
            for (int i = tid; i < 3 * P.nL; i += BA_T) P.pt[i] += P.x[n + i];
            __syncthreads();
            double tempChi = compute_errors(P, robust, delta, dsqr, red);
            if (!ok2) tempChi = 1.7976931348623157e308;
            double sc = 0;                                                              // computeScale()
            for (int j = tid; j < nx; j += BA_T) {
                const double xj = P.x[j], bj = j < n ? P.bp[j] : P.bl[j - n];
                sc += xj * (lambda * xj + bj);
            }
            sc = block_sum(sc, red) + 1e-3;
            if (tid == 0) {
                const double rho = (S->currentChi - tempChi) / sc;
                S->rho = rho;
                S->trials++;
                if (rho > 0 && isfinite(tempChi)) {
                    double alpha = 1. - pow(2 * rho - 1, 3.0);
                    alpha = fmin(alpha, 2. / 3.);
                    S->lambda *= fmax(1. / 3., alpha);
                    S->ni = 2;
                    S->currentChi = tempChi;
                    S->result = 1;
                } else {
                    S->lambda *= S->ni;
                    S->ni *= 2;
                    S->result = 0;
                }
                S->qmax++;
            }
            __syncthreads();
            if (S->result == 0) {                                                       // pop()
                for (int i = tid; i < 7 * P.nP; i += BA_T) P.pose[i] = P.pose_bak[i];
                for (int i = tid; i < 3 * P.nL; i += BA_T) P.pt[i] = P.pt_bak[i];
            }
            const bool stopped = read_stop(stop, s_stop);
            again = S->rho < 0 && S->qmax < 10 && !stopped;
            __syncthreads();
        }
        if (tid == 0) {
            S->iterations++;
            int res = 0;   // OK
            if (S->qmax == 10 || S->rho == 0) res = 1;
            else {
                if ((S->iniChi - S->currentChi) * 1e3 < S->iniChi) S->nBad++; else S->nBad = 0;
                if (S->nBad >= 3) res = 1;
            }
            S->result = res;
        }
        __syncthreads();
        ok = S->result == 0;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BA_T, 1) ba_kernel(const BAProb* __restrict__ probs, int its1, int its2, double delta, double chi2_th,
                                                     const volatile int* stop, int hs_smem_n) {
    extern __shared__ double sm_hs[];
    __shared__ double red[BA_WARPS];
    __shared__ int s_flag, s_stop;
    __shared__ LMState S;
    __shared__ BAProb sP;
    const int tid = threadIdx.x;
    if (tid == 0) sP = probs[blockIdx.x];
    __syncthreads();
    const BAProb& P = sP;
    double* Hs = P.n <= hs_smem_n ? sm_hs : P.Hs;
    for (int i = tid; i < 7 * P.nP; i += BA_T) P.pose[i] = P.pose0[i];
    for (int i = tid; i < 3 * P.nL; i += BA_T) P.pt[i] = P.pt0[i];
    for (int i = tid; i < P.nE; i += BA_T) { P.level[i] = 0; P.err[2 * i] = 0; P.err[2 * i + 1] = 0; }
    if (tid == 0) { memset(&S, 0, sizeof(S)); }
    __syncthreads();
    const bool robust1 = delta > 0;
    const double dsqr = delta * delta;
    const bool stopped0 = read_stop(stop, &s_stop);
    double initial = 0;
    if (!stopped0 && P.nE > 0) {
        initial = compute_errors(P, robust1, delta, dsqr, red);
        optimize(P, its1, robust1, delta, Hs, red, &s_flag, &s_stop, &S, stop);
        const bool more = its2 >= 0 && !read_stop(stop, &s_stop);
        if (more) {
            // chi2 > th or non-positive depth -> level 1; Huber off (src/Optimizer.cc:598-613)
            for (int e = tid; e < P.nE; e += BA_T) {
                const double c2 = (P.err[2 * e] * P.err[2 * e] + P.err[2 * e + 1] * P.err[2 * e + 1]) * P.e_info[e];
                double pc[3];
                project_edge(P, e, pc);
                if (c2 > chi2_th || !(pc[2] > 0.0)) P.level[e] = 1;
            }
            __syncthreads();
            optimize(P, its2, false, delta, Hs, red, &s_flag, &s_stop, &S, stop);
        }
    }
    __syncthreads();
    int nout = 0;
    for (int e = tid; e < P.nE; e += BA_T) {
        const double c2 = (P.err[2 * e] * P.err[2 * e] + P.err[2 * e + 1] * P.err[2 * e + 1]) * P.e_info[e];
        double pc[3];
        project_edge(P, e, pc);
        const bool out = c2 > chi2_th || !(pc[2] > 0.0);
        P.outlier[e] = out;
        nout += out;
    }
    const double tot = block_sum((double)nout, red);
    for (int i = tid; i < P.nP; i += BA_T) {
        double R[9];
        const double* s = P.pose + 7 * i;
        q_to_matrix(s, R);
        double* o = P.poses_out + 12 * (size_t)i;
        for (int r = 0; r < 3; r++) { o[r * 4] = R[r * 3]; o[r * 4 + 1] = R[r * 3 + 1]; o[r * 4 + 2] = R[r * 3 + 2]; o[r * 4 + 3] = s[4 + r]; }
    }
    for (int i = tid; i < 3 * P.nL; i += BA_T) P.points_out[i] = P.pt[i];
    if (tid == 0) {
        orbba_stats_t st;
        st.initial_chi2 = initial; st.final_chi2 = S.currentChi; st.final_lambda = S.lambda;
        st.iterations = S.iterations; st.trials = S.trials; st.outliers = (int)tot;
        st.status = stopped0 ? ORB_E_ABORTED : ORB_OK;
        *P.stats = st;
    }
}

// ================================================================================================ host side
struct orbba {
    int device = 0, max_problems = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // uploaded batch
    int n = 0;
    std::vector<BAProb> probs;
    std::vector<int> nP, nL, nE;
    uint8_t* d_static = nullptr; size_t static_cap = 0;
    uint8_t* d_dynamic = nullptr; size_t dynamic_cap = 0;
    BAProb* d_probs = nullptr; size_t probs_cap = 0;
    std::vector<size_t> out_off;     // per problem: offsets of poses_out, points_out, outlier, stats inside d_dynamic
    int* h_stop = nullptr;           // pinned + mapped: device-visible stop flag
    int* d_stop = nullptr;
    int max_n = 0, hs_smem_n = 0;
    size_t smem_bytes = 0;
    long long launches = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool profile = false;
    double prof_ms = 0; int prof_calls = 0; bool prof_pending = false;
};

static void orbba_free(orbba* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    cudaFree(b->d_static); cudaFree(b->d_dynamic); cudaFree(b->d_probs);
    if (b->h_stop) cudaFreeHost(b->h_stop);
    if (b->ev[0]) { cudaEventDestroy(b->ev[0]); cudaEventDestroy(b->ev[1]); }
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    delete b;
}

namespace {
struct Blob {
    std::vector<uint8_t> bytes;
    size_t add(const void* p, size_t n) {
        const size_t off = (bytes.size() + 15) & ~(size_t)15;
        bytes.resize(off + n);
        if (p && n) memcpy(bytes.data() + off, p, n);
        return off;
    }
};
size_t bump(size_t& cur, size_t n) { const size_t off = (cur + 15) & ~(size_t)15; cur = off + n; return off; }
}  // namespace

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
    cudaError_t ce = cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaHostAlloc((void**)&b->h_stop, sizeof(int), cudaHostAllocMapped);
    if (ce == cudaSuccess) { *b->h_stop = 0; ce = cudaHostGetDevicePointer((void**)&b->d_stop, b->h_stop, 0); }
    if (ce == cudaSuccess) ce = cudaEventCreate(&b->ev[0]);
    if (ce == cudaSuccess) ce = cudaEventCreate(&b->ev[1]);
    // reduced camera system in shared memory up to 200 KB
    b->hs_smem_n = 156;   // 156^2 * 8 = 194,688 B  (26 free poses)
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(ba_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, b->hs_smem_n * b->hs_smem_n * 8);
    if (ce != cudaSuccess) { int rc = orbhost::check_cuda(ce, "orbba_create", __FILE__, __LINE__); orbba_free(b); return rc; }
    b->stream = b->own_stream;
    *out = b;
    return ORB_OK;
}

void orbba_destroy(orbba_t* b) { orbba_free(b); }

int orbba_set_stream(orbba_t* b, void* s) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_set_stream: NULL handle");
    b->stream = s ? (cudaStream_t)s : b->own_stream;
    return ORB_OK;
}
int orbba_synchronize(orbba_t* b) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_synchronize: NULL handle");
    ORB_CUDA(cudaSetDevice(b->device));
    ORB_CUDA(cudaStreamSynchronize(b->stream));
    return ORB_OK;
}
long long orbba_launch_count(const orbba_t* b) { return b ? b->launches : 0; }

// Flattens, indexes and uploads a batch of problems (host -> device, asynchronous on the handle's stream).
int orbba_upload(orbba_t* b, const orbba_problem_t* problems, int n) {
    if (!b || (!problems && n > 0)) ORB_FAIL(ORB_E_INVALID, "orbba_upload: bad argument");
    if (n < 0 || n > b->max_problems) ORB_FAIL(ORB_E_INVALID, "orbba_upload: n=%d exceeds max_problems=%d", n, b->max_problems);
    ORB_CUDA(cudaSetDevice(b->device));
    Blob blob;
    b->probs.assign(n, BAProb());
    b->nP.assign(n, 0); b->nL.assign(n, 0); b->nE.assign(n, 0);
    b->out_off.assign((size_t)n * 4, 0);
    std::vector<std::vector<size_t>> soff(n);
    size_t dyn = 0;
    std::vector<std::vector<size_t>> doff(n);
    b->max_n = 0;
    for (int p = 0; p < n; p++) {
        const orbba_problem_t& Q = problems[p];
        const int nP = Q.n_poses, nL = Q.n_points, nE = Q.n_edges, nC = Q.n_cams;
        if (nP < 0 || nL < 0 || nE < 0 || nC < 1) ORB_FAIL(ORB_E_INVALID, "orbba_upload: problem %d has negative sizes", p);
        if ((nP && (!Q.poses || !Q.pose_fixed)) || (nL && !Q.points) || (nE && (!Q.edge_pose || !Q.edge_point || !Q.edge_cam || !Q.edge_obs || !Q.edge_inv_sigma2)) ||
            !Q.cam_K || !Q.cam_ext || !Q.cam_adj)
            ORB_FAIL(ORB_E_INVALID, "orbba_upload: problem %d has a NULL array", p);
        std::vector<int> pose_free(nP, -1);
        int K = 0;
        for (int i = 0; i < nP; i++) if (!Q.pose_fixed[i]) pose_free[i] = K++;
        for (int e = 0; e < nE; e++)
            if (Q.edge_pose[e] < 0 || Q.edge_pose[e] >= nP || Q.edge_point[e] < 0 || Q.edge_point[e] >= nL || Q.edge_cam[e] < 0 || Q.edge_cam[e] >= nC)
                ORB_FAIL(ORB_E_INVALID, "orbba_upload: problem %d edge %d indexes out of range", p, e);
        // CSR by landmark / by free pose (edge ids ascending inside every list)
        std::vector<int> pt_off(nL + 1, 0), pose_off(K + 1, 0);
        for (int e = 0; e < nE; e++) { pt_off[Q.edge_point[e] + 1]++; const int k = pose_free[Q.edge_pose[e]]; if (k >= 0) pose_off[k + 1]++; }
        for (int i = 0; i < nL; i++) pt_off[i + 1] += pt_off[i];
        for (int i = 0; i < K; i++) pose_off[i + 1] += pose_off[i];
        std::vector<int> pt_edges(nE), pose_edges(pose_off[K]);
        {
            std::vector<int> c1(pt_off.begin(), pt_off.end() - 1), c2(pose_off.begin(), pose_off.end() - 1);
            for (int e = 0; e < nE; e++) { pt_edges[c1[Q.edge_point[e]]++] = e; const int k = pose_free[Q.edge_pose[e]]; if (k >= 0) pose_edges[c2[k]++] = e; }
        }
        // (edge, edge) tuples per pose pair (upper block triangle incl. diagonal)
        std::vector<std::vector<int>> per_pair((size_t)K * K);
        for (int l = 0; l < nL; l++)
            for (int a = pt_off[l]; a < pt_off[l + 1]; a++) {
                const int ea = pt_edges[a], ka = pose_free[Q.edge_pose[ea]];
                if (ka < 0) continue;
                for (int c = a; c < pt_off[l + 1]; c++) {
                    const int ec = pt_edges[c], kc = pose_free[Q.edge_pose[ec]];
                    if (kc < 0) continue;
                    if (ka <= kc) { per_pair[(size_t)ka * K + kc].push_back(ea); per_pair[(size_t)ka * K + kc].push_back(ec); }
                    else { per_pair[(size_t)kc * K + ka].push_back(ec); per_pair[(size_t)kc * K + ka].push_back(ea); }
                }
            }
        std::vector<int> pair_off(1, 0), pair_ij, tuples;
        for (int i = 0; i < K; i++)
            for (int j = i; j < K; j++) {
                const std::vector<int>& v = per_pair[(size_t)i * K + j];
                if (v.empty()) continue;
                pair_ij.push_back(i); pair_ij.push_back(j);
                tuples.insert(tuples.end(), v.begin(), v.end());
                pair_off.push_back((int)tuples.size() / 2);
            }
        const int nPairs = (int)pair_ij.size() / 2, nTuples = (int)tuples.size() / 2;
        // cameras and initial estimates
        std::vector<double> cam((size_t)nC * BA_CAM_STRIDE), pose0((size_t)7 * nP);
        for (int c = 0; c < nC; c++) {
            double* D = &cam[(size_t)c * BA_CAM_STRIDE];
            for (int i = 0; i < 4; i++) D[i] = Q.cam_K[4 * c + i];
            const double* T = Q.cam_ext + 12 * c;
            const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
            q_from_matrix(R, D + 4);
            if (D[7] < 0) for (int i = 4; i < 8; i++) D[i] = -D[i];
            const double nn = sqrt(D[4] * D[4] + D[5] * D[5] + D[6] * D[6] + D[7] * D[7]);
            for (int i = 4; i < 8; i++) D[i] /= nn;
            D[8] = T[3]; D[9] = T[7]; D[10] = T[11];
            for (int i = 0; i < 36; i++) D[11 + i] = Q.cam_adj[36 * c + i];
        }
        for (int i = 0; i < nP; i++) {   // Converter::toSE3Quat + SE3Quat(R, t)
            const double* T = Q.poses + 12 * i;
            double* D = &pose0[(size_t)7 * i];
            const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
            q_from_matrix(R, D);
            if (D[3] < 0) for (int k = 0; k < 4; k++) D[k] = -D[k];
            const double nn = sqrt(D[0] * D[0] + D[1] * D[1] + D[2] * D[2] + D[3] * D[3]);
            for (int k = 0; k < 4; k++) D[k] /= nn;
            D[4] = T[3]; D[5] = T[7]; D[6] = T[11];
        }
        BAProb& P = b->probs[p];
        P.nP = nP; P.nL = nL; P.nE = nE; P.nC = nC; P.K = K; P.n = 6 * K; P.nPairs = nPairs; P.nTuples = nTuples;
        b->nP[p] = nP; b->nL[p] = nL; b->nE[p] = nE;
        b->max_n = std::max(b->max_n, 6 * K);
        std::vector<size_t>& so = soff[p];
        so.push_back(blob.add(Q.edge_pose, sizeof(int) * nE));
        so.push_back(blob.add(Q.edge_point, sizeof(int) * nE));
        so.push_back(blob.add(Q.edge_cam, sizeof(int) * nE));
        so.push_back(blob.add(pose_free.data(), sizeof(int) * nP));
        so.push_back(blob.add(Q.edge_obs, sizeof(double) * 2 * nE));
        so.push_back(blob.add(Q.edge_inv_sigma2, sizeof(double) * nE));
        so.push_back(blob.add(cam.data(), sizeof(double) * cam.size()));
        so.push_back(blob.add(pt_off.data(), sizeof(int) * pt_off.size()));
        so.push_back(blob.add(pt_edges.data(), sizeof(int) * pt_edges.size()));
        so.push_back(blob.add(pose_off.data(), sizeof(int) * pose_off.size()));
        so.push_back(blob.add(pose_edges.data(), sizeof(int) * pose_edges.size()));
        so.push_back(blob.add(pair_off.data(), sizeof(int) * pair_off.size()));
        so.push_back(blob.add(pair_ij.data(), sizeof(int) * pair_ij.size()));
        so.push_back(blob.add(tuples.data(), sizeof(int) * tuples.size()));
        so.push_back(blob.add(pose0.data(), sizeof(double) * pose0.size()));
        so.push_back(blob.add(Q.points, sizeof(double) * 3 * nL));
        // dynamic arena
        std::vector<size_t>& d = doff[p];
        const size_t nn = (size_t)6 * K;
        const size_t sizes[] = {sizeof(double) * 7 * nP, sizeof(double) * 7 * nP, sizeof(double) * 3 * nL, sizeof(double) * 3 * nL,
                                sizeof(double) * 2 * nE, sizeof(double) * BA_JSTRIDE * nE, sizeof(double) * 9 * nL, sizeof(double) * 3 * nL,
                                sizeof(double) * 9 * nL, sizeof(double) * 3 * nL, sizeof(double) * 36 * K, sizeof(double) * 6 * K,
                                sizeof(double) * 6 * K, sizeof(double) * (nn + 3 * nL), nn > (size_t)b->hs_smem_n ? sizeof(double) * nn * nn : 0,
                                (size_t)nE, sizeof(double) * 12 * nP, sizeof(double) * 3 * nL, (size_t)nE, sizeof(orbba_stats_t)};
        for (size_t s : sizes) d.push_back(bump(dyn, s));
    }
    if (blob.bytes.size() > b->static_cap) {
        cudaFree(b->d_static); b->d_static = nullptr; b->static_cap = 0;
        ORB_CUDA(cudaMalloc((void**)&b->d_static, blob.bytes.size() + 64));
        b->static_cap = blob.bytes.size();
    }
    if (dyn > b->dynamic_cap) {
        cudaFree(b->d_dynamic); b->d_dynamic = nullptr; b->dynamic_cap = 0;
        ORB_CUDA(cudaMalloc((void**)&b->d_dynamic, dyn + 64));
        b->dynamic_cap = dyn;
    }
    if ((size_t)n > b->probs_cap) {
        cudaFree(b->d_probs); b->d_probs = nullptr; b->probs_cap = 0;
        ORB_CUDA(cudaMalloc((void**)&b->d_probs, sizeof(BAProb) * (size_t)n));
        b->probs_cap = n;
    }
    for (int p = 0; p < n; p++) {
        BAProb& P = b->probs[p];
        const std::vector<size_t>& so = soff[p];
        const uint8_t* S = b->d_static;
        P.e_pose = (const int*)(S + so[0]); P.e_pt = (const int*)(S + so[1]); P.e_cam = (const int*)(S + so[2]); P.pose_free = (const int*)(S + so[3]);
        P.e_obs = (const double*)(S + so[4]); P.e_info = (const double*)(S + so[5]); P.cam = (const double*)(S + so[6]);
        P.pt_off = (const int*)(S + so[7]); P.pt_edges = (const int*)(S + so[8]); P.pose_off = (const int*)(S + so[9]); P.pose_edges = (const int*)(S + so[10]);
        P.pair_off = (const int*)(S + so[11]); P.pair_ij = (const int*)(S + so[12]); P.tuples = (const int*)(S + so[13]);
        P.pose0 = (const double*)(S + so[14]); P.pt0 = (const double*)(S + so[15]);
        const std::vector<size_t>& d = doff[p];
        uint8_t* D = b->d_dynamic;
        P.pose = (double*)(D + d[0]); P.pose_bak = (double*)(D + d[1]); P.pt = (double*)(D + d[2]); P.pt_bak = (double*)(D + d[3]);
        P.err = (double*)(D + d[4]); P.J = (double*)(D + d[5]); P.Hll = (double*)(D + d[6]); P.bl = (double*)(D + d[7]);
        P.Dinv = (double*)(D + d[8]); P.db = (double*)(D + d[9]); P.Hpp = (double*)(D + d[10]); P.bp = (double*)(D + d[11]);
        P.bs = (double*)(D + d[12]); P.x = (double*)(D + d[13]); P.Hs = (double*)(D + d[14]);
        P.level = D + d[15]; P.poses_out = (double*)(D + d[16]); P.points_out = (double*)(D + d[17]); P.outlier = D + d[18];
        P.stats = (orbba_stats_t*)(D + d[19]);
        b->out_off[(size_t)p * 4 + 0] = d[16]; b->out_off[(size_t)p * 4 + 1] = d[17]; b->out_off[(size_t)p * 4 + 2] = d[18]; b->out_off[(size_t)p * 4 + 3] = d[19];
    }
    b->n = n;
    if (n == 0) return ORB_OK;
    // the blob / descriptor vectors are pageable: these copies complete before returning
    ORB_CUDA(cudaMemcpyAsync(b->d_static, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice, b->stream));
    ORB_CUDA(cudaMemcpyAsync(b->d_probs, b->probs.data(), sizeof(BAProb) * (size_t)n, cudaMemcpyHostToDevice, b->stream));
    ORB_CUDA(cudaStreamSynchronize(b->stream));
    return ORB_OK;
}

// Runs the uploaded batch from its uploaded initial estimates (asynchronous on the handle's stream).
// its2 < 0: single round, no outlier pass (Optimizer::BundleAdjustment); huber_delta <= 0: no robust kernel.
int orbba_run(orbba_t* b, int its1, int its2, double huber_delta, double chi2_th) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_run: NULL handle");
    if (b->n == 0) return ORB_OK;
    if (its1 < 0) ORB_FAIL(ORB_E_INVALID, "orbba_run: its1 < 0");
    ORB_CUDA(cudaSetDevice(b->device));
    const int hs_n = std::min(b->max_n, b->hs_smem_n);
    const size_t smem = (size_t)hs_n * hs_n * sizeof(double);
    if (b->profile) ORB_CUDA(cudaEventRecord(b->ev[0], b->stream));
    ba_kernel<<<b->n, BA_T, smem, b->stream>>>(b->d_probs, its1, its2, huber_delta, chi2_th, b->d_stop, b->hs_smem_n);
    b->launches++;
    if (b->profile) { ORB_CUDA(cudaEventRecord(b->ev[1], b->stream)); b->prof_pending = true; }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orbba_profile(orbba_t* b, int enable) {
    if (!b) ORB_FAIL(ORB_E_INVALID, "orbba_profile: NULL handle");
    b->profile = enable != 0; b->prof_ms = 0; b->prof_calls = 0; b->prof_pending = false;
    return ORB_OK;
}
int orbba_stage_ms(orbba_t* b, double* ms1, int* calls) {   // only the LAST run is kept per synchronisation: call after each run
    if (!b || !ms1) ORB_FAIL(ORB_E_INVALID, "orbba_stage_ms: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
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

// Copies results of problem `p` of the last run to the host (synchronises the stream).  Any pointer may be NULL.
int orbba_download(orbba_t* b, int p, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats) {
    if (!b || p < 0 || p >= b->n) ORB_FAIL(ORB_E_INVALID, "orbba_download: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
    const size_t* o = &b->out_off[(size_t)p * 4];
    cudaStream_t st = b->stream;
    if (poses_out && b->nP[p]) ORB_CUDA(cudaMemcpyAsync(poses_out, b->d_dynamic + o[0], sizeof(double) * 12 * b->nP[p], cudaMemcpyDeviceToHost, st));
    if (points_out && b->nL[p]) ORB_CUDA(cudaMemcpyAsync(points_out, b->d_dynamic + o[1], sizeof(double) * 3 * b->nL[p], cudaMemcpyDeviceToHost, st));
    if (edge_outlier && b->nE[p]) ORB_CUDA(cudaMemcpyAsync(edge_outlier, b->d_dynamic + o[2], b->nE[p], cudaMemcpyDeviceToHost, st));
    if (stats) ORB_CUDA(cudaMemcpyAsync(stats, b->d_dynamic + o[3], sizeof(orbba_stats_t), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

// Optimizer::LocalBundleAdjustment for ONE problem, host buffers in and out, synchronous.  `stop` (may be NULL) is
// polled while the kernel runs and forwarded to the device, which checks it before every LM iteration and trial
// (g2o: SparseOptimizer::terminate(), sparse_optimizer.cpp:376, optimization_algorithm_levenberg.cpp:149).
int orbba_local(orbba_t* b, const orbba_problem_t* problem, int its1, int its2, double huber_delta, double chi2_th,
                const volatile uint8_t* stop, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats) {
    if (!b || !problem) ORB_FAIL(ORB_E_INVALID, "orbba_local: bad argument");
    ORB_CUDA(cudaSetDevice(b->device));
    *b->h_stop = (stop && *stop) ? 1 : 0;
    int rc = orbba_upload(b, problem, 1);
    if (rc != ORB_OK) return rc;
    rc = orbba_run(b, its1, its2, huber_delta, chi2_th);
    if (rc != ORB_OK) return rc;
    cudaEvent_t done = b->ev[1];
    if (!b->profile) ORB_CUDA(cudaEventRecord(done, b->stream));
    for (;;) {
        const cudaError_t q = cudaEventQuery(done);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) return orbhost::check_cuda(q, "cudaEventQuery", __FILE__, __LINE__);
        if (stop && *stop) *b->h_stop = 1;
    }
    orbba_stats_t st;
    rc = orbba_download(b, 0, poses_out, points_out, edge_outlier, &st);
    *b->h_stop = 0;
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
