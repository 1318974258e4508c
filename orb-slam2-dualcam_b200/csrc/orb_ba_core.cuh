// orb_ba_core.cuh -- device building blocks shared by the bundle-adjustment translation units (orb_ba.cu: batched LocalBA,
// orb_gba.cu: distributed GlobalBA): SE3 as unit quaternion + translation (g2o SE3Quat, Thirdparty/g2o/g2o/types/se3quat.h),
// the dual-camera reprojection of EdgeSE3ProjectXYZ (types_six_dof_expmap.cpp:109-169) and the Huber cost.
#pragma once
#include <math.h>

#define BA_CAM_STRIDE 48       // fx fy cx cy | ext quat xyzw | ext t | pad | adj[36]   (adjoint 16-byte aligned at +12)
#define BA_CAM_ADJ 12
#define BA_REC 22              // per-edge linearisation record: Jl[6] W r0 r1 Jp[12] pad (176 B: 16-byte aligned for 128-bit loads)

// ------------------------------------------------------------------------------------------------ SE3 (unit quaternion xyzw + t)
static __device__ __forceinline__ void q_rotate(const double* q, const double* v, double* o) {
    double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
    ux += ux; uy += uy; uz += uz;
    o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
static __device__ __forceinline__ void se3_map(const double* s, const double* p, double* o) {
    q_rotate(s, p, o);
    o[0] += s[4]; o[1] += s[5]; o[2] += s[6];
}
static __device__ __forceinline__ void q_normalize(double* q) {
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static __device__ __forceinline__ void q_mul(const double* a, const double* b, double* r) {
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
static __host__ __device__ inline void q_from_matrix(const double* m, double* q) {   // Eigen::Quaterniond(Matrix3d)
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
static __host__ __device__ inline void q_to_matrix(const double* q, double* R) {    // Eigen toRotationMatrix
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// out <- exp(u) * s  (VertexSE3Expmap::oplusImpl, SE3Quat::exp, SE3Quat::operator*)
static __device__ void se3_oplus(const double* u, const double* s, double* out) {
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
    out[0] = rq[0]; out[1] = rq[1]; out[2] = rq[2]; out[3] = rq[3];
    out[4] = e[4] + rt[0]; out[5] = e[5] + rt[1]; out[6] = e[6] + rt[2];
}

// One definition of the reprojection (point -> rig -> camera) and of the residual, never inlined, so that every kernel
// that evaluates an edge produces the same bits.
static __device__ __noinline__ void edge_project(const double* pose7, const double* pt3, const double* cam, double* pc) {
    double pr[3];
    se3_map(pose7, pt3, pr);
    se3_map(cam + 4, pr, pc);
}
static __device__ __noinline__ void edge_error(const double* pc, const double* cam, const double* obs, double* e2) {
    e2[0] = obs[0] - (pc[0] / pc[2] * cam[0] + cam[2]);
    e2[1] = obs[1] - (pc[1] / pc[2] * cam[1] + cam[3]);
}
static __device__ __forceinline__ double huber_rho0(double e, double delta, double dsqr) { return e <= dsqr ? e : 2 * sqrt(e) * delta - dsqr; }

