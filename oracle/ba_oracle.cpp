// ba_oracle.cpp -- CPU oracle of Optimizer::LocalBundleAdjustment / BundleAdjustment and the g2o pieces under them.
//
// TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  Dependency-free FP64 restatement; every block cites the reference
// lines it follows (paths relative to /root/reference).  Eigen (un-vendored dependency of g2o, >= 3.1.0) supplies
// Quaterniond <-> Matrix3d, the 3x3 inverse and SimplicialLDLT in the reference; their published algorithms are restated
// here (Shepperd-style matrix->quaternion, cofactor inverse, LDL^T without pivoting on the dense reduced camera system --
// the sparse ordering Eigen applies changes only the rounding, ~1e-13 relative, far inside the 1e-5 parity budget).
// PARITY UNPINNED by the reference (it has no tests / fixtures); tests/test_ba_oracle.py cross-checks one LM step
// against an independent numpy solve of the same normal equations and the convergence to the planted ground truth.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "orb_oracle.h"

namespace {

// ---------------------------------------------------------------- SE3Quat (Thirdparty/g2o/g2o/types/se3quat.h)
struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

void quat_normalize_rotation(Quat& q) {  // se3quat.h:280-285
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

// Eigen::Quaterniond(Matrix3d): trace branch, else largest diagonal element
Quat quat_from_matrix(const double m[9]) {
    Quat q;
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m[7] - m[5]) * t;
        q.y = (m[2] - m[6]) * t;
        q.z = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m[k * 3 + j] - m[j * 3 + k]) * t;
        v[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        v[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}

// Eigen::Quaterniond::toRotationMatrix
void quat_to_matrix(const Quat& q, double R[9]) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

Quat quat_mul(const Quat& a, const Quat& b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}

// Eigen quaternion * vector: v + w*uv + q.vec x uv with uv = 2 * (q.vec x v)
void quat_rotate(const Quat& q, const double v[3], double out[3]) {
    double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}

SE3 se3_from_Rt(const double* T12) {  // Converter::toSE3Quat src/Converter.cc:58-69 ; SE3Quat(R,t) se3quat.h:59-61
    SE3 s;
    const double R[9] = {T12[0], T12[1], T12[2], T12[4], T12[5], T12[6], T12[8], T12[9], T12[10]};
    s.r = quat_from_matrix(R);
    quat_normalize_rotation(s.r);
    s.t[0] = T12[3]; s.t[1] = T12[7]; s.t[2] = T12[11];
    return s;
}

void se3_to_Rt(const SE3& s, double* T12) {  // SE3Quat::to_homogeneous_matrix se3quat.h:270-278
    double R[9];
    quat_to_matrix(s.r, R);
    for (int i = 0; i < 3; i++) {
        T12[i * 4 + 0] = R[i * 3 + 0]; T12[i * 4 + 1] = R[i * 3 + 1]; T12[i * 4 + 2] = R[i * 3 + 2];
        T12[i * 4 + 3] = s.t[i];
    }
}

void se3_map(const SE3& s, const double p[3], double out[3]) {  // se3quat.h:217-220
    quat_rotate(s.r, p, out);
    out[0] += s.t[0]; out[1] += s.t[1]; out[2] += s.t[2];
}

SE3 se3_mul(const SE3& a, const SE3& b) {  // se3quat.h:104-110
    SE3 r = a;
    double rt[3];
    quat_rotate(a.r, b.t, rt);
    r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
    r.r = quat_mul(a.r, b.r);
    quat_normalize_rotation(r.r);
    return r;
}

void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

SE3 se3_exp(const double u[6]) {  // SE3Quat::exp se3quat.h:223-257 (rotation first, translation last)
    const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
    const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double Om2[9];
    mat3_mul(Om, Om, Om2);
    double R[9], V[9];
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) { R[i] = I[i] + Om[i] + Om2[i]; V[i] = R[i]; }
    } else {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
        const double c = (theta - std::sin(theta)) / std::pow(theta, 3);
        for (int i = 0; i < 9; i++) { R[i] = I[i] + a * Om[i] + b * Om2[i]; V[i] = I[i] + b * Om[i] + c * Om2[i]; }
    }
    SE3 s;
    s.r = quat_from_matrix(R);
    quat_normalize_rotation(s.r);
    for (int i = 0; i < 3; i++) s.t[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
    return s;
}

// 3x3 inverse by cofactors (Eigen's fixed-size inverse)
void inv3(const double m[9], double o[9]) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// dense LDL^T without pivoting (stand-in for Eigen::SimplicialLDLT, linear_solver_eigen.h:94-124); false on zero pivot
bool ldlt_solve(std::vector<double>& A, int n, const double* b, double* x) {
    std::vector<double> d(n);
    for (int j = 0; j < n; j++) {
        double dj = A[(size_t)j * n + j];
        for (int k = 0; k < j; k++) dj -= A[(size_t)j * n + k] * A[(size_t)j * n + k] * d[k];
        if (dj == 0.0 || !std::isfinite(dj)) return false;
        d[j] = dj;
        for (int i = j + 1; i < n; i++) {
            double v = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k] * d[k];
            A[(size_t)i * n + j] = v / dj;
        }
    }
    for (int i = 0; i < n; i++) {
        double v = b[i];
        for (int k = 0; k < i; k++) v -= A[(size_t)i * n + k] * x[k];
        x[i] = v;
    }
    for (int i = 0; i < n; i++) x[i] /= d[i];
    for (int i = n - 1; i >= 0; i--) {
        double v = x[i];
        for (int k = i + 1; k < n; k++) v -= A[(size_t)k * n + i] * x[k];
        x[i] = v;
    }
    return true;
}

// ---------------------------------------------------------------- the optimiser
struct Cam { double fx, fy, cx, cy; SE3 ext; double adj[36]; };

struct BA {
    int nP = 0, nL = 0, nE = 0;
    std::vector<SE3> pose;
    std::vector<uint8_t> fixed;
    std::vector<double> pt;          // 3 per point
    std::vector<int> ePose, ePt, eCam, eLevel;
    std::vector<uint8_t> eRobust;
    std::vector<double> obs, info, err;   // err: 2 per edge, as of the last computeActiveErrors (kept for inactive edges)
    std::vector<Cam> cam;
    double delta = 0, dsqr = 0;
    // active set (SparseOptimizer::initializeOptimization(level) sparse_optimizer.cpp:243-300)
    std::vector<int> activeE, poseIdx, ptIdx, freePoses, activePts;
    // system
    std::vector<double> Hpp, bp, Hll, bl, Hpl, x;   // Hpp: 36 per free pose; Hll: 9 per active point; Hpl: 18 per active edge (6x3)
    double lambda = 0, ni = 2;
    int nBad = 0;
    orc_ba_stats_t st{};
    const volatile uint8_t* stop = nullptr;

    bool terminate() const { return stop && *stop; }

    void project(int e, double pc[3]) const {  // EdgeSE3ProjectXYZ::computeError types_six_dof_expmap.cpp:109-114
        double pr[3];
        se3_map(pose[ePose[e]], &pt[3 * ePt[e]], pr);
        se3_map(cam[eCam[e]].ext, pr, pc);
    }

    void computeActiveErrors() {
        for (int e : activeE) {
            double pc[3];
            project(e, pc);
            const Cam& c = cam[eCam[e]];
            err[2 * e] = obs[2 * e] - (pc[0] / pc[2] * c.fx + c.cx);       // cam_project :163-169
            err[2 * e + 1] = obs[2 * e + 1] - (pc[1] / pc[2] * c.fy + c.cy);
        }
    }
    double chi2(int e) const { return (err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1]) * info[e]; }   // base_edge.h:58-61, Omega = I*invSigma2
    void robustify(double e, double rho[3]) const {  // RobustKernelHuber::robustify robust_kernel_impl.cpp:78-91
        if (e <= dsqr) { rho[0] = e; rho[1] = 1; rho[2] = 0; }
        else { const double s = std::sqrt(e); rho[0] = 2 * s * delta - dsqr; rho[1] = delta / s; rho[2] = -0.5 * rho[1] / e; }
    }
    double activeRobustChi2() const {  // sparse_optimizer.cpp:100-114
        double chi = 0;
        for (int e : activeE) {
            if (eRobust[e]) { double rho[3]; robustify(chi2(e), rho); chi += rho[0]; }
            else chi += chi2(e);
        }
        return chi;
    }

    void initializeOptimization(int level) {
        activeE.clear();
        for (int e = 0; e < nE; e++) if (eLevel[e] == level) activeE.push_back(e);
        std::vector<uint8_t> pa(nP, 0), la(nL, 0);
        for (int e : activeE) { pa[ePose[e]] = 1; la[ePt[e]] = 1; }
        // index mapping: non-marginalised (poses) then marginalised (points), ascending id (sparse_optimizer.cpp:166-190)
        poseIdx.assign(nP, -1); ptIdx.assign(nL, -1); freePoses.clear(); activePts.clear();
        for (int i = 0; i < nP; i++) if (pa[i] && !fixed[i]) { poseIdx[i] = (int)freePoses.size(); freePoses.push_back(i); }
        for (int i = 0; i < nL; i++) if (la[i]) { ptIdx[i] = (int)activePts.size(); activePts.push_back(i); }
    }

    // BlockSolver::buildSystem block_solver.hpp:502-560 : linearizeOplus + constructQuadraticForm per active edge
    void buildSystem() {
        const int K = (int)freePoses.size(), M = (int)activePts.size();
        Hpp.assign((size_t)36 * K, 0); bp.assign((size_t)6 * K, 0); Hll.assign((size_t)9 * M, 0); bl.assign((size_t)3 * M, 0);
        Hpl.assign((size_t)18 * activeE.size(), 0);
        for (size_t a = 0; a < activeE.size(); a++) {
            const int e = activeE[a];
            const Cam& c = cam[eCam[e]];
            double pc[3];
            project(e, pc);
            const double X = pc[0], Y = pc[1], Z = pc[2];
            // linearizeOplus types_six_dof_expmap.cpp:123-161
            const double tmp[6] = {c.fx, 0, -X / Z * c.fx, 0, c.fy, -Y / Z * c.fy};
            const double J3[18] = {0, Z, -Y, 1, 0, 0, -Z, 0, X, 0, 1, 0, Y, -X, 0, 0, 0, 1};
            double tJ[12];
            for (int i = 0; i < 2; i++)
                for (int j = 0; j < 6; j++) tJ[i * 6 + j] = (-1. / Z * tmp[i * 3]) * J3[j] + (-1. / Z * tmp[i * 3 + 1]) * J3[6 + j] + (-1. / Z * tmp[i * 3 + 2]) * J3[12 + j];
            double Jp[12];
            for (int i = 0; i < 2; i++)
                for (int j = 0; j < 6; j++) {
                    double s = 0;
                    for (int k = 0; k < 6; k++) s += tJ[i * 6 + k] * c.adj[k * 6 + j];
                    Jp[i * 6 + j] = s;
                }
            const SE3 T = se3_mul(c.ext, pose[ePose[e]]);
            double R[9];
            quat_to_matrix(T.r, R);
            double Jl[6];
            for (int i = 0; i < 2; i++)
                for (int j = 0; j < 3; j++) Jl[i * 3 + j] = (-1. / Z * tmp[i * 3]) * R[j] + (-1. / Z * tmp[i * 3 + 1]) * R[3 + j] + (-1. / Z * tmp[i * 3 + 2]) * R[6 + j];
            // constructQuadraticForm base_binary_edge.hpp:55-120 (robust: weighted Omega = rho1 * Omega, base_edge.h:96-102)
            double w = info[e], wr = 1;
            if (eRobust[e]) { double rho[3]; robustify(chi2(e), rho); wr = rho[1]; }
            const double W = wr * w;
            const double r0 = -w * err[2 * e] * wr, r1 = -w * err[2 * e + 1] * wr;   // omega_r * rho1
            const int li = ptIdx[ePt[e]], pi = poseIdx[ePose[e]];
            double* HL = &Hll[(size_t)9 * li];
            for (int i = 0; i < 3; i++) {
                bl[3 * li + i] += Jl[i] * r0 + Jl[3 + i] * r1;
                for (int j = 0; j < 3; j++) HL[i * 3 + j] += (Jl[i] * Jl[j] + Jl[3 + i] * Jl[3 + j]) * W;
            }
            if (pi >= 0) {
                double* HP = &Hpp[(size_t)36 * pi];
                for (int i = 0; i < 6; i++) {
                    bp[6 * pi + i] += Jp[i] * r0 + Jp[6 + i] * r1;
                    for (int j = 0; j < 6; j++) HP[i * 6 + j] += (Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j]) * W;
                }
                double* B = &Hpl[(size_t)18 * a];   // pose x landmark block (6x3)
                for (int i = 0; i < 6; i++)
                    for (int j = 0; j < 3; j++) B[i * 3 + j] += (Jp[i] * Jl[j] + Jp[6 + i] * Jl[3 + j]) * W;
            }
        }
    }

    // BlockSolver::setLambda + solve (Schur) + restoreDiagonal, block_solver.hpp:354-486,564-604
    bool solve() {
        const int K = (int)freePoses.size(), M = (int)activePts.size(), n = 6 * K;
        x.assign((size_t)n + 3 * M, 0);
        std::vector<double> Hs((size_t)n * n, 0), coeff(n, 0), Dinv((size_t)9 * M), db((size_t)3 * M);
        for (int k = 0; k < K; k++)
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < 6; j++) Hs[(size_t)(6 * k + i) * n + 6 * k + j] = Hpp[(size_t)36 * k + i * 6 + j] + (i == j ? lambda : 0);
        for (int l = 0; l < M; l++) {
            double D[9];
            for (int i = 0; i < 9; i++) D[i] = Hll[(size_t)9 * l + i];
            D[0] += lambda; D[4] += lambda; D[8] += lambda;
            inv3(D, &Dinv[(size_t)9 * l]);
            for (int i = 0; i < 3; i++) db[3 * l + i] = Dinv[9 * l + i * 3] * bl[3 * l] + Dinv[9 * l + i * 3 + 1] * bl[3 * l + 1] + Dinv[9 * l + i * 3 + 2] * bl[3 * l + 2];
        }
        // edges grouped by landmark
        std::vector<std::vector<int>> byPt(M);
        for (size_t a = 0; a < activeE.size(); a++) {
            const int e = activeE[a];
            if (poseIdx[ePose[e]] >= 0) byPt[ptIdx[ePt[e]]].push_back((int)a);
        }
        for (int l = 0; l < M; l++) {
            const double* Di = &Dinv[(size_t)9 * l];
            for (int a1 : byPt[l]) {
                const int i1 = poseIdx[ePose[activeE[a1]]];
                const double* Bi = &Hpl[(size_t)18 * a1];
                double BD[18];
                for (int i = 0; i < 6; i++)
                    for (int j = 0; j < 3; j++) BD[i * 3 + j] = Bi[i * 3] * Di[j] + Bi[i * 3 + 1] * Di[3 + j] + Bi[i * 3 + 2] * Di[6 + j];
                for (int i = 0; i < 6; i++) coeff[6 * i1 + i] += Bi[i * 3] * db[3 * l] + Bi[i * 3 + 1] * db[3 * l + 1] + Bi[i * 3 + 2] * db[3 * l + 2];
                for (int a2 : byPt[l]) {
                    const int i2 = poseIdx[ePose[activeE[a2]]];
                    if (i2 < i1) continue;   // upper block triangle (block_solver.hpp:418-431)
                    const double* Bj = &Hpl[(size_t)18 * a2];
                    for (int i = 0; i < 6; i++)
                        for (int j = 0; j < 6; j++)
                            Hs[(size_t)(6 * i1 + i) * n + 6 * i2 + j] -= BD[i * 3] * Bj[j * 3] + BD[i * 3 + 1] * Bj[j * 3 + 1] + BD[i * 3 + 2] * Bj[j * 3 + 2];
                }
            }
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++) Hs[(size_t)i * n + j] = Hs[(size_t)j * n + i];   // the solver reads the upper triangle
        std::vector<double> bs(n);
        for (int i = 0; i < n; i++) bs[i] = bp[i] - coeff[i];
        if (n > 0 && !ldlt_solve(Hs, n, bs.data(), x.data())) return false;
        // landmarks: xl = Dinv * (bl - B^T xp)   (block_solver.hpp:461-481)
        std::vector<double> cl(bl);
        for (int l = 0; l < M; l++)
            for (int a1 : byPt[l]) {
                const int i1 = poseIdx[ePose[activeE[a1]]];
                const double* Bi = &Hpl[(size_t)18 * a1];
                for (int j = 0; j < 3; j++)
                    for (int i = 0; i < 6; i++) cl[3 * l + j] -= Bi[i * 3 + j] * x[6 * i1 + i];
            }
        for (int l = 0; l < M; l++)
            for (int i = 0; i < 3; i++) x[n + 3 * l + i] = Dinv[9 * l + i * 3] * cl[3 * l] + Dinv[9 * l + i * 3 + 1] * cl[3 * l + 1] + Dinv[9 * l + i * 3 + 2] * cl[3 * l + 2];
        return true;
    }

    void update() {  // SparseOptimizer::update + oplusImpl (types_six_dof_expmap.h:73-76, types_sba.h:53-57)
        const int K = (int)freePoses.size(), M = (int)activePts.size();
        for (int k = 0; k < K; k++) pose[freePoses[k]] = se3_mul(se3_exp(&x[6 * k]), pose[freePoses[k]]);
        for (int l = 0; l < M; l++)
            for (int i = 0; i < 3; i++) pt[3 * activePts[l] + i] += x[6 * K + 3 * l + i];
    }

    enum Result { OK, Terminate, Fail };

    // OptimizationAlgorithmLevenberg::solve optimization_algorithm_levenberg.cpp:61-164
    Result lmIteration(int iteration) {
        computeActiveErrors();
        double currentChi = activeRobustChi2(), tempChi = currentChi;
        const double iniChi = currentChi;
        buildSystem();
        const int K = (int)freePoses.size(), M = (int)activePts.size();
        if (iteration == 0) {  // computeLambdaInit :166-180
            double maxDiagonal = 0;
            for (int k = 0; k < K; k++) for (int j = 0; j < 6; j++) maxDiagonal = std::max(std::fabs(Hpp[(size_t)36 * k + j * 7]), maxDiagonal);
            for (int l = 0; l < M; l++) for (int j = 0; j < 3; j++) maxDiagonal = std::max(std::fabs(Hll[(size_t)9 * l + j * 4]), maxDiagonal);
            lambda = 1e-5 * maxDiagonal;
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const std::vector<SE3> savedPose = pose;     // push()
            const std::vector<double> savedPt = pt;
            const bool ok2 = solve();
            update();
            computeActiveErrors();
            tempChi = activeRobustChi2();
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = 0;   // computeScale :182-189
            for (int j = 0; j < 6 * K; j++) scale += x[j] * (lambda * x[j] + bp[j]);
            for (int j = 0; j < 3 * M; j++) scale += x[6 * K + j] * (lambda * x[6 * K + j] + bl[j]);
            scale += 1e-3;
            rho /= scale;
            st.trials++;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow(2 * rho - 1, 3);
                alpha = std::min(alpha, 2. / 3.);
                const double scaleFactor = std::max(1. / 3., alpha);
                lambda *= scaleFactor;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                pose = savedPose;   // pop()
                pt = savedPt;
            }
            qmax++;
        } while (rho < 0 && qmax < 10 && !terminate());
        st.final_chi2 = currentChi;
        st.final_lambda = lambda;
        if (qmax == 10 || rho == 0) return Terminate;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) return Terminate;
        return OK;
    }

    int optimize(int iterations) {  // SparseOptimizer::optimize sparse_optimizer.cpp:354-419
        if (freePoses.empty() && activePts.empty()) return -1;
        int n = 0;
        bool ok = true;
        for (int i = 0; i < iterations && !terminate() && ok; i++) {
            ok = lmIteration(i) == OK;
            n++;
            st.iterations++;
        }
        return n;
    }

    bool depthPositive(int e) const { double pc[3]; project(e, pc); return pc[2] > 0.0; }
};

// ---------------------------------------------------------------- pose-only optimisation (Optimizer::PoseOptimization)
// One VertexSE3Expmap, unary EdgeSE3ProjectXYZOnlyPose edges (types_six_dof_expmap.cpp:200-255), BlockSolver_6_3 over
// LinearSolverDense (6x6 Eigen::LDLT; restated as LDL^T without pivoting + positivity check), the same LM as above.
struct PoseOnly {
    SE3 pose;
    int nE = 0;
    std::vector<double> Xw, obs, info, err;
    std::vector<int> eCam, level;
    std::vector<Cam> cam;
    bool robust = true;
    double delta = 0, dsqr = 0, lambda = 0, ni = 2;
    int nBad = 0, trials = 0, iterations = 0;
    double H[36], b[6], x[6];

    void project(int e, double pc[3]) const { double pr[3]; se3_map(pose, &Xw[3 * e], pr); se3_map(cam[eCam[e]].ext, pr, pc); }
    void computeError(int e) {
        double pc[3];
        project(e, pc);
        const Cam& c = cam[eCam[e]];
        err[2 * e] = obs[2 * e] - (pc[0] / pc[2] * c.fx + c.cx);
        err[2 * e + 1] = obs[2 * e + 1] - (pc[1] / pc[2] * c.fy + c.cy);
    }
    double chi2(int e) const { return (err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1]) * info[e]; }
    void computeActiveErrors() { for (int e = 0; e < nE; e++) if (level[e] == 0) computeError(e); }
    double activeRobustChi2() const {
        double chi = 0;
        for (int e = 0; e < nE; e++) {
            if (level[e]) continue;
            const double c2 = chi2(e);
            chi += (robust && c2 > dsqr) ? 2 * std::sqrt(c2) * delta - dsqr : c2;
        }
        return chi;
    }
    void buildSystem() {
        for (double& v : H) v = 0;
        for (double& v : b) v = 0;
        for (int e = 0; e < nE; e++) {
            if (level[e]) continue;
            const Cam& c = cam[eCam[e]];
            double pc[3];
            project(e, pc);
            const double X = pc[0], Y = pc[1], Z = pc[2];
            const double tmp[6] = {c.fx, 0, -X / Z * c.fx, 0, c.fy, -Y / Z * c.fy};
            const double J3[18] = {0, Z, -Y, 1, 0, 0, -Z, 0, X, 0, 1, 0, Y, -X, 0, 0, 0, 1};
            double tJ[12], Jp[12];
            for (int i = 0; i < 2; i++)
                for (int j = 0; j < 6; j++) tJ[i * 6 + j] = (-1. / Z * tmp[i * 3]) * J3[j] + (-1. / Z * tmp[i * 3 + 1]) * J3[6 + j] + (-1. / Z * tmp[i * 3 + 2]) * J3[12 + j];
            for (int i = 0; i < 2; i++)
                for (int j = 0; j < 6; j++) {
                    double s = 0;
                    for (int k = 0; k < 6; k++) s += tJ[i * 6 + k] * c.adj[k * 6 + j];
                    Jp[i * 6 + j] = s;
                }
            const double w = info[e];
            double wr = 1;
            if (robust) { const double c2 = chi2(e); if (c2 > dsqr) wr = delta / std::sqrt(c2); }
            const double W = wr * w, r0 = -w * err[2 * e] * wr, r1 = -w * err[2 * e + 1] * wr;
            for (int i = 0; i < 6; i++) {
                b[i] += Jp[i] * r0 + Jp[6 + i] * r1;
                for (int j = 0; j < 6; j++) H[i * 6 + j] += (Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j]) * W;
            }
        }
    }
    bool solve() {
        std::vector<double> A(36);
        for (int i = 0; i < 36; i++) A[i] = H[i] + ((i % 7 == 0) ? lambda : 0.0);
        for (double& v : x) v = 0;
        return ldlt_solve(A, 6, b, x);
    }
    bool lmIteration(int iteration) {
        computeActiveErrors();
        double currentChi = activeRobustChi2(), tempChi = currentChi;
        const double iniChi = currentChi;
        buildSystem();
        if (iteration == 0) {
            double md = 0;
            for (int j = 0; j < 6; j++) md = std::max(std::fabs(H[j * 7]), md);
            lambda = 1e-5 * md; ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const SE3 saved = pose;
            const bool ok2 = solve();
            pose = se3_mul(se3_exp(x), pose);
            computeActiveErrors();
            tempChi = activeRobustChi2();
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
            scale += 1e-3;
            rho /= scale;
            trials++;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow(2 * rho - 1, 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                pose = saved;
            }
            qmax++;
        } while (rho < 0 && qmax < 10);
        if (qmax == 10 || rho == 0) return false;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        return nBad < 3;
    }
    void optimize(int its) {
        bool any = false;
        for (int e = 0; e < nE; e++) any |= level[e] == 0;
        if (!any) return;                                  // "0 vertices to optimize"
        bool ok = true;
        for (int i = 0; i < its && ok; i++) { ok = lmIteration(i); iterations++; }
    }
};

void load(BA& b, const orc_ba_problem_t* p, double huber_delta) {
    b.nP = p->n_poses; b.nL = p->n_points; b.nE = p->n_edges;
    b.pose.resize(b.nP); b.fixed.assign(p->pose_fixed, p->pose_fixed + b.nP);
    for (int i = 0; i < b.nP; i++) b.pose[i] = se3_from_Rt(p->poses + 12 * i);
    b.pt.assign(p->points, p->points + 3 * (size_t)b.nL);
    b.ePose.assign(p->edge_pose, p->edge_pose + b.nE);
    b.ePt.assign(p->edge_point, p->edge_point + b.nE);
    b.eCam.assign(p->edge_cam, p->edge_cam + b.nE);
    b.eLevel.assign(b.nE, 0);
    b.eRobust.assign(b.nE, huber_delta > 0 ? 1 : 0);
    b.obs.assign(p->edge_obs, p->edge_obs + 2 * (size_t)b.nE);
    b.info.assign(p->edge_inv_sigma2, p->edge_inv_sigma2 + b.nE);
    b.err.assign(2 * (size_t)b.nE, 0);
    b.cam.resize(p->n_cams);
    for (int c = 0; c < p->n_cams; c++) {
        Cam& C = b.cam[c];
        C.fx = p->cam_K[4 * c]; C.fy = p->cam_K[4 * c + 1]; C.cx = p->cam_K[4 * c + 2]; C.cy = p->cam_K[4 * c + 3];
        C.ext = se3_from_Rt(p->cam_ext + 12 * c);
        std::memcpy(C.adj, p->cam_adj + 36 * c, sizeof(C.adj));
    }
    b.delta = huber_delta; b.dsqr = huber_delta * huber_delta;
}

void store(const BA& b, double* poses_out, double* points_out) {
    if (poses_out) for (int i = 0; i < b.nP; i++) se3_to_Rt(b.pose[i], poses_out + 12 * i);
    if (points_out) std::memcpy(points_out, b.pt.data(), sizeof(double) * 3 * (size_t)b.nL);
}

}  // namespace

extern "C" {

// Optimizer::LocalBundleAdjustment src/Optimizer.cc:582-660 (graph is given flattened; see orb_oracle.h)
int orc_local_ba(const orc_ba_problem_t* p, int its1, int its2, double huber_delta, double chi2_th, const volatile uint8_t* stop,
                 double* poses_out, double* points_out, uint8_t* edge_outlier, orc_ba_stats_t* stats) {
    if (!p) return -1;
    BA b;
    load(b, p, huber_delta);
    b.stop = stop;
    if (stop && *stop) { store(b, poses_out, points_out); return -5; }   // :582-584
    b.initializeOptimization(0);
    b.computeActiveErrors();
    b.st.initial_chi2 = b.activeRobustChi2();
    b.optimize(its1);
    const bool doMore = !(stop && *stop);
    if (doMore) {
        for (int e = 0; e < b.nE; e++) {   // :598-613
            if (b.chi2(e) > chi2_th || !b.depthPositive(e)) b.eLevel[e] = 1;
            b.eRobust[e] = 0;
        }
        b.initializeOptimization(0);
        b.optimize(its2);
    }
    int nout = 0;
    for (int e = 0; e < b.nE; e++) {       // :641-655
        const bool out = b.chi2(e) > chi2_th || !b.depthPositive(e);
        if (edge_outlier) edge_outlier[e] = out;
        nout += out;
    }
    b.st.outliers = nout;
    store(b, poses_out, points_out);
    if (stats) *stats = b.st;
    return 0;
}

// Optimizer::BundleAdjustment src/Optimizer.cc:70-248 (single optimize(nIterations); Huber thHuber2D = sqrt(3.99) if robust, :108; no outlier pass)
int orc_global_ba(const orc_ba_problem_t* p, int iterations, double huber_delta, const volatile uint8_t* stop,
                  double* poses_out, double* points_out, orc_ba_stats_t* stats) {
    if (!p) return -1;
    BA b;
    load(b, p, huber_delta);
    b.stop = stop;
    b.initializeOptimization(0);
    b.computeActiveErrors();
    b.st.initial_chi2 = b.activeRobustChi2();
    b.optimize(iterations);
    store(b, poses_out, points_out);
    if (stats) *stats = b.st;
    return 0;
}

// one buildSystem at the given state: exposes H / b for the numpy cross-check (tests only)
int orc_ba_normal_equations(const orc_ba_problem_t* p, double huber_delta, double* Hpp /*[K][36]*/, double* bp, double* Hll /*[M][9]*/,
                            double* bl, double* Hpl /*[E][18]*/, double* chi2) {
    BA b;
    load(b, p, huber_delta);
    b.initializeOptimization(0);
    b.computeActiveErrors();
    if (chi2) *chi2 = b.activeRobustChi2();
    b.buildSystem();
    std::memcpy(Hpp, b.Hpp.data(), b.Hpp.size() * 8); std::memcpy(bp, b.bp.data(), b.bp.size() * 8);
    std::memcpy(Hll, b.Hll.data(), b.Hll.size() * 8); std::memcpy(bl, b.bl.data(), b.bl.size() * 8);
    std::memcpy(Hpl, b.Hpl.data(), b.Hpl.size() * 8);
    return (int)b.freePoses.size();
}


// Optimizer::PoseOptimization(pFrame)  src/Optimizer.cc:250-405.  One edge per keypoint that holds a map point (in keypoint
// order).  Returns nInitialCorrespondences - nBad (0 when fewer than 3 correspondences: nothing is touched then);
// outlier[e] = pFrame->mvbOutlier of the keypoint, pose_out = the pose SetPose() receives.
int orc_pose_optimization(const double* pose12, int n, const double* Xw, const double* obs, const double* inv_sigma2, const int32_t* cam,
                          int n_cams, const double* cam_K, const double* cam_ext, const double* cam_adj, double* pose_out, uint8_t* outlier,
                          int32_t* lm_counts /* [2] iterations, trials; may be NULL */) {
    for (int i = 0; i < 12; i++) pose_out[i] = pose12[i];
    for (int e = 0; e < n; e++) outlier[e] = 0;
    if (lm_counts) { lm_counts[0] = 0; lm_counts[1] = 0; }
    if (n < 3) return 0;
    PoseOnly P;
    P.nE = n;
    P.Xw.assign(Xw, Xw + 3 * (size_t)n); P.obs.assign(obs, obs + 2 * (size_t)n); P.info.assign(inv_sigma2, inv_sigma2 + n);
    P.err.assign(2 * (size_t)n, 0); P.eCam.assign(cam, cam + n); P.level.assign(n, 0);
    P.cam.resize(n_cams);
    for (int c = 0; c < n_cams; c++) {
        Cam& C = P.cam[c];
        C.fx = cam_K[4 * c]; C.fy = cam_K[4 * c + 1]; C.cx = cam_K[4 * c + 2]; C.cy = cam_K[4 * c + 3];
        C.ext = se3_from_Rt(cam_ext + 12 * c);
        for (int i = 0; i < 36; i++) C.adj[i] = cam_adj[36 * c + i];
    }
    P.delta = (double)(float)std::sqrt(5.991); P.dsqr = P.delta * P.delta;   // const float deltaMono = sqrt(5.991)
    const SE3 initial = se3_from_Rt(pose12);
    P.pose = initial;
    const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f};
    int nBad = 0;
    for (int it = 0; it < 4; it++) {
        P.pose = initial;                                   // vSE3->setEstimate(toSE3Quat(pFrame->mTcw)): every round restarts
        P.optimize(10);
        nBad = 0;
        for (int e = 0; e < n; e++) {
            if (outlier[e]) P.computeError(e);
            const float chi2 = (float)P.chi2(e);
            if (chi2 > chi2Mono[it]) { outlier[e] = 1; P.level[e] = 1; nBad++; }
            else { outlier[e] = 0; P.level[e] = 0; }
        }
        if (it == 2) P.robust = false;                      // e->setRobustKernel(0)
        if (n < 10) break;                                  // optimizer.edges().size() < 10
    }
    se3_to_Rt(P.pose, pose_out);
    if (lm_counts) { lm_counts[0] = P.iterations; lm_counts[1] = P.trials; }
    return n - nBad;
}

}  // extern "C"
