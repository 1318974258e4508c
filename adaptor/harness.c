/* adaptor/harness.c -- plain C99 program that links liborbslam2_dualcam_b200.so and drives the two ends of the hot path natively
 * (no Python, no ctypes): orbx_extract on a synthetic 640x480 image and orbba_local on a small synthetic dual-camera window.
 *
 *   gcc -std=c99 -O1 -Iinclude adaptor/harness.c -o adaptor/_build/harness_c -Lorb-slam2-dualcam_b200/lib -lorbslam2_dualcam_b200 -lm
 *
 * Exit code 0 and a line "harness ok ..." on success; needs an sm_100 GPU (the library has no CPU fallback). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orbslam2_dualcam_b200.h"

static uint32_t rng_state = 12345u;
static double rnd(void) { rng_state = rng_state * 1664525u + 1013904223u; return (rng_state >> 8) / 16777216.0; }   /* [0, 1) */

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_ != ORB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, orb_last_error()); return 1; } \
    } while (0)

static int run_extract(void) {
    const int W = 640, H = 480;
    uint8_t* img = (uint8_t*)malloc((size_t)W * H);
    /* blocks of random grey levels with a soft ramp: plenty of FAST corners at block junctions */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) img[y * W + x] = 0;
    for (int b = 0; b < 500; b++) {
        const int x0 = (int)(rnd() * W), y0 = (int)(rnd() * H), w = 8 + (int)(rnd() * 70), h = 8 + (int)(rnd() * 70), g = (int)(rnd() * 256);
        for (int y = y0; y < y0 + h && y < H; y++)
            for (int x = x0; x < x0 + w && x < W; x++) img[y * W + x] = (uint8_t)((g + (x - x0)) & 255);
    }
    orbx_t* ex = NULL;
    CHECK(orbx_create(&ex, 0, W, H, 1, 1, 1000, 1.2f, 8, 20, 7));
    const int cap = orbx_max_keypoints(ex);
    orb_keypoint_t* kps = (orb_keypoint_t*)malloc(sizeof(orb_keypoint_t) * (size_t)cap);
    uint8_t* desc = (uint8_t*)malloc((size_t)cap * 32);
    int32_t n = 0;
    CHECK(orbx_extract(ex, img, 1, (size_t)W, kps, desc, &n, cap));
    unsigned long long sum = 0;
    for (int i = 0; i < n * 32; i++) sum = sum * 131u + desc[i];
    int in_range = 1;
    for (int i = 0; i < n; i++) in_range &= kps[i].x >= 0 && kps[i].x < W && kps[i].y >= 0 && kps[i].y < H && kps[i].octave >= 0 && kps[i].octave < 8;
    printf("extract: %d keypoints (capacity %d), descriptor checksum %016llx\n", n, cap, sum);
    orbx_destroy(ex);
    free(img); free(kps); free(desc);
    return (n > 200 && n <= cap && in_range) ? 0 : 2;
}

/* row-major 3x4 [R | t] of a rotation about y by `a` and a translation */
static void pose_y(double a, double tx, double ty, double tz, double* T) {
    const double c = cos(a), s = sin(a);
    const double R[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
    for (int r = 0; r < 3; r++) { for (int q = 0; q < 3; q++) T[4 * r + q] = R[3 * r + q]; }
    T[3] = tx; T[7] = ty; T[11] = tz;
}
static void apply(const double* T, const double* X, double* o) {
    for (int r = 0; r < 3; r++) o[r] = T[4 * r] * X[0] + T[4 * r + 1] * X[1] + T[4 * r + 2] * X[2] + T[4 * r + 3];
}

static int run_ba(void) {
    enum { NP = 6, NL = 240, NC = 2 };
    const double K[NC][4] = {{520, 520, 320, 240}, {515, 515, 318, 242}};
    double ext[NC][12], adj[NC][36];
    pose_y(0.0, 0, 0, 0, ext[0]);
    pose_y(0.35, -0.12, 0.0, 0.02, ext[1]);
    for (int c = 0; c < NC; c++) {          /* Cameras::setExtrinsics (src/Cameras.cc:17-40): [R 0; skew(t) R, R], lower-left block as the caller holds it (0 here) */
        memset(adj[c], 0, sizeof(adj[c]));
        for (int r = 0; r < 3; r++)
            for (int q = 0; q < 3; q++) { adj[c][6 * r + q] = ext[c][4 * r + q]; adj[c][6 * (r + 3) + q + 3] = ext[c][4 * r + q]; }
        const double t[3] = {ext[c][3], ext[c][7], ext[c][11]};
        const double S[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
        for (int r = 0; r < 3; r++)
            for (int q = 0; q < 3; q++) { double v = 0; for (int k = 0; k < 3; k++) v += S[3 * r + k] * ext[c][4 * k + q]; adj[c][6 * r + q + 3] = v; }
    }
    double gt[NP][12], poses[NP][12], pts[NL][3], gtp[NL][3];
    uint8_t fixed[NP] = {1, 0, 0, 0, 0, 0};
    for (int i = 0; i < NP; i++) pose_y(0.02 * i, -0.25 * i, 0.01 * i, 0.05 * i, gt[i]);
    for (int l = 0; l < NL; l++) { gtp[l][0] = -1.5 + 4.5 * rnd(); gtp[l][1] = -1.5 + 3.0 * rnd(); gtp[l][2] = 4.0 + 8.0 * rnd(); }
    int32_t ep[NP * NL], el[NP * NL], ec[NP * NL];
    double obs[NP * NL][2], info[NP * NL];
    int nE = 0;
    for (int l = 0; l < NL; l++)
        for (int i = 0; i < NP; i++) {
            double pr[3], pc[3];
            apply(gt[i], gtp[l], pr);
            for (int c = 0; c < NC; c++) {
                apply(ext[c], pr, pc);
                const double u = K[c][0] * pc[0] / pc[2] + K[c][2], v = K[c][1] * pc[1] / pc[2] + K[c][3];
                if (pc[2] > 0.2 && u >= 0 && u < 640 && v >= 0 && v < 480) {      /* one observation per (landmark, key frame) */
                    const int oct = (int)(rnd() * 8);
                    const double sc = pow(1.2, oct);
                    ep[nE] = i; el[nE] = l; ec[nE] = c;
                    obs[nE][0] = (float)(u + (rnd() - 0.5) * sc); obs[nE][1] = (float)(v + (rnd() - 0.5) * sc);
                    info[nE] = 1.0 / (sc * sc);
                    nE++;
                    break;
                }
            }
        }
    for (int i = 0; i < NP; i++) {
        memcpy(poses[i], gt[i], sizeof(gt[i]));
        if (i > 0) { poses[i][3] += 0.03 * (rnd() - 0.5); poses[i][7] += 0.03 * (rnd() - 0.5); poses[i][11] += 0.03 * (rnd() - 0.5); }
    }
    for (int l = 0; l < NL; l++) for (int k = 0; k < 3; k++) pts[l][k] = gtp[l][k] + 0.05 * (rnd() - 0.5);
    orbba_problem_t P;
    P.n_poses = NP; P.n_points = NL; P.n_edges = nE; P.n_cams = NC;
    P.poses = &poses[0][0]; P.pose_fixed = fixed; P.points = &pts[0][0];
    P.edge_pose = ep; P.edge_point = el; P.edge_cam = ec; P.edge_obs = &obs[0][0]; P.edge_inv_sigma2 = info;
    P.cam_K = &K[0][0]; P.cam_ext = &ext[0][0]; P.cam_adj = &adj[0][0];
    orbba_t* ba = NULL;
    CHECK(orbba_create(&ba, 0, 1));
    static double pout[NP][12], lout[NL][3];
    static uint8_t outl[NP * NL];
    orbba_stats_t st;
    CHECK(orbba_local(ba, &P, 5, 10, sqrt(5.991), 5.991, NULL, &pout[0][0], &lout[0][0], outl, &st));
    double worst = 0;
    for (int i = 0; i < NP; i++)
        for (int k = 0; k < 12; k++) { const double d = fabs(pout[i][k] - gt[i][k]); if (d > worst) worst = d; }
    printf("localBA: %d edges, chi2 %.1f -> %.1f, %d iterations, %d trials, %d outliers, worst pose entry error vs ground truth %.2e\n", nE, st.initial_chi2,
           st.final_chi2, st.iterations, st.trials, st.outliers, worst);
    orbba_destroy(ba);
    return (st.final_chi2 < 0.2 * st.initial_chi2 && worst < 0.02 && st.iterations > 0) ? 0 : 3;
}

int main(void) {
    int rc = run_extract();
    if (rc) { fprintf(stderr, "harness: extract failed (%d)\n", rc); return rc; }
    rc = run_ba();
    if (rc) { fprintf(stderr, "harness: localBA failed (%d)\n", rc); return rc; }
    printf("harness ok: orbx_extract and orbba_local called natively from C\n");
    return 0;
}
