// orb_geometry.h -- host-side derivation of everything ORBextractor computes once per (parameters, image size):
// scale tables and per-level quotas (src/ORBextractor.cc:415-446), umax (:454-469), level sizes (:1111-1112),
// the 30-px cell grid (:773-806), quadtree roots (:543-563) and the fixed-point bilinear tables of cv::resize
// (OpenCV imgproc, INTER_LINEAR 8U: 11-bit coefficients).  Host only; results are uploaded to the device.
#pragma once
#include <stdint.h>

#include <cmath>
#include <vector>

namespace orbgeo {

inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_floor(float v) { int i = (int)v; return i - (i > v); }
inline int cv_ceil(float v) { int i = (int)v; return i + (i < v); }

struct Level {
    int w = 0, h = 0;          // level image size
    int pitch = 0;             // row pitch in bytes of the device storage (multiple of 16)
    float scale = 1.f;         // mvScaleFactor[l]
    int patch_size = 31;       // (int)(31*scale): cv::KeyPoint::size
    int quota = 0;             // mnFeaturesPerLevel[l]
    // cell grid, coordinates relative to (16,16)
    int width = 0, height = 0; // maxBorder - minBorder
    int nCols = 0, nRows = 0, wCell = 0, hCell = 0;
    int nColsEff = 0, nRowsEff = 0;  // cells that are not skipped by the `continue`s at :794-795,:803-804
    bool valid = false;        // at least one cell
    // quadtree roots
    int nIni = 0;
    float hX = 0.f;
    // resize tables (level l from level l-1); empty for level 0
    std::vector<int32_t> xofs, yofs;       // source column / row
    std::vector<int16_t> ialpha, ibeta;    // 2 coefficients per destination column / row
};

struct Geometry {
    int nlevels = 0, nfeatures = 0, iniTh = 0, minTh = 0;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> umax;     // 16 entries
    std::vector<Level> lv;
    int max_keypoints = 0;     // per image output capacity
};

inline void resize_tables(int sw, int sh, int dw, int dh, Level& L) {
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    L.xofs.resize(dw); L.ialpha.resize(2 * dw); L.yofs.resize(dh); L.ibeta.resize(2 * dh);
    auto sat = [](int v) { return (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); };
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        L.xofs[dx] = sx;
        L.ialpha[2 * dx] = sat(cv_round((1.f - fx) * 2048));
        L.ialpha[2 * dx + 1] = sat(cv_round(fx * 2048));
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        L.yofs[dy] = sy;
        L.ibeta[2 * dy] = sat(cv_round((1.f - fy) * 2048));
        L.ibeta[2 * dy + 1] = sat(cv_round(fy * 2048));
    }
}

inline Geometry make_geometry(int W, int H, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    Geometry g;
    g.nlevels = nlevels; g.nfeatures = nfeatures; g.iniTh = iniTh; g.minTh = minTh;
    g.scale.assign(nlevels, 1.f); g.sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; i++) {
        g.scale[i] = g.scale[i - 1] * scaleFactor;
        g.sigma2[i] = g.scale[i] * g.scale[i];
    }
    g.inv_scale.resize(nlevels); g.inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) { g.inv_scale[i] = 1.0f / g.scale[i]; g.inv_sigma2[i] = 1.0f / g.sigma2[i]; }
    std::vector<int> quota(nlevels);
    const float factor = 1.0f / scaleFactor;
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        quota[l] = cv_round(nDesired);
        sum += quota[l];
        nDesired *= factor;
    }
    quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    g.umax.assign(16, 0);
    int v, v0, vmax = cv_floor(15 * std::sqrt(2.f) / 2 + 1);
    const int vmin = cv_ceil(15 * std::sqrt(2.f) / 2);
    for (v = 0; v <= vmax; ++v) g.umax[v] = cv_round(std::sqrt(225.0 - v * v));
    for (v = 15, v0 = 0; v >= vmin; --v) {
        while (g.umax[v0] == g.umax[v0 + 1]) ++v0;
        g.umax[v] = v0;
        ++v0;
    }
    g.lv.resize(nlevels);
    g.max_keypoints = 0;
    for (int l = 0; l < nlevels; l++) {
        Level& L = g.lv[l];
        L.w = cv_round((float)W * g.inv_scale[l]);
        L.h = cv_round((float)H * g.inv_scale[l]);
        L.pitch = (L.w + 15) & ~15;
        L.scale = g.scale[l];
        L.patch_size = (int)(31 * g.scale[l]);
        L.quota = quota[l];
        if (l > 0 && L.w > 0 && L.h > 0) resize_tables(g.lv[l - 1].w, g.lv[l - 1].h, L.w, L.h, L);
        const int minB = 16, maxBX = L.w - 16, maxBY = L.h - 16;
        L.width = maxBX - minB; L.height = maxBY - minB;
        const float width = (float)L.width, height = (float)L.height;
        L.nCols = L.width > 0 ? (int)(width / 30.f) : 0;
        L.nRows = L.height > 0 ? (int)(height / 30.f) : 0;
        L.valid = L.nCols >= 1 && L.nRows >= 1;
        if (L.valid) {
            L.wCell = (int)std::ceil(width / L.nCols);
            L.hCell = (int)std::ceil(height / L.nRows);
            for (int j = 0; j < L.nCols; j++) if (!((float)(minB + j * L.wCell) >= (float)(maxBX - 6))) L.nColsEff = j + 1;
            for (int i = 0; i < L.nRows; i++) if (!((float)(minB + i * L.hCell) >= (float)(maxBY - 3))) L.nRowsEff = i + 1;
            L.nIni = (int)std::round((float)L.width / (float)L.height);
            if (L.nIni < 1) L.valid = false;   // reference indexes an empty vector here; rejected at create()
            else L.hX = (float)L.width / L.nIni;
        }
        // the quadtree stops at >= quota nodes but a split adds up to 3: <= quota + 2 (and >= 4*nIni after sweep 1)
        int cap = L.quota + 2;
        if (cap < 4 * L.nIni) cap = 4 * L.nIni;
        g.max_keypoints += L.valid ? cap : 0;
    }
    return g;
}

}  // namespace orbgeo
