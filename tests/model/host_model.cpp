// host_model.cpp -- CPU emulation of the GPU FORMULATION of the extractor stages (test support, built by tests/).
// It runs the same orb_core.h / orb_geometry.h code the sm_100a kernels use, with the block-parallel parts
// replaced by serial loops, so that the reformulations (closed-form FAST score, cell-masked NMS, order keys,
// level-synchronous quadtree, glibc sincosf restatement) can be checked against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../orb-slam2-dualcam_b200/csrc/orb_core.h"
#include "../../orb-slam2-dualcam_b200/csrc/orb_geometry.h"

using namespace orbcore;

extern "C" {

// geometry query: fills ints {w,h,width,height,nCols,nRows,wCell,hCell,nColsEff,nRowsEff,nIni,quota,valid} per level
int hm_geometry(int W, int H, int nfeatures, float sf, int nlevels, int iniTh, int minTh, int32_t* out13, float* hX, int* max_kp) {
    orbgeo::Geometry g = orbgeo::make_geometry(W, H, nfeatures, sf, nlevels, iniTh, minTh);
    for (int l = 0; l < nlevels; l++) {
        const orbgeo::Level& L = g.lv[l];
        int32_t* o = out13 + 13 * l;
        o[0] = L.w; o[1] = L.h; o[2] = L.width; o[3] = L.height; o[4] = L.nCols; o[5] = L.nRows; o[6] = L.wCell; o[7] = L.hCell;
        o[8] = L.nColsEff; o[9] = L.nRowsEff; o[10] = L.nIni; o[11] = L.quota; o[12] = L.valid;
        hX[l] = L.hX;
    }
    *max_kp = g.max_keypoints;
    return 0;
}

// FAST + cell-local NMS + per-cell threshold fallback on one dense level image (GPU formulation).
// Returns number of candidates; out = packed candidates (unordered).
int hm_fast_level(const uint8_t* img, int w, int h, int iniTh, int minTh, uint32_t* out, int cap) {
    orbgeo::Geometry g = orbgeo::make_geometry(w, h, 1000, 1.2f, 1, iniTh, minTh);
    const orbgeo::Level& L = g.lv[0];
    if (!L.valid) return 0;
    const int rdx[16] = ORB_RING_DX, rdy[16] = ORB_RING_DY;
    iniTh = std::min(std::max(iniTh, 0), 255); minTh = std::min(std::max(minTh, 0), 255);
    const int tlow = std::min(iniTh, minTh);
    // score map over relative coords [0,width) x [0,height); detection region [3,width-3) x [3,height-3)
    std::vector<int> score((size_t)L.width * L.height, -1);
    for (int y = 3; y < L.height - 3; y++)
        for (int x = 3; x < L.width - 3; x++) {
            const uint8_t* p = img + (size_t)(y + 16) * w + (x + 16);
            const int v = p[0];
            if (!fast16_pretest(v, p[rdy[0] * w + rdx[0]], p[rdy[4] * w + rdx[4]], p[rdy[8] * w + rdx[8]], p[rdy[12] * w + rdx[12]], tlow)) continue;
            int ring[16];
            for (int k = 0; k < 16; k++) ring[k] = p[rdy[k] * w + rdx[k]];
            score[(size_t)y * L.width + x] = fast16_score(v, ring);
        }
    auto cell_of = [&](int x, int y, int& ci, int& cj) {
        cj = std::min((x - 3) / L.wCell, L.nColsEff - 1);
        ci = std::min((y - 3) / L.hCell, L.nRowsEff - 1);
    };
    auto keep = [&](int x, int y, int t) {
        const int s = score[(size_t)y * L.width + x];
        if (s < t) return false;
        int ci, cj;
        cell_of(x, y, ci, cj);
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy) continue;
                const int xx = x + dx, yy = y + dy;
                if (xx < 3 || xx >= L.width - 3 || yy < 3 || yy >= L.height - 3) continue;
                int ni, nj;
                cell_of(xx, yy, ni, nj);
                if (ni != ci || nj != cj) continue;
                const int sn = score[(size_t)yy * L.width + xx];
                if (sn >= t && sn >= s) return false;   // neighbour is a corner at t with score >= ours
            }
        return s > 0;   // a corner whose score is 0 never beats the zero background (strict >)
    };
    std::vector<int> cellCount((size_t)L.nRowsEff * L.nColsEff, 0);
    for (int y = 3; y < L.height - 3; y++)
        for (int x = 3; x < L.width - 3; x++)
            if (keep(x, y, iniTh)) { int ci, cj; cell_of(x, y, ci, cj); cellCount[ci * L.nColsEff + cj]++; }
    int n = 0;
    for (int y = 3; y < L.height - 3; y++)
        for (int x = 3; x < L.width - 3; x++) {
            int ci, cj;
            cell_of(x, y, ci, cj);
            const int t = cellCount[ci * L.nColsEff + cj] > 0 ? iniTh : minTh;
            if (keep(x, y, t)) { if (n < cap) out[n] = cand_pack(x, y, score[(size_t)y * L.width + x]); n++; }
        }
    return n;
}

uint32_t hm_order_key(uint32_t c, int wCell, int hCell, int nColsEff, int nRowsEff) {
    return cand_order_key(cand_x(c), cand_y(c), wCell, hCell, nColsEff, nRowsEff);
}

// Level-synchronous quadtree (GPU formulation).  cands unordered packed; out = selected packed, list order.
int hm_quadtree(const uint32_t* cands, int n, int width, int height, int nIni, float hX, int N,
                int wCell, int hCell, int nColsEff, int nRowsEff, uint32_t* out, int cap) {
    if (n == 0) return 0;
    const int MAXL = std::max(N + 8, 4 * nIni + 8);
    std::vector<QtNode> bufA(MAXL), bufB(MAXL);
    QtNode* cur = bufA.data();
    QtNode* nxt = bufB.data();
    std::vector<int> cc(MAXL * 4), childpos(MAXL * 4), newpos(MAXL), order(MAXL);
    std::vector<int> node_of(n), quad(n);
    // roots
    std::vector<int> rootCnt(nIni, 0);
    for (int p = 0; p < n; p++) { node_of[p] = (int)((float)cand_x(cands[p]) / hX); rootCnt[node_of[p]]++; }
    std::vector<int> rootPos(nIni, -1);
    int m = 0;
    for (int i = 0; i < nIni; i++) {
        if (rootCnt[i] == 0) continue;
        QtNode r;
        r.x0 = (int16_t)(int)(hX * (float)i); r.x1 = (int16_t)(int)(hX * (float)(i + 1)); r.y0 = 0; r.y1 = (int16_t)height;
        r.cnt = rootCnt[i]; r.seq = i;
        rootPos[i] = m; cur[m++] = r;
    }
    for (int p = 0; p < n; p++) node_of[p] = rootPos[node_of[p]];
    bool finish = false, phase2 = false;
    while (!finish) {
        // parallel part 1: child histograms of multi-point nodes
        std::fill(cc.begin(), cc.begin() + m * 4, 0);
        for (int p = 0; p < n; p++) {
            const QtNode& nd = cur[node_of[p]];
            if (nd.cnt > 1) { quad[p] = qt_quadrant(nd, cand_x(cands[p]), cand_y(cands[p])); cc[node_of[p] * 4 + quad[p]]++; }
        }
        // split order
        int nx = 0;
        for (int i = 0; i < m; i++) if (cur[i].cnt > 1) order[nx++] = i;
        if (phase2) {
            // ascending (size, seq), processed from the back  => descending
            std::sort(order.begin(), order.begin() + nx, [&](int a, int b) {
                if (cur[a].cnt != cur[b].cnt) return cur[a].cnt > cur[b].cnt;
                return cur[a].seq > cur[b].seq;
            });
        }
        int nToExpand = 0;
        const int m2 = qt_rebuild(cur, m, cc.data(), order.data(), nx, phase2, N, nxt, childpos.data(), newpos.data(), &nToExpand);
        // parallel part 2: relabel
        for (int p = 0; p < n; p++) {
            const int i = node_of[p];
            node_of[p] = newpos[i] >= 0 ? newpos[i] : childpos[i * 4 + quad[p]];
        }
        std::swap(cur, nxt);
        if (m2 >= N || m2 == m) finish = true;
        else if (!phase2 && m2 + 3 * nToExpand > N) phase2 = true;
        m = m2;
        if (m > MAXL - 4) return -1;
    }
    // best point per node: max score, then earliest in vToDistributeKeys order
    std::vector<uint64_t> best(m, 0);
    for (int p = 0; p < n; p++) {
        const uint32_t c = cands[p];
        const uint32_t key = cand_order_key(cand_x(c), cand_y(c), wCell, hCell, nColsEff, nRowsEff);
        const uint64_t k = ((uint64_t)cand_score(c) << 32) | (uint64_t)(0xffffffffu - key);
        if (k + 1 > best[node_of[p]]) best[node_of[p]] = k + 1;
    }
    for (int i = 0; i < m && i < cap; i++) {
        const uint64_t k = best[i] - 1;
        const uint32_t key = 0xffffffffu - (uint32_t)(k & 0xffffffffu);
        const int ci = key >> 24, cj = (key >> 16) & 255, ly = (key >> 8) & 255, lx = key & 255;
        out[i] = cand_pack(cj * wCell + lx, ci * hCell + ly, (int)(k >> 32));
    }
    return m;
}

void hm_sincos(const float* x, int n, float* c, float* s) {
    for (int i = 0; i < n; i++) { c[i] = glibc_sincosf(x[i], true); s[i] = glibc_sincosf(x[i], false); }
}
// exhaustive sweep over the float bit patterns [lo, hi]: mismatches of the restated sincosf against this box's libm cosf / sinf
// (first mismatching pattern in *first_bad)
long long hm_sincos_sweep(uint32_t lo, uint32_t hi, uint32_t* first_bad) {
    long long bad = 0;
    for (uint64_t u = lo; u <= hi; u++) {
        const uint32_t bits = (uint32_t)u;
        float x;
        std::memcpy(&x, &bits, 4);
        const float c = glibc_sincosf(x, true), s = glibc_sincosf(x, false);
        const float rc = cosf(x), rs = sinf(x);
        if (std::memcmp(&c, &rc, 4) != 0 || std::memcmp(&s, &rs, 4) != 0) { if (!bad && first_bad) *first_bad = bits; bad++; }
    }
    return bad;
}
float hm_atan2(float y, float x) { return fast_atan2_deg(y, x); }

}  // extern "C"
