// orb_oracle.cpp -- CPU oracle, extractor + brute-force matcher part.
//
// TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  Sequential, single-threaded, dependency-free C++17
// restatement of ORBextractor (/root/reference/src/ORBextractor.cc) including the OpenCV primitives
// it calls (OpenCV itself is NOT in /root/reference: un-vendored system dependency, README pins
// 3.3.1/3.4.0; the only OpenCV runnable here is the cv2 4.13.0 wheel, and the primitives below are
// pinned bit-exactly to it by tests/test_oracle_cv2.py).
//
// Build: g++ -O3 -std=c++17 -ffp-contract=off (no -march=native: reference CMakeLists.txt:11-12).
//
// Determinisations (the reference itself is not deterministic here):
//  * quadtree tie-break between equal-size nodes: the reference compares node ADDRESSES
//    (ORBextractor.cc:684, sort of pair<int,ExtractorNode*>).  Pinned here to creation order:
//    a node created later compares as the higher address.
//  * cos/sin (ORBextractor.cc:113): with <cmath> and `using namespace std` the float overloads are
//    selected, i.e. glibc cosf/sinf.  The oracle calls this box's glibc (2.39) directly.
#include "orb_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <vector>

#include "rbrief_pattern_oracle.h"

namespace {

// ---------------------------------------------------------------- OpenCV scalar helpers
// cvRound = round-half-even under the default FP environment (SSE cvtss2si / lrint).
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_floor(float v) { int i = (int)v; return i - (i > v); }
inline int cv_ceil(float v) { int i = (int)v; return i + (i < v); }
inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- cv::resize INTER_LINEAR, 8UC1
// OpenCV imgproc resize.cpp, fixed-point bilinear: 11-bit coefficients (INTER_RESIZE_COEF_SCALE=2048),
// horizontal pass to int32, vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2.
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = sat_short(cv_round((1.f - fx) * 2048));
        ialpha[2 * dx + 1] = sat_short(cv_round(fx * 2048));
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[2 * dy] = sat_short(cv_round((1.f - fy) * 2048));
        ibeta[2 * dy + 1] = sat_short(cv_round(fy * 2048));
    }
    std::vector<int> row0(dw), row1(dw);
    auto hpass = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            out[dx] = S[sx] * ialpha[2 * dx] + S[sx1] * ialpha[2 * dx + 1];
        }
    };
    auto clip = [&](int y) { return y < 0 ? 0 : (y < sh ? y : sh - 1); };
    int have0 = -1, have1 = -1;
    for (int dy = 0; dy < dh; dy++) {
        int r0 = clip(yofs[dy]), r1 = clip(yofs[dy] + 1);
        if (have0 != r0) {
            if (have1 == r0) { row0.swap(row1); std::swap(have0, have1); }
            else { hpass(r0, row0); have0 = r0; }
        }
        if (have1 != r1) { hpass(r1, row1); have1 = r1; }
        const int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int x = 0; x < dw; x++) {
            int v = (((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2;
            D[x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
}

// ---------------------------------------------------------------- cv::GaussianBlur(7x7, sigma 2) 8UC1
// OpenCV >= 4 fixed-point path: 8.8 kernel {18,34,48,56,48,34,18}/256, BORDER_REFLECT_101, row pass exact
// in 16 bit, column pass in 32 bit, single rounding (acc + 2^15) >> 16.
const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};
inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}
void gaussian7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* T = &tmp[(size_t)y * w];
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3) for (int k = 0; k < 7; k++) acc += kGauss7[k] * S[x + k - 3];
            else for (int k = 0; k < 7; k++) acc += kGauss7[k] * S[reflect101(x + k - 3, w)];
            T[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* R[7];
        for (int k = 0; k < 7; k++) R[k] = &tmp[(size_t)reflect101(y + k - 3, h) * w];
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; k++) acc += (uint32_t)kGauss7[k] * R[k][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------- cv::FAST (TYPE_9_16), features2d fast.cpp
// Bresenham circle of radius 3, order as in OpenCV's makeOffsets.
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// cornerScore<16>: largest threshold for which the pixel is still a corner.
int fast_corner_score(const uint8_t* p, const int* off, int threshold) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 25; k++) d[k] = v - p[off[k]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], d[k + 2]);
        a = std::min(a, d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, d[k + 4]); a = std::min(a, d[k + 5]); a = std::min(a, d[k + 6]);
        a = std::min(a, d[k + 7]); a = std::min(a, d[k + 8]);
        a0 = std::max(a0, std::min(a, d[k]));
        a0 = std::max(a0, std::min(a, d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], d[k + 2]);
        b = std::max(b, d[k + 3]); b = std::max(b, d[k + 4]); b = std::max(b, d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, d[k + 6]); b = std::max(b, d[k + 7]); b = std::max(b, d[k + 8]);
        b0 = std::min(b0, std::max(b, d[k]));
        b0 = std::min(b0, std::max(b, d[k + 9]));
    }
    return -b0 - 1;
}

struct FastKp { int x, y, score; };

void fast9(const uint8_t* img, int w, int h, int stride, int threshold, bool nms, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    int off[25];
    for (int k = 0; k < 25; k++) off[k] = kRingDy[k % 16] * stride + kRingDx[k % 16];
    threshold = std::min(std::max(threshold, 0), 255);
    // three rolling rows of scores + corner positions (same structure as the published algorithm)
    std::vector<uint8_t> sbuf((size_t)3 * w, 0);
    std::vector<std::vector<int>> cpos(3);
    for (int i = 3; i < h - 2; i++) {
        uint8_t* curr = &sbuf[(size_t)((i - 3) % 3) * w];
        std::vector<int>& cp = cpos[(i - 3) % 3];
        std::fill(curr, curr + w, 0);
        cp.clear();
        if (i < h - 3) {
            const uint8_t* row = img + (size_t)i * stride;
            for (int j = 3; j < w - 3; j++) {
                const uint8_t* p = row + j;
                const int v = p[0];
                // 9 contiguous ring pixels all darker than v - t, or all brighter than v + t
                bool corner = false;
                int cd = 0, cb = 0;
                for (int k = 0; k < 25; k++) {
                    int x = p[off[k]];
                    if (x < v - threshold) { if (++cd > 8) { corner = true; break; } } else cd = 0;
                    if (x > v + threshold) { if (++cb > 8) { corner = true; break; } } else cb = 0;
                }
                if (corner) {
                    cp.push_back(j);
                    if (nms) curr[j] = (uint8_t)fast_corner_score(p, off, threshold);
                }
            }
        }
        if (i == 3) continue;
        const uint8_t* prev = &sbuf[(size_t)((i - 4 + 3) % 3) * w];
        const uint8_t* pprev = &sbuf[(size_t)((i - 5 + 3) % 3) * w];
        const std::vector<int>& pc = cpos[(i - 4 + 3) % 3];
        for (int j : pc) {
            int score = prev[j];
            if (!nms || (score > prev[j + 1] && score > prev[j - 1] && score > pprev[j - 1] && score > pprev[j] &&
                         score > pprev[j + 1] && score > curr[j - 1] && score > curr[j] && score > curr[j + 1]))
                out.push_back({j, i - 1, score});
        }
    }
}

// ---------------------------------------------------------------- cv::fastAtan2 (scalar path, degrees)
float fast_atan2(float y, float x) {
    const float scale = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

inline int hamming256(const uint8_t* a, const uint8_t* b) {
    // ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2015-2031): 8 x 32-bit SWAR popcount
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4);
        std::memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

// ---------------------------------------------------------------- quadtree (DistributeOctTree)
struct Cand { float x, y, response; };  // coords relative to (minBorderX, minBorderY), integer valued

struct QNode {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<int> keys;  // indices into the candidate array, in vToDistributeKeys order
    bool noMore = false;
    long seq = 0;           // creation order: stands in for the node's address (see header)
    std::list<QNode>::iterator self;
};

// ExtractorNode::DivideNode, ORBextractor.cc:481-537
void divide_node(const QNode& n, const std::vector<Cand>& c, QNode& n1, QNode& n2, QNode& n3, QNode& n4) {
    const int halfX = (int)std::ceil((float)(n.URx - n.ULx) / 2);
    const int halfY = (int)std::ceil((float)(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (int k : n.keys) {
        const Cand& kp = c[k];
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.keys.push_back(k); else n3.keys.push_back(k);
        } else if (kp.y < n1.BRy) n2.keys.push_back(k);
        else n4.keys.push_back(k);
    }
    n1.noMore = n1.keys.size() == 1; n2.noMore = n2.keys.size() == 1;
    n3.noMore = n3.keys.size() == 1; n4.noMore = n4.keys.size() == 1;
}

// DistributeOctTree, ORBextractor.cc:539-763.  Returns indices of the retained candidates in list order.
std::vector<int> distribute_quadtree(const std::vector<Cand>& cand, int minX, int maxX, int minY, int maxY, int N) {
    std::vector<int> result;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return result;  // reference would index an empty vector; out of contract
    const float hX = (float)(maxX - minX) / nIni;
    std::list<QNode> nodes;
    long seq = 0;
    std::vector<QNode*> ini(nIni);
    for (int i = 0; i < nIni; i++) {
        QNode ni;
        ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
        ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        nodes.push_back(ni);
        ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < cand.size(); i++) ini[(int)(cand[i].x / hX)]->keys.push_back((int)i);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->noMore = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    typedef std::pair<int, long> SizeSeq;  // (size, creation order) == (size, address) of the reference
    std::vector<std::pair<SizeSeq, QNode*>> expandable;
    auto push_children = [&](QNode* kids[4], int& nToExpand) {
        for (int q = 0; q < 4; q++) {
            if (kids[q]->keys.empty()) continue;
            kids[q]->seq = seq++;
            nodes.push_front(*kids[q]);
            if (kids[q]->keys.size() > 1) {
                nToExpand++;
                expandable.push_back({{(int)kids[q]->keys.size(), nodes.front().seq}, &nodes.front()});
                nodes.front().self = nodes.begin();
            }
        }
    };
    bool finish = false;
    while (!finish) {
        const int prevSize = (int)nodes.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->noMore) { ++it; continue; }
            QNode n1, n2, n3, n4;
            divide_node(*it, cand, n1, n2, n3, n4);
            QNode* kids[4] = {&n1, &n2, &n3, &n4};
            push_children(kids, nToExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {
            while (!finish) {
                const int prev2 = (int)nodes.size();
                std::vector<std::pair<SizeSeq, QNode*>> prevExp = expandable;
                expandable.clear();
                std::sort(prevExp.begin(), prevExp.end(),
                          [](const std::pair<SizeSeq, QNode*>& a, const std::pair<SizeSeq, QNode*>& b) { return a.first < b.first; });
                for (int j = (int)prevExp.size() - 1; j >= 0; j--) {
                    QNode n1, n2, n3, n4;
                    divide_node(*prevExp[j].second, cand, n1, n2, n3, n4);
                    QNode* kids[4] = {&n1, &n2, &n3, &n4};
                    int dummy = 0;
                    push_children(kids, dummy);
                    nodes.erase(prevExp[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev2) finish = true;
            }
        }
    }
    result.reserve(nodes.size());
    for (auto& n : nodes) {
        int best = n.keys[0];
        float maxResponse = cand[best].response;
        for (size_t k = 1; k < n.keys.size(); k++)
            if (cand[n.keys[k]].response > maxResponse) { best = n.keys[k]; maxResponse = cand[best].response; }
        result.push_back(best);
    }
    return result;
}

}  // namespace

// ================================================================ extractor object
struct orc_extractor {
    int nfeatures, nlevels, iniTh, minTh;
    float scaleFactor;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> perLevel, umax;
    // per-call state
    std::vector<int> lw, lh;
    std::vector<std::vector<uint8_t>> pyr, blurred;
    std::vector<std::vector<Cand>> cands;
    std::vector<std::vector<orc_keypoint_t>> sel;
    double t[6] = {0, 0, 0, 0, 0, 0};
};

extern "C" {

orc_extractor* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    if (nlevels < 1 || nlevels > 32 || nfeatures < 1) return nullptr;
    orc_extractor* e = new orc_extractor();
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = iniThFAST; e->minTh = minThFAST;
    e->scaleFactor = scaleFactor;
    // ORBextractor.cc:415-446
    e->scale.assign(nlevels, 1.f); e->sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = e->scale[i - 1] * scaleFactor;
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    e->invScale.resize(nlevels); e->invSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) { e->invScale[i] = 1.0f / e->scale[i]; e->invSigma2[i] = 1.0f / e->sigma2[i]; }
    e->perLevel.resize(nlevels);
    float factor = 1.0f / scaleFactor;
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->perLevel[l] = cv_round(nDesired);
        sum += e->perLevel[l];
        nDesired *= factor;
    }
    e->perLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
    // ORBextractor.cc:454-469 (circular patch row half-widths)
    e->umax.assign(16, 0);
    int v, v0, vmax = cv_floor(15 * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(15 * std::sqrt(2.f) / 2);
    const double hp2 = 15 * 15;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = 15, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    e->lw.resize(nlevels); e->lh.resize(nlevels);
    e->pyr.resize(nlevels); e->blurred.resize(nlevels); e->cands.resize(nlevels); e->sel.resize(nlevels);
    return e;
}

void orc_extractor_destroy(orc_extractor* e) { delete e; }
int orc_extractor_nlevels(const orc_extractor* e) { return e->nlevels; }

void orc_extractor_tables(const orc_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int32_t* features_per_level, int32_t* umax16) {
    for (int i = 0; i < e->nlevels; i++) {
        if (scale) scale[i] = e->scale[i];
        if (inv_scale) inv_scale[i] = e->invScale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = e->invSigma2[i];
        if (features_per_level) features_per_level[i] = e->perLevel[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, int stride, orc_keypoint_t* kps, uint8_t* desc, int cap) {
    if (!e || !img || w <= 0 || h <= 0) return -1;
    const int L = e->nlevels;
    // ---- ComputePyramid (ORBextractor.cc:1107-1132).  The 19-px REFLECT_101 border the reference adds is
    // never read by anything below (all keypoints are >= 19 px inside), so levels are stored dense.
    double t0 = now_s();
    for (int l = 0; l < L; l++) {
        float s = e->invScale[l];
        e->lw[l] = cv_round((float)w * s);
        e->lh[l] = cv_round((float)h * s);
        e->pyr[l].resize((size_t)e->lw[l] * e->lh[l]);
        if (l == 0) for (int y = 0; y < h; y++) std::memcpy(&e->pyr[0][(size_t)y * w], img + (size_t)y * stride, w);
        else resize_linear_u8(e->pyr[l - 1].data(), e->lw[l - 1], e->lh[l - 1], e->lw[l - 1], e->pyr[l].data(), e->lw[l], e->lh[l], e->lw[l]);
    }
    double t1 = now_s();
    e->t[0] += t1 - t0;
    // ---- ComputeKeyPointsOctTree (ORBextractor.cc:765-853)
    const float W = 30;
    std::vector<FastKp> cell;
    for (int l = 0; l < L; l++) {
        double ta = now_s();
        const int lw = e->lw[l], lh = e->lh[l];
        const uint8_t* im = e->pyr[l].data();
        const int minBorderX = 19 - 3, minBorderY = minBorderX;
        const int maxBorderX = lw - 19 + 3, maxBorderY = lh - 19 + 3;
        std::vector<Cand>& cands = e->cands[l];
        cands.clear();
        e->sel[l].clear();
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) { e->t[1] += now_s() - ta; continue; }  // level too small for one cell (reference divides by zero)
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                fast9(im + (size_t)y0 * lw + x0, cw, ch, lw, e->iniTh, true, cell);
                if (cell.empty()) fast9(im + (size_t)y0 * lw + x0, cw, ch, lw, e->minTh, true, cell);
                for (const FastKp& k : cell)
                    cands.push_back({(float)k.x + j * wCell, (float)k.y + i * hCell, (float)k.score});
            }
        }
        double tb = now_s();
        e->t[1] += tb - ta;
        std::vector<int> keep = distribute_quadtree(cands, minBorderX, maxBorderX, minBorderY, maxBorderY, e->perLevel[l]);
        const int scaledPatchSize = (int)(31 * e->scale[l]);
        for (int k : keep) {
            orc_keypoint_t kp;
            kp.x = cands[k].x + minBorderX; kp.y = cands[k].y + minBorderY;
            kp.size = (float)scaledPatchSize; kp.angle = -1.f; kp.response = cands[k].response;
            kp.octave = l; kp.class_id = -1;
            e->sel[l].push_back(kp);
        }
        e->t[2] += now_s() - tb;
    }
    // ---- computeOrientation / IC_Angle (ORBextractor.cc:77-104, 472-479)
    double t2 = now_s();
    for (int l = 0; l < L; l++) {
        const int step = e->lw[l];
        for (orc_keypoint_t& kp : e->sel[l]) {
            const uint8_t* center = e->pyr[l].data() + (size_t)cv_round(kp.y) * step + cv_round(kp.x);
            int m_01 = 0, m_10 = 0;
            for (int u = -15; u <= 15; ++u) m_10 += u * center[u];
            for (int v = 1; v <= 15; ++v) {
                int v_sum = 0, d = e->umax[v];
                for (int u = -d; u <= d; ++u) {
                    int val_plus = center[u + v * step], val_minus = center[u - v * step];
                    v_sum += (val_plus - val_minus);
                    m_10 += u * (val_plus + val_minus);
                }
                m_01 += v * v_sum;
            }
            kp.angle = fast_atan2((float)m_01, (float)m_10);
        }
    }
    e->t[3] += now_s() - t2;
    // ---- operator() (ORBextractor.cc:1059-1104): blur, descriptors, scale, concatenate
    int n = 0;
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    for (int l = 0; l < L; l++) {
        e->blurred[l].clear();
        if (e->sel[l].empty()) continue;
        double ta = now_s();
        const int lw = e->lw[l], lh = e->lh[l];
        e->blurred[l].resize((size_t)lw * lh);
        gaussian7_u8(e->pyr[l].data(), lw, lh, lw, e->blurred[l].data(), lw);
        double tb = now_s();
        e->t[4] += tb - ta;
        for (orc_keypoint_t& kp : e->sel[l]) {
            if (n < cap) {
                // computeOrbDescriptor (ORBextractor.cc:108-147)
                float angle = kp.angle * factorPI;
                float a = cosf(angle), b = sinf(angle);
                const uint8_t* center = e->blurred[l].data() + (size_t)cv_round(kp.y) * lw + cv_round(kp.x);
                const int8_t* pat = orb_oracle_pattern;
                uint8_t* d = desc + (size_t)n * 32;
                for (int i = 0; i < 32; i++, pat += 32) {
                    int val = 0;
                    for (int bit = 0; bit < 8; bit++) {
                        const int x0 = pat[4 * bit], y0 = pat[4 * bit + 1], x1 = pat[4 * bit + 2], y1 = pat[4 * bit + 3];
                        int t0 = center[cv_round(x0 * b + y0 * a) * lw + cv_round(x0 * a - y0 * b)];
                        int t1 = center[cv_round(x1 * b + y1 * a) * lw + cv_round(x1 * a - y1 * b)];
                        val |= (t0 < t1) << bit;
                    }
                    d[i] = (uint8_t)val;
                }
                orc_keypoint_t out = kp;
                if (l != 0) { out.x *= e->scale[l]; out.y *= e->scale[l]; }
                kps[n] = out;
            }
            n++;
        }
        e->t[5] += now_s() - tb;
    }
    return n;
}

int orc_level_size(const orc_extractor* e, int level, int* w, int* h) {
    if (level < 0 || level >= e->nlevels) return -1;
    *w = e->lw[level]; *h = e->lh[level];
    return 0;
}
const uint8_t* orc_level_pixels(const orc_extractor* e, int level) { return e->pyr[level].data(); }
const uint8_t* orc_level_blurred(const orc_extractor* e, int level) { return e->blurred[level].empty() ? nullptr : e->blurred[level].data(); }
int orc_level_candidates(const orc_extractor* e, int level, int32_t* xys, int cap) {
    const auto& c = e->cands[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) {
        xys[3 * i] = (int)c[i].x; xys[3 * i + 1] = (int)c[i].y; xys[3 * i + 2] = (int)c[i].response;
    }
    return (int)c.size();
}
int orc_level_selected(const orc_extractor* e, int level, int32_t* xys, int cap) {
    const auto& s = e->sel[level];
    for (size_t i = 0; i < s.size() && (int)i < cap; i++) {
        xys[3 * i] = (int)s[i].x; xys[3 * i + 1] = (int)s[i].y; xys[3 * i + 2] = (int)s[i].response;
    }
    return (int)s.size();
}
void orc_extractor_timers(const orc_extractor* e, double* six) { for (int i = 0; i < 6; i++) six[i] = e->t[i]; }

// ---- primitives
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_gaussian7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    gaussian7_u8(src, w, h, sstride, dst, dstride);
}
int orc_fast9(const uint8_t* img, int w, int h, int stride, int threshold, int nonmax, int32_t* xys, int cap) {
    std::vector<FastKp> out;
    fast9(img, w, h, stride, threshold, nonmax != 0, out);
    for (size_t i = 0; i < out.size() && (int)i < cap; i++) {
        xys[3 * i] = out[i].x; xys[3 * i + 1] = out[i].y; xys[3 * i + 2] = out[i].score;
    }
    return (int)out.size();
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orc_cvround_f(float v) { return cv_round(v); }
int orc_hamming256(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }
void orc_cosf_sinf(const float* x, int n, float* c, float* s) {
    for (int i = 0; i < n; i++) { volatile float v = x[i]; c[i] = cosf(v); s[i] = sinf(v); }
}

// ---- brute-force matcher: per query the nearest and second-nearest train descriptor (first index wins ties).
// Same inner loop shape as the reference's SearchBy* best/second-best scans (e.g. ORBmatcher.cc:575-610).
void orc_match_bruteforce(const uint8_t* dq, int nq, const uint8_t* dt, int nt, int32_t* best_idx, int32_t* best_d, int32_t* second_d) {
    for (int q = 0; q < nq; q++) {
        int bestDist = 256, bestDist2 = 256, bestIdx = -1;
        for (int t = 0; t < nt; t++) {
            const int dist = hamming256(dq + (size_t)q * 32, dt + (size_t)t * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = t; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        best_idx[q] = bestIdx; best_d[q] = bestDist; second_d[q] = bestDist2;
    }
}

}  // extern "C"
