// orb_core.h -- exact-arithmetic building blocks shared by the sm_100a kernels and by the CPU "host model"
// that tests/ uses to check the GPU formulation of each stage without a GPU.  Everything here is
// `__host__ __device__`, integer or strictly-IEEE float (no FMA contraction: the translation units are compiled
// with -fmad=false / -ffp-contract=off), so host and device produce identical bits.
//
// What each block reproduces (reference file:line; OpenCV itself is not vendored in the reference):
//   cv_round_f        cvRound(float)               -- src/ORBextractor.cc:81,119-120 call sites
//   fast16_score      cv::FAST cornerScore<16>     -- call sites src/ORBextractor.cc:809-815
//   fast_atan2_deg    cv::fastAtan2                -- src/ORBextractor.cc:103
//   glibc_sincosf     glibc 2.39 cosf/sinf         -- src/ORBextractor.cc:113 (std::cos/std::sin float overloads)
//   quadtree sweep    ORBextractor::DistributeOctTree / ExtractorNode::DivideNode -- src/ORBextractor.cc:481-763
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define ORB_HD __host__ __device__ __forceinline__
#define ORB_HD_NOINLINE __host__ __device__
#else
#define ORB_HD inline
#define ORB_HD_NOINLINE inline
#endif

namespace orbcore {

// ------------------------------------------------------------------------------------------ rounding
ORB_HD int cv_round_f(float v) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}

ORB_HD int imin(int a, int b) { return a < b ? a : b; }
ORB_HD int imax(int a, int b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------ FAST-9/16
// Ring offsets (dx,dy) in OpenCV's order.
#define ORB_RING_DX {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}
#define ORB_RING_DY {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3}

// Corner score of cv::FAST (cornerScore<16>) in closed form.  With d[k] = v - ring[k]:
//   A = max over the 16 arcs of 9 contiguous ring pixels of min(d)   (centre brighter than the arc)
//   B = max over arcs of min(-d)                                      (centre darker than the arc)
// the pixel is a corner at threshold t  <=>  max(A,B) > t, and OpenCV's score (the largest threshold that still
// detects it) is max(A,B) - 1 for any pixel that is a corner at the threshold FAST was called with.
// Returns max(A,B) - 1 (can be negative for flat pixels).
ORB_HD int fast16_score(int v, const int* ring) {
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - ring[k];
    // sliding min / max over windows of 9 on the circular array, log-step
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn2[k] = imin(d[k], d[(k + 1) & 15]); mx2[k] = imax(d[k], d[(k + 1) & 15]); }
    int mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn4[k] = imin(mn2[k], mn2[(k + 2) & 15]); mx4[k] = imax(mx2[k], mx2[(k + 2) & 15]); }
    int A = -256, Bn = 256;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int mn9 = imin(imin(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        int mx9 = imax(imax(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        A = imax(A, mn9);
        Bn = imin(Bn, mx9);
    }
    return imax(A, -Bn) - 1;
}

// Necessary condition for a 9-arc at threshold t: every 9-arc contains two ADJACENT compass pixels (0,4,8,12).
ORB_HD bool fast16_pretest(int v, int p0, int p4, int p8, int p12, int t) {
    const int hi = v + t, lo = v - t;
    const int b = (p0 > hi) | ((p4 > hi) << 1) | ((p8 > hi) << 2) | ((p12 > hi) << 3);
    const int k = (p0 < lo) | ((p4 < lo) << 1) | ((p8 < lo) << 2) | ((p12 < lo) << 3);
    // adjacent pairs on the 4-cycle: (0,1) (1,2) (2,3) (3,0)
    const int bb = b & ((b >> 1) | (b << 3));
    const int kk = k & ((k >> 1) | (k << 3));
    return ((bb | kk) & 15) != 0;
}

// ------------------------------------------------------------------------------------------ fastAtan2
ORB_HD float fmul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
ORB_HD float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
ORB_HD float fsub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
ORB_HD float fdiv_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
ORB_HD double dmul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
ORB_HD double dadd_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// cv::fastAtan2 scalar path (degrees, [0,360]); 7th-order odd polynomial, float32, no FMA.
ORB_HD float fast_atan2_deg(float y, float x) {
    const float scale = 57.295779513082323f;  // (float)(180/CV_PI)
    const float p1 = fmul_rn(0.9997878412794807f, scale), p3 = fmul_rn(-0.3258083974640975f, scale);
    const float p5 = fmul_rn(0.1555786518463281f, scale), p7 = fmul_rn(-0.04432655554792128f, scale);
    const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
    const float ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
    float a, c, c2;
    if (ax >= ay) {
        c = fdiv_rn(ay, fadd_rn(ax, eps));
        c2 = fmul_rn(c, c);
        a = fmul_rn(fadd_rn(fmul_rn(fadd_rn(fmul_rn(fadd_rn(fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = fdiv_rn(ax, fadd_rn(ay, eps));
        c2 = fmul_rn(c, c);
        a = fsub_rn(90.f, fmul_rn(fadd_rn(fmul_rn(fadd_rn(fmul_rn(fadd_rn(fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = fsub_rn(180.f, a);
    if (y < 0) a = fsub_rn(360.f, a);
    return a;
}

// ------------------------------------------------------------------------------------------ cosf / sinf
// glibc 2.39 (ARM optimized-routines) sincosf algorithm on the domain this path uses, x in [0, 2*pi]:
// double-precision range reduction by pi/2 and degree-8/7 minimax polynomials, rounded once to float.
// tests/test_exact_arith.py checks it against this box's libm for EVERY float in [0, 6.2832].
ORB_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
#endif
}

ORB_HD float sincosf_poly(double x, double x2, int n, bool negate_cos) {
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
           C4 = 0x1.99343027bf8c3p-16;
    if (negate_cos) { C0 = -C0; C1 = -C1; C2 = -C2; C3 = -C3; C4 = -C4; }
    if ((n & 1) == 0) {
        const double x3 = dmul_rn(x, x2);
        const double s1 = dadd_rn(S2, dmul_rn(x2, S3));
        const double x7 = dmul_rn(x3, x2);
        const double s = dadd_rn(x, dmul_rn(x3, S1));
        return (float)dadd_rn(s, dmul_rn(x7, s1));
    } else {
        const double x4 = dmul_rn(x2, x2);
        const double c2 = dadd_rn(C3, dmul_rn(x2, C4));
        const double c1 = dadd_rn(C0, dmul_rn(x2, C1));
        const double x6 = dmul_rn(x4, x2);
        const double c = dadd_rn(c1, dmul_rn(x4, C2));
        return (float)dadd_rn(c, dmul_rn(x6, c2));
    }
}

// valid for 0 <= y < 120; is_cos selects cosf, else sinf
ORB_HD float glibc_sincosf(float y, bool is_cos) {
    const uint32_t top = (f32_bits(y) >> 20) & 0x7ff;
    double x = (double)y;
    if (top < 0x3f4) {                 // |y| < 0.75 (abstop12 of pi/4)
        if (top < 0x398) return is_cos ? 1.0f : y;   // |y| < 2^-12
        return sincosf_poly(x, dmul_rn(x, x), is_cos ? 1 : 0, false);
    }
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double r = dmul_rn(x, hpi_inv);
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = dadd_rn(x, -dmul_rn((double)n, hpi));
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[] = {1,-1,-1,1}
    const bool neg = (n & 2) != 0;
    return sincosf_poly(dmul_rn(x, s), dmul_rn(x, x), is_cos ? (n ^ 1) : n, neg);
}

// ------------------------------------------------------------------------------------------ Hamming
ORB_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// ------------------------------------------------------------------------------------------ candidates
// A FAST survivor is one 32-bit word: x (12 bits) | y (12 bits) << 12 | score (8 bits) << 24, x/y relative to the
// 16-px border (minBorderX/Y of src/ORBextractor.cc:773-776).
ORB_HD uint32_t cand_pack(int x, int y, int score) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)score << 24); }
ORB_HD int cand_x(uint32_t c) { return (int)(c & 0xfff); }
ORB_HD int cand_y(uint32_t c) { return (int)((c >> 12) & 0xfff); }
ORB_HD int cand_score(uint32_t c) { return (int)(c >> 24); }

// Position of a candidate in the reference's vToDistributeKeys order (cells row-major, then pixels row-major inside
// the cell, src/ORBextractor.cc:789-829).  Detection regions of the cells tile the level exactly, so the order is a
// pure function of (x, y): key = cell_row | cell_col | y in cell | x in cell.
ORB_HD uint32_t cand_order_key(int x, int y, int wCell, int hCell, int nColsEff, int nRowsEff) {
    int cj = (x - 3) / wCell, ci = (y - 3) / hCell;
    if (cj > nColsEff - 1) cj = nColsEff - 1;
    if (ci > nRowsEff - 1) ci = nRowsEff - 1;
    const int lx = x - cj * wCell, ly = y - ci * hCell;   // < 256 by construction (cells are < 60 px + 6)
    return ((uint32_t)ci << 24) | ((uint32_t)cj << 16) | ((uint32_t)ly << 8) | (uint32_t)lx;
}

// ------------------------------------------------------------------------------------------ quadtree
// Level-synchronous formulation of DistributeOctTree.  At any time all splittable nodes were created by the previous
// sweep, so one sweep = (parallel) count the 4 children of every multi-point node, (sequential, O(#nodes)) rebuild
// the list, (parallel) re-label the points.  The list is an array in list order; push_front of children in creation
// order means a child with creation index c of a sweep that creates T children sits at position T-1-c, and the
// surviving old nodes follow in their old order.
struct QtNode {
    int16_t x0, x1, y0, y1;   // UL.x, UR.x, UL.y, BL.y
    int32_t cnt;              // vKeys.size()
    int32_t seq;              // creation index inside the sweep that created it (address stand-in for the tie-break)
};

ORB_HD int qt_quadrant(const QtNode& n, int x, int y) {
    const int mx = n.x0 + ((n.x1 - n.x0 + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
    const int my = n.y0 + ((n.y1 - n.y0 + 1) >> 1);
    return (x >= mx ? 1 : 0) | (y >= my ? 2 : 0);       // n1,n2,n3,n4 = 0,1,2,3
}

ORB_HD QtNode qt_child(const QtNode& n, int q, int cnt, int seq) {
    const int mx = n.x0 + ((n.x1 - n.x0 + 1) >> 1);
    const int my = n.y0 + ((n.y1 - n.y0 + 1) >> 1);
    QtNode c;
    c.x0 = (int16_t)((q & 1) ? mx : n.x0);
    c.x1 = (int16_t)((q & 1) ? n.x1 : mx);
    c.y0 = (int16_t)((q & 2) ? my : n.y0);
    c.y1 = (int16_t)((q & 2) ? n.y1 : my);
    c.cnt = cnt;
    c.seq = seq;
    return c;
}

// One sweep's sequential part.
//   cur[0..m)      current list (list order)           cc[i*4+q]   points of node i falling in child q (multi nodes)
//   order[0..nx)   positions of the nodes to split, in split order
//   phase2         true: stop splitting as soon as the list holds >= N nodes (src/ORBextractor.cc:730-731)
// Outputs: nxt[0..m') new list, childpos[i*4+q] new position of each created child (-1 otherwise), newpos[i] new
// position of every old node that stays (-1 if split), *nToExpand = created children holding > 1 point.
// Returns m'.
ORB_HD_NOINLINE int qt_rebuild(const QtNode* cur, int m, const int* cc, const int* order, int nx, bool phase2, int N,
                               QtNode* nxt, int* childpos, int* newpos, int* nToExpand) {
    int created = 0, size = m, expandable = 0;
    for (int i = 0; i < m; i++) newpos[i] = 0;
    // pass 1: creation indices
    int nsplit = 0;
    for (int e = 0; e < nx; e++) {
        const int i = order[e];
        for (int q = 0; q < 4; q++) {
            const int c = cc[i * 4 + q];
            if (c > 0) { childpos[i * 4 + q] = created++; size++; if (c > 1) expandable++; }
            else childpos[i * 4 + q] = -1;
        }
        newpos[i] = -1;
        size--;
        nsplit++;
        if (phase2 && size >= N) break;
    }
    // pass 2: positions.  children: created-1-c ; kept old nodes: created + rank
    for (int e = 0; e < nsplit; e++) {
        const int i = order[e];
        for (int q = 0; q < 4; q++) {
            const int c = childpos[i * 4 + q];
            if (c >= 0) {
                const int pos = created - 1 - c;
                childpos[i * 4 + q] = pos;
                nxt[pos] = qt_child(cur[i], q, cc[i * 4 + q], c);
            }
        }
    }
    int rank = created;
    for (int i = 0; i < m; i++) {
        if (newpos[i] < 0) {
            // a node that is in `order` but was not reached because of the early break stays in the list
            continue;
        }
        newpos[i] = rank;
        nxt[rank] = cur[i];
        rank++;
    }
    *nToExpand = expandable;
    return rank;
}

}  // namespace orbcore
