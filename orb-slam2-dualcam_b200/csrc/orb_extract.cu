// orb_extract.cu -- sm_100a implementation of ORBextractor::operator() batched over frames x cameras.
//
// Reference path (file:line under /root/reference): src/ORBextractor.cc:1043-1105 operator(), :1107-1132 ComputePyramid,
// :765-853 ComputeKeyPointsOctTree (+ cv::FAST per 30-px cell), :539-763 DistributeOctTree, :77-104 IC_Angle,
// :108-147 computeOrbDescriptor (+ cv::GaussianBlur 7x7 s=2), bit_pattern_31_ :150-408.
//
// Kernel pipeline per orbx_extract_device() call (all images of the batch in every launch):
//   resize4_kernel        x (nlevels-1)   level l from level l-1, cv::resize INTER_LINEAR 8U fixed-point arithmetic: 4 pixels per thread
//                                         from aligned source words, horizontal taps with dp2a (resize_level_kernel: any scale factor)
//   fast_cells_kernel     x 1             one CTA per strip of up to 7 consecutive 30-px cells of a cell row: the strip's tile of the pyramid level
//                                         is staged in shared memory by TMA (cp.async.bulk.tensor.3d + mbarrier, one tensor map per level);
//                                         compass pre-test and FAST-9/16 score over the strip, NMS and the ini/min threshold fallback per cell,
//                                         unordered packed candidate list per (image, level)
//   quadtree_kernel       x 1             one CTA per (image, level): level-synchronous DistributeOctTree
//   blur_level_kernel     x nlevels       cv::GaussianBlur(7x7, s=2, REFLECT_101) of every level: 64x58 tile per CTA from a 96x64 TMA box
//   describe_kernel       x 1             one warp per keypoint: the blurred 64x37 window by TMA, IC angle from the raw level,
//                                         steered rBRIEF, cv::KeyPoint + 32-byte descriptor output
// There is no CPU fallback: every entry point needs a CUDA device.
#include <cuda.h>              // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "orb_common.h"
#include "orb_core.h"
#include "orb_geometry.h"
#include "rbrief_pattern.h"

using namespace orbcore;

#define ORB_MAX_LEVELS 16
#define ORB_MAX_ROOTS 16

// ------------------------------------------------------------------------------------------------ device-side geometry
struct LevelDev {
    int w, h, pitch;                 // level image
    unsigned long long img_stride;   // bytes between consecutive images of this level
    int width, height;               // maxBorder - minBorder (detection frame, relative coordinates start at (16,16))
    int wCell, hCell, nColsEff, nRowsEff;
    int cand_off, cand_cap;          // slice of the per-image candidate array
    int sel_off, sel_cap;            // slice of the per-image selected array
    int quota, nIni;
    float hX, scale;
    int patch_size, valid;
};

struct CellDesc {                  // a strip of up to FAST_STRIP_PX pixels of consecutive cells of one cell row of ComputeKeyPointsOctTree's grid
    short x0, y0, dw, dh;          // detection region of the whole strip (border-relative), see fast_cells_kernel
    short level, ncell, cw, pad;   // cells of the strip: cell j spans [j cw, (j + 1) cw), the last one runs to dw
    unsigned m_dw, m_cw;           // fastdiv_magic(dw), fastdiv_magic(cw)
};

// One tiled tensor map per pyramid level: (x, y, image) over uint8, box = (tile width rounded up to 16 bytes, tile height, 1).
struct alignas(64) FastMaps {
    CUtensorMap lv[ORB_MAX_LEVELS];
    int bw[ORB_MAX_LEVELS], bh[ORB_MAX_LEVELS];        // box size = tile pitch and rows the TMA writes
};

struct ExtractParams {
    const CellDesc* cells;         // [cells_per_image]
    int nlevels, n_images;
    int iniTh, minTh;
    int cells_per_image;
    int cand_per_image, sel_per_image;
    int cell_begin[ORB_MAX_LEVELS + 1];
    LevelDev lv[ORB_MAX_LEVELS];
    const uint8_t* base[ORB_MAX_LEVELS];   // level images (level 0 = the caller's batch)
    uint8_t* blur[ORB_MAX_LEVELS];         // the same levels after GaussianBlur(7x7, sigma 2, REFLECT_101); pitch / stride of the level, level 0: blur_pitch0 / blur_stride0
    int blur_pitch0;
    unsigned long long blur_stride0;
};

// ------------------------------------------------------------------------------------------------ pyramid
// cv::resize(INTER_LINEAR, 8UC1) as called at src/ORBextractor.cc:1120.  One thread = 4 consecutive destination pixels of one row.
__global__ void __launch_bounds__(128) resize_level_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                           unsigned long long sstride, uint8_t* __restrict__ dst, int dw, int dh,
                                                           int dpitch, unsigned long long dstride, const int* __restrict__ xofs,
                                                           const int* __restrict__ ialpha /* 2 x int16 packed */,
                                                           const int* __restrict__ yofs, const int* __restrict__ ibeta) {
    // items (row, group of 4 pixels) are numbered row-major over the whole level, so no lane idles at the end of a row
    const int nx4 = (dw + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nx4 * dh) return;
    const int dy = item / nx4, dx0 = (item - dy * nx4) * 4;
    const uint8_t* S = src + (unsigned long long)blockIdx.y * sstride;
    uint8_t* D = dst + (unsigned long long)blockIdx.y * dstride + (size_t)dy * dpitch;
    const int sy = yofs[dy];
    const int r0 = min(max(sy, 0), sh - 1), r1 = min(max(sy + 1, 0), sh - 1);
    const int bpk = ibeta[dy];
    const int b0 = (short)(bpk & 0xffff), b1 = (short)(bpk >> 16);
    const uint8_t* R0 = S + (size_t)r0 * spitch;
    const uint8_t* R1 = S + (size_t)r1 * spitch;
    uint32_t outw = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int dx = dx0 + k;
        if (dx < dw) {
            const int sx = xofs[dx];
            const int sx1 = min(sx + 1, sw - 1);
            const int apk = ialpha[dx];
            const int a0 = (short)(apk & 0xffff), a1 = (short)(apk >> 16);
            const int S0 = R0[sx] * a0 + R0[sx1] * a1;
            const int S1 = R1[sx] * a0 + R1[sx1] * a1;
            int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            outw |= (uint32_t)v << (8 * k);
        }
    }
    // rows are padded to a multiple of 16 bytes, so the full word store never leaves the row
    *reinterpret_cast<uint32_t*>(D + dx0) = outw;
}

// The same arithmetic for scale factors up to 2.33, 4 destination pixels per thread with a third of the load instructions: one
// 32-byte table entry per group {aligned source offset, byte offsets of the 4 left taps inside a 12-byte window, the 4 coefficient
// pairs}, three aligned words per source row, the two taps of a pixel funnel-shifted into the low bytes of a register and multiplied
// by the coefficient pair with one dp2a (S = p0 a0 + p1 a1, the integers cv::resize forms).  The byte-gather form above spent
// 69 thread instructions per pixel, 79 % of the LSU wavefront budget.
__global__ void __launch_bounds__(128) resize4_kernel(const uint8_t* __restrict__ src, int sh, int spitch, unsigned long long sstride,
                                                      uint8_t* __restrict__ dst, int dw, int dh, int dpitch, unsigned long long dstride,
                                                      const uint4* __restrict__ xtab, const int* __restrict__ yofs, const int* __restrict__ ibeta) {
    const int nx4 = (dw + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nx4 * dh) return;
    const int dy = item / nx4, g = item - dy * nx4;
    const uint4 t0 = __ldg(xtab + 2 * g), t1 = __ldg(xtab + 2 * g + 1);
    const uint8_t* S = src + (unsigned long long)blockIdx.y * sstride;
    const int sy = __ldg(yofs + dy);
    const int r0 = min(max(sy, 0), sh - 1), r1 = min(max(sy + 1, 0), sh - 1);
    const int bpk = __ldg(ibeta + dy);
    const int b0 = (short)(bpk & 0xffff), b1 = (short)(bpk >> 16);
    // three aligned words per row; the last word of a row is re-read rather than crossed (the taps it would hold carry weight 0)
    const int o0 = (int)t0.x, o1 = min(o0 + 4, spitch - 4), o2 = min(o0 + 8, spitch - 4);
    const uint8_t* R0 = S + (size_t)r0 * spitch;
    const uint8_t* R1 = S + (size_t)r1 * spitch;
    const uint32_t a0 = __ldg(reinterpret_cast<const uint32_t*>(R0 + o0)), a1 = __ldg(reinterpret_cast<const uint32_t*>(R0 + o1)), a2 = __ldg(reinterpret_cast<const uint32_t*>(R0 + o2));
    const uint32_t c0 = __ldg(reinterpret_cast<const uint32_t*>(R1 + o0)), c1 = __ldg(reinterpret_cast<const uint32_t*>(R1 + o1)), c2 = __ldg(reinterpret_cast<const uint32_t*>(R1 + o2));
    const uint32_t apk[4] = {t1.x, t1.y, t1.z, t1.w};
    uint32_t outw = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t o = (t0.y >> (8 * k)) & 0xffu, idx = o >> 2, sh8 = 8 * (o & 3);
        const uint32_t lo0 = idx == 0 ? a0 : (idx == 1 ? a1 : a2), hi0 = idx == 0 ? a1 : a2;
        const uint32_t lo1 = idx == 0 ? c0 : (idx == 1 ? c1 : c2), hi1 = idx == 0 ? c1 : c2;
        const int S0 = (int)__dp2a_lo(apk[k], __funnelshift_r(lo0, hi0, sh8), 0u);
        const int S1 = (int)__dp2a_lo(apk[k], __funnelshift_r(lo1, hi1, sh8), 0u);
        int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        outw |= (uint32_t)v << (8 * k);
    }
    // rows are padded to a multiple of 16 bytes, so the full word store never leaves the row
    *reinterpret_cast<uint32_t*>(dst + (unsigned long long)blockIdx.y * dstride + (size_t)dy * dpitch + 4 * g) = outw;
}

// ------------------------------------------------------------------------------------------------ TMA / mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the initialised barrier must be visible to the async proxy (TMA) as well
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// box of the level's tensor map at (x, y, image) -> shared memory; completion is counted on `bar` in bytes
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ FAST per strip of cells
// ComputeKeyPointsOctTree runs cv::FAST on every 30-px cell on its own (src/ORBextractor.cc:789-829): the detection region of cell
// (ci, cj) is [cj*wCell+3, (cj+1)*wCell+3) x [ci*hCell+3, (ci+1)*hCell+3) in border-relative coordinates (the last effective cell
// runs to width-3 / height-3); regions of different cells tile the level exactly, NMS only looks at neighbours inside the same
// region, and the iniTh -> minTh fallback is decided per cell.  One CTA takes a STRIP of consecutive cells of a cell row (as many as
// fit a 256-byte TMA box): the pre-test and the score do not depend on the cell, so they run over the whole strip at iniTh (full
// thread rows, one barrier set per ~7000 pixels instead of per 900); the NMS clips its 3x3 window to the pixel's own cell; cells
// that kept nothing are then redone at minTh, cell by cell -- exactly what the 815 separate cv::FAST calls of an image do.
#define FAST_THREADS 128
#define FAST_STRIP_PX 216              // + 6 (ring) + 4 (word right of the tile) + 2 x 15 (16-byte alignment of the box) <= 256
#define FAST_MAX_CELLS 16

// floor(i / d) for the small operands of this kernel: m = ceil(2^32 / d) (d >= 2), exact while i * d < 2^32; m = 0 stands for d = 1
__device__ __forceinline__ int fastdiv20(int i, unsigned m) { return m ? (int)__umulhi((unsigned)i, m) : i; }
static inline unsigned fastdiv_magic(int d) { return d <= 1 ? 0u : (unsigned)(((1ull << 32) + (unsigned)d - 1) / (unsigned)d); }

// (the tensor maps live in global memory: a kernel-parameter array indexed by the level would be copied to local memory, where TMA
// cannot read it)
__global__ void __launch_bounds__(FAST_THREADS) fast_cells_kernel(const __grid_constant__ ExtractParams P, const FastMaps* __restrict__ TMp,
                                                                  uint32_t* __restrict__ cand, int* __restrict__ cand_count,
                                                                  int tile_cap, int pix_cap) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    uint8_t* tile = smem;                                  // (dh+6) rows of raw pixels, 4-byte aligned like the source rows
    uint8_t* score = tile + tile_cap;                      // dh x dw
    uint16_t* list = reinterpret_cast<uint16_t*>(score + pix_cap);   // pixels that pass the compass pre-test
    uint16_t* list2 = list + pix_cap;                      // pixels that are corners at their cell's threshold
    __shared__ int s_n1, s_n2, s_base, s_emit, s_cellkeep[FAST_MAX_CELLS];

    const int img = blockIdx.y;
    const CellDesc cd = P.cells[blockIdx.x];
    const int l = cd.level;
    const LevelDev& L = P.lv[l];
    const int x0 = cd.x0, y0 = cd.y0, dw = cd.dw, dh = cd.dh, ncell = cd.ncell, cw = cd.cw;
    if (dw <= 0 || dh <= 0) return;
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;

    // ---- stage the tile with ONE TMA box load: absolute pixel (16 + x0 - 3 + tx, 16 + y0 - 3 + ty) -> tile[ty * tp + dx + tx], tp = box width.
    //      The innermost box coordinate has to be a multiple of 16 bytes (the copy engine traps otherwise), so the box starts at the
    //      16-byte boundary left of the tile; it is the level's largest tile (+ 15) rounded up to 16 bytes, and what it covers beyond
    //      this strip's tile (or beyond the image: zero fill) is never read.
    const int gx0 = 16 + x0 - 3, ax0 = gx0 & ~15, dx = gx0 - ax0;
    const int tp = TMp->bw[l], nw = tp >> 2;
    if (tid == 0) mbar_init(&s_bar, 1);
    if (tid < FAST_MAX_CELLS) s_cellkeep[tid] = 0;
    if (tid == 0) { s_n1 = 0; s_n2 = 0; s_emit = 0; }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, (unsigned)(tp * TMp->bh[l]));
        tma_load_tile(tile, &TMp->lv[l], &s_bar, ax0, 16 + y0 - 3, img);
    }
    const int npix = dw * dh;
    for (int i = tid; i < ((npix + 3) >> 2); i += FAST_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0;
    mbar_wait(&s_bar, 0);
    const unsigned m_dw = cd.m_dw, m_cw = cd.m_cw;
    const uint8_t* T0 = tile + 3 * tp + dx + 3;            // pixel (0, 0) of the detection region
    const int iniTh = min(max(P.iniTh, 0), 255), minTh = min(max(P.minTh, 0), 255);
    const int rdx[16] = ORB_RING_DX, rdy[16] = ORB_RING_DY;

    // ---- pass 1: compass pre-test at threshold t over the columns [rx0, rx0 + rw) of the strip.  A thread takes one aligned tile word
    //      (4 pixels) in 4 consecutive rows: SWAR compares on 16-bit lanes, a 16-bit survivor mask, ONE warp scan / list reservation per
    //      16 pixels (the compaction, not the compares, was two thirds of this pass when it ran per word)
    auto pretest = [&](int rx0, int rw, int t) {
        const int c0 = dx + 3 + rx0;                               // tile column of region x = 0
        const int wc0 = c0 >> 2, nwc = ((c0 + rw - 1) >> 2) - wc0 + 1;
        const unsigned m_nwc = nwc <= 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)nwc - 1) / (unsigned)nwc);
        const uint32_t TH1 = (uint32_t)(t + 1) * 0x00010001u, TL1 = (uint32_t)(512 - t - 1) * 0x00010001u;
        const uint32_t LM = 0x00ff00ffu, B9 = 0x02000200u;
        const int nitems = ((dh + 3) >> 2) * nwc;
        for (int it0 = 0; it0 < nitems; it0 += FAST_THREADS) {
            const int it = it0 + tid;
            uint32_t M16 = 0;
            int ibase = 0;
            if (it < nitems) {
                const int rg = fastdiv20(it, m_nwc), wc = wc0 + (it - rg * nwc), y4 = 4 * rg;
                // (rows past dh read whatever follows in shared memory; their bits are masked)
                const uint32_t* col = reinterpret_cast<const uint32_t*>(tile + y4 * tp) + wc;     // tile row y4 = region row y4 - 3
                uint32_t V[10];
#pragma unroll
                for (int r = 0; r < 10; r++) V[r] = col[r * nw];
                // bytes of this word column that lie inside the region
                const int xlo = wc * 4 - c0;                                 // region x of byte 0
                uint32_t valid = 0xfu;
                if (xlo < 0) valid = (valid << (-xlo)) & 0xfu;
                if (xlo + 3 >= rw) valid &= 0xfu >> (xlo + 4 - rw);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t C = V[r + 3], Cp = col[(r + 3) * nw - 1], Cn = col[(r + 3) * nw + 1];
                    const uint32_t P12 = __byte_perm(Cp, C, 0x4321), P4 = __byte_perm(C, Cn, 0x6543);
                    // with a 512 bias, bit 9 of  p + 512 - (c + t + 1)  is set iff p > c + t, and bit 9 of  (c + 512 - t - 1) - p  iff
                    // p < c - t  (no lane can borrow: all terms stay in [1, 766])
                    const uint32_t Ce = C & LM, Co = (C >> 8) & LM;
                    const uint32_t hie = B9 - (Ce + TH1), hio = B9 - (Co + TH1);      // 512 - (c + t + 1)
                    const uint32_t loe = Ce + TL1, loo = Co + TL1;                    // c + 512 - t - 1
                    uint32_t be[4], bo[4], ke[4], ko[4];
                    const uint32_t W4[4] = {V[r + 6], P4, V[r], P12};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t pe = W4[q] & LM, po = (W4[q] >> 8) & LM;
                        be[q] = pe + hie; bo[q] = po + hio;
                        ke[q] = loe - pe; ko[q] = loo - po;
                    }
                    // two ADJACENT compass pixels brighter, or two darker
                    const uint32_t Me = (((be[0] | be[2]) & (be[1] | be[3])) | ((ke[0] | ke[2]) & (ke[1] | ke[3]))) & B9;
                    const uint32_t Mo = (((bo[0] | bo[2]) & (bo[1] | bo[3])) | ((ko[0] | ko[2]) & (ko[1] | ko[3]))) & B9;
                    const uint32_t M = (Me >> 9) | (Mo >> 1);                           // bit 8k of M <-> byte k of the word
                    uint32_t m4 = ((M * 0x00204081u) >> 21) & valid;                    // bits 0, 8, 16, 24 -> 0..3
                    if (y4 + r >= dh) m4 = 0;
                    M16 |= m4 << (4 * r);
                }
                ibase = y4 * dw + rx0 + xlo;
            }
            const int cnt = __popc(M16);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
            const int wtot = __shfl_sync(0xffffffffu, incl, 31);
            int base = 0;
            if (lane == 31 && wtot) base = atomicAdd(&s_n1, wtot);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
            while (M16) {
                const int b = __ffs(M16) - 1;
                list[base++] = (uint16_t)(ibase + (b >> 2) * dw + (b & 3));
                M16 &= M16 - 1;
            }
        }
    };
    // ---- pass 2: exact corner score of the survivors; pass 3: strict 3x3 maximum inside the pixel's cell (non-corners score 0)
    auto score_and_nms = [&](int t, int n2_begin) {
        const int n1 = s_n1;
        for (int e = tid; e < n1; e += FAST_THREADS) {
            const int i = list[e];
            const int y = fastdiv20(i, m_dw), x = i - y * dw;
            const uint8_t* p = T0 + y * tp + x;
            int ring[16];
#pragma unroll
            for (int k = 0; k < 16; k++) ring[k] = p[rdy[k] * tp + rdx[k]];
            const int s = fast16_score(p[0], ring);
            if (s >= t && s > 0) {
                score[i] = (uint8_t)s;
                list2[atomicAdd(&s_n2, 1)] = (uint16_t)i;
            }
        }
        __syncthreads();
        const int n2 = s_n2;
        for (int e = n2_begin + tid; e < n2; e += FAST_THREADS) {
            const int i = list2[e];
            const int y = fastdiv20(i, m_dw), x = i - y * dw;
            const int c = min(fastdiv20(x, m_cw), ncell - 1);
            const int cx0 = c * cw, cx1 = c == ncell - 1 ? dw : cx0 + cw;
            const int s = score[i];
            bool ismax = true;
#pragma unroll
            for (int ddy = -1; ddy <= 1; ddy++)
#pragma unroll
                for (int ddx = -1; ddx <= 1; ddx++) {
                    if (ddx == 0 && ddy == 0) continue;
                    const int xx = x + ddx, yy = y + ddy;
                    if (xx < cx0 || xx >= cx1 || yy < 0 || yy >= dh) continue;
                    if (score[yy * dw + xx] >= s) ismax = false;
                }
            if (ismax) { atomicAdd(&s_cellkeep[c], 1); list2[e] = (uint16_t)(i | 0x8000); }
        }
        __syncthreads();
    };

    // cv::FAST(cell, iniTh) for every cell of the strip ...
    pretest(0, dw, iniTh);
    __syncthreads();
    score_and_nms(iniTh, 0);
    // ... and, only for the cells where that found nothing, cv::FAST(cell, minTh)   (src/ORBextractor.cc:809-816)
    if (minTh < iniTh) {
        bool any = false;
        for (int c = 0; c < ncell; c++) any |= s_cellkeep[c] == 0;
        if (any) {
            const int n2_begin = s_n2;
            __syncthreads();
            if (tid == 0) s_n1 = 0;
            __syncthreads();
            for (int c = 0; c < ncell; c++)
                if (s_cellkeep[c] == 0) pretest(c * cw, (c == ncell - 1 ? dw : (c + 1) * cw) - c * cw, minTh);
            __syncthreads();
            score_and_nms(minTh, n2_begin);
        }
    }
    int nkeep = 0;
    for (int c = 0; c < ncell; c++) nkeep += s_cellkeep[c];
    if (nkeep == 0) return;

    // ---- emit (unordered; the reference order is a function of (x, y), see cand_order_key)
    if (tid == 0) s_base = atomicAdd(&cand_count[img * P.nlevels + l], nkeep);
    __syncthreads();
    uint32_t* out = cand + (size_t)img * P.cand_per_image + L.cand_off;
    const int n2 = s_n2;
    for (int e = tid; e < n2; e += FAST_THREADS) {
        const int v = list2[e];
        if (!(v & 0x8000)) continue;
        const int i = v & 0x7fff;
        const int y = fastdiv20(i, m_dw), x = i - y * dw;
        const int pos = s_base + atomicAdd(&s_emit, 1);
        if (pos < L.cand_cap) out[pos] = cand_pack(x0 + x, y0 + y, score[i]);
    }
}

// ------------------------------------------------------------------------------------------------ quadtree
// DistributeOctTree (src/ORBextractor.cc:539-763) in its level-synchronous form (orb_core.h): every sweep is
// (parallel) child histograms -> (parallel) split order -> (one thread) list rebuild -> (parallel) relabel.
#define QT_THREADS 256

__global__ void __launch_bounds__(QT_THREADS) quadtree_kernel(const __grid_constant__ ExtractParams P,
                                                              const uint32_t* __restrict__ cand, const int* __restrict__ cand_count,
                                                              uint16_t* __restrict__ node_of_all, uint32_t* __restrict__ sel,
                                                              int* __restrict__ sel_count, int maxl, int* __restrict__ overflow) {
    extern __shared__ __align__(128) uint8_t smem[];
    QtNode* bufA = reinterpret_cast<QtNode*>(smem);
    QtNode* bufB = bufA + maxl;
    int* cc = reinterpret_cast<int*>(bufB + maxl);
    int* childpos = cc + 4 * maxl;
    int* newpos = childpos + 4 * maxl;
    int* order = newpos + maxl;
    unsigned long long* best = reinterpret_cast<unsigned long long*>(order + maxl);   // 72*maxl bytes in: 8-byte aligned
    __shared__ int s_root_cnt[ORB_MAX_ROOTS], s_root_pos[ORB_MAX_ROOTS];
    __shared__ int s_m, s_nx, s_m2, s_nexp;

    const int l = blockIdx.x, img = blockIdx.y;
    const LevelDev& L = P.lv[l];
    const int tid = threadIdx.x;
    int n = L.valid ? cand_count[img * P.nlevels + l] : 0;
    if (n > L.cand_cap) { n = L.cand_cap; if (tid == 0) atomicExch(overflow, 1); }
    if (n == 0) { if (tid == 0) sel_count[img * P.nlevels + l] = 0; return; }
    const uint32_t* C = cand + (size_t)img * P.cand_per_image + L.cand_off;
    uint16_t* node_of = node_of_all + (size_t)img * P.cand_per_image + L.cand_off;
    const int N = L.quota;

    // ---- roots (src/ORBextractor.cc:543-583)
    if (tid < ORB_MAX_ROOTS) s_root_cnt[tid] = 0;
    __syncthreads();
    for (int p = tid; p < n; p += QT_THREADS) {
        const int r = (int)fdiv_rn((float)cand_x(C[p]), L.hX);
        node_of[p] = (uint16_t)r;
        atomicAdd(&s_root_cnt[r], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < L.nIni; i++) {
            s_root_pos[i] = -1;
            if (s_root_cnt[i] == 0) continue;
            QtNode r;
            r.x0 = (int16_t)(int)fmul_rn(L.hX, (float)i);
            r.x1 = (int16_t)(int)fmul_rn(L.hX, (float)(i + 1));
            r.y0 = 0;
            r.y1 = (int16_t)L.height;
            r.cnt = s_root_cnt[i];
            r.seq = i;
            s_root_pos[i] = m;
            bufA[m++] = r;
        }
        s_m = m;
    }
    __syncthreads();
    for (int p = tid; p < n; p += QT_THREADS) node_of[p] = (uint16_t)s_root_pos[node_of[p]];
    QtNode* cur = bufA;
    QtNode* nxt = bufB;
    int m = s_m;
    bool finish = false, phase2 = false;
    __syncthreads();

    while (!finish) {
        for (int i = tid; i < 4 * m; i += QT_THREADS) cc[i] = 0;
        if (tid == 0) s_nx = 0;
        __syncthreads();
        // child histograms of the nodes that will (may) be split
        for (int p = tid; p < n; p += QT_THREADS) {
            const int i = node_of[p];
            const QtNode nd = cur[i];
            if (nd.cnt > 1) {
                const uint32_t c = C[p];
                atomicAdd(&cc[i * 4 + qt_quadrant(nd, cand_x(c), cand_y(c))], 1);
            }
        }
        // split order: list order in the first phase (:598-664); descending (size, creation index) afterwards (:673-738)
        for (int i = tid; i < m; i += QT_THREADS) {
            const QtNode a = cur[i];
            if (a.cnt <= 1) continue;
            int rank = 0;
            if (!phase2) {
                for (int j = 0; j < i; j++) rank += cur[j].cnt > 1;
            } else {
                for (int j = 0; j < m; j++) {
                    const QtNode b = cur[j];
                    rank += (b.cnt > 1) && (b.cnt > a.cnt || (b.cnt == a.cnt && b.seq > a.seq));
                }
            }
            order[rank] = i;
            atomicAdd(&s_nx, 1);
        }
        __syncthreads();
        if (tid == 0) {
            int nexp = 0;
            s_m2 = qt_rebuild(cur, m, cc, order, s_nx, phase2, N, nxt, childpos, newpos, &nexp);
            s_nexp = nexp;
        }
        __syncthreads();
        for (int p = tid; p < n; p += QT_THREADS) {
            const int i = node_of[p];
            int np = newpos[i];
            if (np < 0) {
                const uint32_t c = C[p];
                np = childpos[i * 4 + qt_quadrant(cur[i], cand_x(c), cand_y(c))];
            }
            node_of[p] = (uint16_t)np;
        }
        const int m2 = s_m2;
        if (m2 >= N || m2 == m) finish = true;
        else if (!phase2 && m2 + 3 * s_nexp > N) phase2 = true;
        m = m2;
        QtNode* t = cur; cur = nxt; nxt = t;
        __syncthreads();
        if (m > maxl - 4) { if (tid == 0) atomicExch(overflow, 2); break; }
    }

    // ---- per node: highest response, first in vToDistributeKeys order on ties (src/ORBextractor.cc:744-760)
    for (int i = tid; i < m; i += QT_THREADS) best[i] = 0ull;
    __syncthreads();
    for (int p = tid; p < n; p += QT_THREADS) {
        const uint32_t c = C[p];
        const uint32_t key = cand_order_key(cand_x(c), cand_y(c), L.wCell, L.hCell, L.nColsEff, L.nRowsEff);
        const unsigned long long k = (((unsigned long long)cand_score(c) << 32) | (unsigned long long)(0xffffffffu - key)) + 1ull;
        atomicMax(&best[node_of[p]], k);
    }
    __syncthreads();
    uint32_t* S = sel + (size_t)img * P.sel_per_image + L.sel_off;
    const int mout = min(m, L.sel_cap);
    for (int i = tid; i < mout; i += QT_THREADS) {
        const unsigned long long k = best[i] - 1ull;
        const uint32_t key = 0xffffffffu - (uint32_t)(k & 0xffffffffull);
        const int ci = key >> 24, cj = (key >> 16) & 255, ly = (key >> 8) & 255, lx = key & 255;
        S[i] = cand_pack(cj * L.wCell + lx, ci * L.hCell + ly, (int)(k >> 32));
    }
    if (tid == 0) {
        sel_count[img * P.nlevels + l] = mout;
        if (m > L.sel_cap) atomicExch(overflow, 3);
    }
}

// ------------------------------------------------------------------------------------------------ Gaussian blur of the levels
// cv::GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) as the reference applies it to every level before sampling descriptors
// (src/ORBextractor.cc:1085-1086): integer kernel {18,34,48,56,48,34,18} / 256 per axis, rows then columns, one rounding
// (acc + 2^15) >> 16.  One CTA = a 64 x 58 output tile: the 96 x 64 raw box around it comes in by TMA (zero fill outside the image; the
// REFLECT_101 ring of border tiles is patched in shared memory from the tile itself), the row pass forms four outputs per thread from
// three aligned words with two 16-bit lanes per register (a row sum is at most 255 * 256 < 2^16), the column pass slides down a column.
#define BL_TW 64
#define BL_TH 58
#define BL_BOXW 96             // 16 columns left of the tile (the box must start on a 16-byte boundary) + 64 + 3, rounded up to 16
#define BL_BOXH 64             // 3 + 58 + 3
#define BL_T 256

struct alignas(64) BlurMaps {
    CUtensorMap raw[ORB_MAX_LEVELS];       // box (96, 64, 1) over the raw levels: input of blur_level_kernel
    CUtensorMap blr[ORB_MAX_LEVELS];       // box (64, 37, 1) over the blurred levels: input of describe_kernel
};

__global__ void __launch_bounds__(BL_T) blur_level_kernel(const __grid_constant__ ExtractParams P, const BlurMaps* __restrict__ BM, int l, int tiles_x) {
    __shared__ __align__(128) uint8_t raw[BL_BOXW * BL_BOXH];
    __shared__ __align__(16) uint16_t rowp[BL_BOXH * BL_TW];
    __shared__ __align__(16) uint8_t outt[BL_TH * BL_TW];
    __shared__ __align__(8) uint64_t s_bar;
    const LevelDev& L = P.lv[l];
    const int tid = threadIdx.x, img = blockIdx.y;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * BL_TW, y0 = ty * BL_TH;
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, BL_BOXW * BL_BOXH);
        tma_load_tile(raw, &BM->raw[l], &s_bar, x0 - 16, y0 - 3, img);      // tile pixel (x, y) -> raw[(y - y0 + 3) * 96 + (x - x0 + 16)]
    }
    mbar_wait(&s_bar, 0);
    // ---- BORDER_REFLECT_101 for the tiles on the image border: columns first, then rows (the 2-d reflection is separable)
    const int w = L.w, h = L.h;
    const bool left = x0 == 0, right = x0 + BL_TW + 3 > w, top = y0 == 0, bottom = y0 + BL_TH + 3 > h;
    if (left || right) {
        for (int i = tid; i < BL_BOXH * 3; i += BL_T) {
            const int r = i / 3, k = i - 3 * r + 1;                  // k = 1..3
            uint8_t* row = raw + r * BL_BOXW + 16 - x0;              // row[x] = pixel x of this box row
            if (left) row[-k] = row[k];
            if (right && w - 1 + k < x0 + BL_TW + 3) row[w - 1 + k] = row[w - 1 - k];
        }
        __syncthreads();
    }
    if (top || bottom) {
        for (int i = tid; i < BL_BOXW * 3; i += BL_T) {
            const int c = i / 3, k = i - 3 * c + 1;
            uint8_t* col = raw + (3 - y0) * BL_BOXW + c;             // col[y * 96] = pixel row y of this box column
            if (top) col[-k * BL_BOXW] = col[k * BL_BOXW];
            if (bottom && h - 1 + k < y0 + BL_TH + 3) col[(h - 1 + k) * BL_BOXW] = col[(h - 1 - k) * BL_BOXW];
        }
        __syncthreads();
    }
    // ---- row pass: task = (box row, group of 4 output columns); outputs 4j .. 4j+3 need box columns 13 + 4j .. 22 + 4j = words 3 + j .. 5 + j
    for (int t = tid; t < BL_BOXH * (BL_TW / 4); t += BL_T) {
        const int r = t >> 4, j = t & 15;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(raw + r * BL_BOXW) + 3 + j;
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];          // bytes b0..b11; output k = sum_t g_t b_(1 + k + t)
        // 16-bit lane pairs (b_i, b_(i+2)) for i = 1..7: outputs (0, 2); and (b_(i+1), b_(i+3)): outputs (1, 3)
        const uint32_t p1 = __byte_perm(w0, w0, 0x4341), p2 = __byte_perm(w0, w1, 0x4442), p3 = __byte_perm(w0, w1, 0x4543), p4 = __byte_perm(w1, w1, 0x4240),
                       p5 = __byte_perm(w1, w1, 0x4341), p6 = __byte_perm(w1, w2, 0x4442), p7 = __byte_perm(w1, w2, 0x4543), p8 = __byte_perm(w2, w2, 0x4240);
        // p_i = b_i | b_(i+2) << 16   (selector nibble 4 = zero byte is not available in PRMT: mask instead)
        const uint32_t M = 0x00ff00ffu;
        const uint32_t q1 = p1 & M, q2 = p2 & M, q3 = p3 & M, q4 = p4 & M, q5 = p5 & M, q6 = p6 & M, q7 = p7 & M, q8 = p8 & M;
        const uint32_t o02 = 18u * (q1 + q7) + 34u * (q2 + q6) + 48u * (q3 + q5) + 56u * q4;     // outputs 0 | 2 << 16
        const uint32_t o13 = 18u * (q2 + q8) + 34u * (q3 + q7) + 48u * (q4 + q6) + 56u * q5;     // outputs 1 | 3 << 16
        uint32_t* o = reinterpret_cast<uint32_t*>(rowp + r * BL_TW + 4 * j);
        o[0] = __byte_perm(o02, o13, 0x5410);                       // (out0, out1)
        o[1] = __byte_perm(o02, o13, 0x7632);                       // (out2, out3)
    }
    __syncthreads();
    // ---- column pass: task = (column, quarter of the 58 rows): 15 + 15 + 14 + 14
    {
        const int x = tid & 63, q = tid >> 6;
        const int yb = q < 2 ? 15 * q : 30 + 14 * (q - 2), n = q < 2 ? 15 : 14;
        const uint16_t* r = rowp + yb * BL_TW + x;
        uint32_t a0 = r[0], a1 = r[BL_TW], a2 = r[2 * BL_TW], a3 = r[3 * BL_TW], a4 = r[4 * BL_TW], a5 = r[5 * BL_TW];
#pragma unroll
        for (int y = 0; y < 15; y++) {
            if (y < n) {
                const uint32_t a6 = r[(y + 6) * BL_TW];
                outt[(yb + y) * BL_TW + x] = (uint8_t)((18u * (a0 + a6) + 34u * (a1 + a5) + 48u * (a2 + a4) + 56u * a3 + 32768u) >> 16);
                a0 = a1; a1 = a2; a2 = a3; a3 = a4; a4 = a5; a5 = a6;
            }
        }
    }
    __syncthreads();
    // ---- write the tile: 16 bytes per thread and step, clipped to the level (rows are padded to 16 bytes)
    const int pitch = l == 0 ? P.blur_pitch0 : L.pitch;
    uint8_t* dst = P.blur[l] + (unsigned long long)img * (l == 0 ? P.blur_stride0 : L.img_stride);
    for (int i = tid; i < BL_TH * (BL_TW / 16); i += BL_T) {
        const int y = i >> 2, c = (i & 3) * 16;
        if (y0 + y < h && x0 + c < pitch)
            *reinterpret_cast<uint4*>(dst + (size_t)(y0 + y) * pitch + x0 + c) = *reinterpret_cast<const uint4*>(outt + y * BL_TW + c);
    }
}

// ------------------------------------------------------------------------------------------------ orientation + descriptor
// One warp per keypoint.  IC_Angle (src/ORBextractor.cc:77-104) reads the radius-15 disc of the raw level straight from global memory
// (lanes along a row: one or two sectors per load); the 37 x 37 window of the BLURRED level that the steered pattern (:108-147) can touch
// comes in by TMA (box 64 x 37 starting on the 16-byte boundary left of the window) and the 512 samples are byte reads of shared memory.
#define DESC_WARPS 4
#define BLR_R 18
#define BLR_W 37
#define BLR_BOXW 64

__device__ const int8_t g_pattern[1024] = {ORB_RBRIEF_PATTERN_VALUES};

__global__ void __launch_bounds__(DESC_WARPS * 32) describe_kernel(const __grid_constant__ ExtractParams P, const BlurMaps* __restrict__ BM,
                                                                   const uint32_t* __restrict__ sel, const int* __restrict__ sel_count,
                                                                   orb_keypoint_t* __restrict__ kps, uint8_t* __restrict__ desc,
                                                                   int* __restrict__ counts, int kp_capacity,
                                                                   const int* __restrict__ umax) {
    __shared__ __align__(128) uint8_t s_win[DESC_WARPS][(BLR_BOXW * BLR_W + 127) & ~127];   // every TMA destination on a 128-byte boundary
    __shared__ __align__(8) uint64_t s_bar[DESC_WARPS];
    __shared__ __align__(4) int8_t spat[1024];
    __shared__ int s_umax[16];
    const int img = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) reinterpret_cast<uint32_t*>(spat)[i] = reinterpret_cast<const uint32_t*>(g_pattern)[i];
    if (threadIdx.x < 16) s_umax[threadIdx.x] = umax[threadIdx.x];
    if (lane == 0) mbar_init(&s_bar[warp], 1);
    __syncthreads();

    // which keypoint: global slot -> (level, index) through the per-level counts of this image
    const int slot = blockIdx.x * DESC_WARPS + warp;
    int l = 0, idx = slot, total = 0;
    bool found = false;
    for (int k = 0; k < P.nlevels; k++) {
        const int c = sel_count[img * P.nlevels + k];
        if (!found && idx < c) { l = k; found = true; }
        if (!found) idx -= c;
        total += c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[img] = min(total, kp_capacity);
    if (!found || slot >= kp_capacity) return;

    const LevelDev& L = P.lv[l];
    const uint32_t c = sel[(size_t)img * P.sel_per_image + L.sel_off + idx];
    const int cx = cand_x(c) + 16, cy = cand_y(c) + 16, response = cand_score(c);
    // ---- the blurred window: pixel (cx - 18 + i, cy - 18 + j) -> win[j * 64 + dx + i]
    const int wx0 = cx - BLR_R, ax0 = wx0 & ~15, dx = wx0 - ax0;
    if (lane == 0) {
        mbar_expect_tx(&s_bar[warp], BLR_BOXW * BLR_W);
        tma_load_tile(s_win[warp], &BM->blr[l], &s_bar[warp], ax0, cy - BLR_R, img);
    }
    // ---- IC_Angle on the raw level (key points lie at least 19 px inside the level, so the disc needs no border handling)
    const uint8_t* I = P.base[l] + (unsigned long long)img * L.img_stride;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15, au = u < 0 ? -u : u;
        const uint8_t* col = I + (size_t)(cy - 15) * L.pitch + cx + u;
        // all 31 row loads are issued before the first use (one memory round trip per key point instead of eight): the rows are
        // independent, and what limits this kernel is the latency of these loads, not their number
        int val[31];
#pragma unroll
        for (int v = 0; v < 31; v++) val[v] = au <= s_umax[v < 15 ? 15 - v : v - 15] ? (int)__ldg(col + (size_t)v * L.pitch) : 0;
        int sv = 0, su = 0;
#pragma unroll
        for (int v = 0; v < 31; v++) { su += val[v]; sv += (v - 15) * val[v]; }
        m10 = u * su;
        m01 = sv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered rBRIEF: pair k of word w is handled by lane k%32, the ballot is the little-endian descriptor word
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    const float arad = fmul_rn(angle, factorPI);
    const float a = glibc_sincosf(arad, true), b = glibc_sincosf(arad, false);
    mbar_wait(&s_bar[warp], 0);
    const uint8_t* B = s_win[warp] + BLR_R * BLR_BOXW + dx + BLR_R;      // pixel (cx, cy)
    uint32_t myword = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t pk = reinterpret_cast<const uint32_t*>(spat)[w * 32 + lane];
        const float px0 = (float)(int8_t)(pk & 0xff), py0 = (float)(int8_t)((pk >> 8) & 0xff);
        const float px1 = (float)(int8_t)((pk >> 16) & 0xff), py1 = (float)(int8_t)(pk >> 24);
        const int r0 = cv_round_f(fadd_rn(fmul_rn(px0, b), fmul_rn(py0, a)));
        const int c0 = cv_round_f(fsub_rn(fmul_rn(px0, a), fmul_rn(py0, b)));
        const int r1 = cv_round_f(fadd_rn(fmul_rn(px1, b), fmul_rn(py1, a)));
        const int c1 = cv_round_f(fsub_rn(fmul_rn(px1, a), fmul_rn(py1, b)));
        const int t0 = B[r0 * BLR_BOXW + c0];
        const int t1 = B[r1 * BLR_BOXW + c1];
        const uint32_t word = __ballot_sync(0xffffffffu, t0 < t1);
        if (lane == w) myword = word;
    }
    const size_t o = (size_t)img * kp_capacity + slot;
    if (lane < 8) reinterpret_cast<uint32_t*>(desc + o * 32)[lane] = myword;
    if (lane == 0) {
        orb_keypoint_t kp;
        kp.x = (float)cx;
        kp.y = (float)cy;
        if (l != 0) { kp.x = fmul_rn(kp.x, L.scale); kp.y = fmul_rn(kp.y, L.scale); }
        kp.size = (float)L.patch_size;
        kp.angle = angle;
        kp.response = (float)response;
        kp.octave = l;
        kp.class_id = -1;
        kps[o] = kp;
    }
}

// ================================================================================================ host side
struct orbx {
    int device = 0;
    int W = 0, H = 0, cameras = 0, max_frames = 0, max_images = 0;
    orbgeo::Geometry geo;
    ExtractParams P;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // device buffers
    uint8_t* d_levels = nullptr;       // levels 1.. of all images, level-major
    uint8_t* d_input = nullptr;        // staging for the host API (level 0)
    size_t input_pitch = 0;
    int* d_tables = nullptr;           // resize tables of all levels
    std::vector<size_t> tab_off;       // per level: offset (ints) of xofs, ialpha, yofs, ibeta
    uint4* d_xtab = nullptr;           // resize4_kernel: per level and group of 4 destination columns, two uint4
    std::vector<size_t> xtab_off;      // per level: offset (uint4) into d_xtab; (size_t)-1 = the level needs the general kernel
    FastMaps TM;                       // tensor maps of the FAST tiles (level 0 is re-encoded when the caller's buffer changes)
    FastMaps* d_TM = nullptr;          // ... and their copy in global memory, where the kernel reads them
    BlurMaps BMh;                      // tensor maps of the blur input (raw levels) and of the descriptor windows (blurred levels)
    BlurMaps* d_BM = nullptr;
    uint8_t* d_blur = nullptr;         // blurred levels 0.. of all images, level-major
    void* encode_fn = nullptr;         // cuTensorMapEncodeTiled
    const uint8_t* tm0_ptr = nullptr; size_t tm0_stride = 0; int tm0_images = 0;
    uint32_t* d_cand = nullptr;
    uint16_t* d_node_of = nullptr;
    int* d_cand_count = nullptr;       // [max_images][nlevels] followed by the overflow flag
    uint32_t* d_sel = nullptr;
    int* d_sel_count = nullptr;
    int* d_umax = nullptr;
    CellDesc* d_cells = nullptr;
    orb_keypoint_t* d_kps = nullptr;   // output staging for the host API
    uint8_t* d_desc = nullptr;
    int* d_counts = nullptr;
    int out_capacity = 0;
    // launch configuration
    int fast_tile_cap = 0, fast_pix_cap = 0;
    size_t fast_smem = 0, qt_smem = 0;
    int qt_maxl = 0;
    long long launches = 0;
    // per-stage device timing (bench.py's roofline): ring of event sets, read back by orbx_stage_ms()
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;  // ORBX_PROF_SETS x 5
    int prof_used = 0;
    // last call (debug taps)
    const uint8_t* last_imgs = nullptr;
    size_t last_stride = 0;
    int last_n_images = 0;
};

#define ORBX_PROF_SETS 256

static void orbx_free(orbx* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
    cudaFree(e->d_levels); cudaFree(e->d_input); cudaFree(e->d_tables); cudaFree(e->d_xtab); cudaFree(e->d_TM); cudaFree(e->d_BM); cudaFree(e->d_blur); cudaFree(e->d_cand); cudaFree(e->d_node_of);
    cudaFree(e->d_cand_count); cudaFree(e->d_sel); cudaFree(e->d_sel_count); cudaFree(e->d_umax); cudaFree(e->d_cells);
    cudaFree(e->d_kps); cudaFree(e->d_desc); cudaFree(e->d_counts);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// (x, y, image) uint8 tensor of one pyramid level, box (bw, bh, 1)
static int encode_map(orbx* e, CUtensorMap* map, int bw, int bh, int l, const uint8_t* base, int w, int h, size_t pitch, size_t img_stride, int n_images) {
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_images};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)img_stride};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = ((EncodeTiledFn)e->encode_fn)(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), gdim, gstr, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ORB_FAIL(ORB_E_CUDA, "cuTensorMapEncodeTiled failed for level %d (%d): %dx%d pitch %zu box %dx%d", l, (int)r, w, h, pitch, bw, bh);
    return ORB_OK;
}
static int encode_level_map(orbx* e, int l, const uint8_t* base, int w, int h, size_t pitch, size_t img_stride, int n_images) {
    return encode_map(e, &e->TM.lv[l], e->TM.bw[l], e->TM.bh[l], l, base, w, h, pitch, img_stride, n_images);
}

extern "C" {

int orbx_create(orbx_t** out, int device, int width, int height, int cameras, int max_frames, int nfeatures, float scaleFactor,
                int nlevels, int iniThFAST, int minThFAST) {
    if (!out) ORB_FAIL(ORB_E_INVALID, "orbx_create: out is NULL");
    *out = nullptr;
    if (width < 64 || height < 64 || width > 4000 || height > 4000) ORB_FAIL(ORB_E_INVALID, "orbx_create: image size %dx%d out of range [64,4000]", width, height);
    if (cameras < 1 || max_frames < 1 || (long long)cameras * max_frames > 65535) ORB_FAIL(ORB_E_INVALID, "orbx_create: cameras*max_frames must be in [1,65535]");
    if (nlevels < 1 || nlevels > ORB_MAX_LEVELS || nfeatures < 1 || !(scaleFactor > 1.0f)) ORB_FAIL(ORB_E_INVALID, "orbx_create: bad extractor parameters");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) ORB_FAIL(ORB_E_NO_DEVICE, "orbx_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) ORB_FAIL(ORB_E_NO_DEVICE, "orbx_create: device %d not present (%d devices)", device, ndev);
    cudaDeviceProp prop;
    ORB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) ORB_FAIL(ORB_E_NO_DEVICE, "orbx_create: device %d is sm_%d%d, the kernels are built for sm_100a only", device, prop.major, prop.minor);
    ORB_CUDA(cudaSetDevice(device));

    orbx* e = new (std::nothrow) orbx();
    if (!e) ORB_FAIL(ORB_E_INVALID, "orbx_create: out of host memory");
    e->device = device; e->W = width; e->H = height; e->cameras = cameras; e->max_frames = max_frames;
    e->max_images = cameras * max_frames;
    e->geo = orbgeo::make_geometry(width, height, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    const orbgeo::Geometry& g = e->geo;
    ExtractParams& P = e->P;
    memset(&P, 0, sizeof(P));
    P.nlevels = nlevels; P.iniTh = iniThFAST; P.minTh = minThFAST;

    size_t level_bytes = 0;           // per image, levels >= 1
    std::vector<size_t> level_off(nlevels, 0);
    int cells = 0, cand_total = 0, sel_total = 0, max_dw = 0, max_dh = 0, max_pix = 0, max_quota_l = 0;
    std::vector<CellDesc> cell_table;
    for (int l = 0; l < nlevels; l++) {
        const orbgeo::Level& G = g.lv[l];
        LevelDev& L = P.lv[l];
        if (G.w < 1 || G.h < 1) { orbx_free(e); ORB_FAIL(ORB_E_INVALID, "orbx_create: pyramid level %d is empty", l); }
        L.w = G.w; L.h = G.h; L.pitch = G.pitch;
        L.width = G.width; L.height = G.height; L.wCell = G.wCell; L.hCell = G.hCell;
        L.nColsEff = G.nColsEff; L.nRowsEff = G.nRowsEff;
        L.quota = G.quota; L.nIni = G.nIni; L.hX = G.hX; L.scale = G.scale; L.patch_size = G.patch_size;
        // a level takes part only if it has at least one cell, a usable quadtree root and room for the 43x43 patch logic
        L.valid = G.valid && G.nColsEff > 0 && G.nRowsEff > 0 && G.width > 6 && G.height > 6 && G.quota > 0;
        if (L.valid && G.nIni > ORB_MAX_ROOTS) { orbx_free(e); ORB_FAIL(ORB_E_INVALID, "orbx_create: aspect ratio needs %d quadtree roots (max %d)", G.nIni, ORB_MAX_ROOTS); }
        P.cell_begin[l] = cells;
        L.cand_off = cand_total; L.sel_off = sel_total;
        e->TM.bw[l] = 16; e->TM.bh[l] = 8;
        if (L.valid) {
            // strict 3x3 maxima inside a cell: at most ceil(w/2)*ceil(h/2) per cell
            int cap = 0, lvl_dw = 0, lvl_dh = 0;
            // a cell row is cut into strips of whole cells, as even as the 216-pixel limit of a strip allows
            const int per_max = std::max(1, std::min(FAST_MAX_CELLS, FAST_STRIP_PX / std::max(G.wCell, 1)));
            const int nstrips = (G.nColsEff + per_max - 1) / per_max, per = (G.nColsEff + nstrips - 1) / nstrips;
            for (int ci = 0; ci < G.nRowsEff; ci++) {
                const int y0 = ci * G.hCell + 3, y1 = ci == G.nRowsEff - 1 ? G.height - 3 : y0 + G.hCell;
                const int dh = std::max(y1 - y0, 0);
                for (int cj = 0; cj < G.nColsEff; cj++) {
                    const int x0 = cj * G.wCell + 3, x1 = cj == G.nColsEff - 1 ? G.width - 3 : x0 + G.wCell;
                    cap += ((std::max(x1 - x0, 0) + 1) / 2) * ((dh + 1) / 2);
                }
                for (int cj0 = 0; cj0 < G.nColsEff; cj0 += per) {
                    const int cj1 = std::min(cj0 + per, G.nColsEff);
                    const int x0 = cj0 * G.wCell + 3, x1 = cj1 == G.nColsEff ? G.width - 3 : cj1 * G.wCell + 3;
                    const int dw = std::max(x1 - x0, 0);
                    max_dw = std::max(max_dw, dw); max_dh = std::max(max_dh, dh);
                    max_pix = std::max(max_pix, dw * dh);
                    lvl_dw = std::max(lvl_dw, dw); lvl_dh = std::max(lvl_dh, dh);
                    CellDesc cd;
                    cd.x0 = (short)x0; cd.y0 = (short)y0; cd.dw = (short)dw; cd.dh = (short)dh; cd.level = (short)l;
                    // (a last cell that is empty -- x0 + wCell*k beyond width-3 -- only shortens the strip)
                    cd.ncell = (short)std::max(1, std::min(cj1 - cj0, (dw + G.wCell - 1) / std::max(G.wCell, 1))); cd.cw = (short)G.wCell; cd.pad = 0;
                    cd.m_dw = fastdiv_magic(dw);
                    cd.m_cw = fastdiv_magic(G.wCell);
                    cell_table.push_back(cd);
                    cells++;
                }
            }
            L.cand_cap = cap;
            // TMA box of the level: its largest tile (+ the 3-px ring on every side, + one word so that the SWAR pass may read the word right of the tile)
            e->TM.bw[l] = (15 + lvl_dw + 6 + 4 + 15) & ~15; e->TM.bh[l] = lvl_dh + 6;
            if (e->TM.bw[l] > 256 || e->TM.bh[l] > 256) { orbx_free(e); ORB_FAIL(ORB_E_INVALID, "orbx_create: FAST tile of level %d (%d x %d) exceeds a TMA box", l, e->TM.bw[l], e->TM.bh[l]); }
            L.sel_cap = std::max(G.quota + 2, 4 * G.nIni);
            max_quota_l = std::max(max_quota_l, L.sel_cap);
        }
        cand_total += (L.cand_cap + 3) & ~3;
        sel_total += (L.sel_cap + 3) & ~3;
        if (l > 0) { level_off[l] = level_bytes; level_bytes += (size_t)G.pitch * G.h; }
    }
    P.cell_begin[nlevels] = cells;
    for (int l = nlevels + 1; l <= ORB_MAX_LEVELS; l++) P.cell_begin[l] = cells;
    P.cells_per_image = cells; P.cand_per_image = cand_total; P.sel_per_image = sel_total;
    if (max_pix > 0x7fff) { orbx_free(e); ORB_FAIL(ORB_E_INVALID, "orbx_create: strip of %dx%d pixels is too large", max_dw, max_dh); }

    const size_t NI = (size_t)e->max_images;
    cudaError_t ce = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (ce == cudaSuccess) ce = cudaMalloc(p, bytes ? bytes : 16); };
    alloc((void**)&e->d_levels, level_bytes * NI);
    alloc((void**)&e->d_cand, (size_t)cand_total * NI * sizeof(uint32_t));
    alloc((void**)&e->d_node_of, (size_t)cand_total * NI * sizeof(uint16_t));
    alloc((void**)&e->d_cand_count, (NI * nlevels + 4) * sizeof(int));
    alloc((void**)&e->d_sel, (size_t)sel_total * NI * sizeof(uint32_t));
    alloc((void**)&e->d_sel_count, NI * nlevels * sizeof(int));
    alloc((void**)&e->d_umax, 16 * sizeof(int));
    alloc((void**)&e->d_cells, std::max<size_t>(cell_table.size(), 1) * sizeof(CellDesc));
    // resize tables
    std::vector<int> tabs;
    e->tab_off.assign((size_t)nlevels * 4, 0);
    for (int l = 1; l < nlevels; l++) {
        const orbgeo::Level& G = g.lv[l];
        e->tab_off[l * 4 + 0] = tabs.size();
        for (int x = 0; x < G.w; x++) tabs.push_back(G.xofs[x]);
        e->tab_off[l * 4 + 1] = tabs.size();
        for (int x = 0; x < G.w; x++) tabs.push_back((int)((uint32_t)(uint16_t)G.ialpha[2 * x] | ((uint32_t)(uint16_t)G.ialpha[2 * x + 1] << 16)));
        e->tab_off[l * 4 + 2] = tabs.size();
        for (int y = 0; y < G.h; y++) tabs.push_back(G.yofs[y]);
        e->tab_off[l * 4 + 3] = tabs.size();
        for (int y = 0; y < G.h; y++) tabs.push_back((int)((uint32_t)(uint16_t)G.ibeta[2 * y] | ((uint32_t)(uint16_t)G.ibeta[2 * y + 1] << 16)));
    }
    alloc((void**)&e->d_tables, tabs.size() * sizeof(int));
    // resize4_kernel tables: per group of 4 destination columns {aligned source byte offset, 4 tap offsets inside the 12-byte window, -, -} {4 coefficient pairs}
    std::vector<uint4> xt;
    e->xtab_off.assign((size_t)nlevels, (size_t)-1);
    for (int l = 1; l < nlevels; l++) {
        const orbgeo::Level& G = g.lv[l];
        const int sw = g.lv[l - 1].w;
        std::vector<uint4> lvl;
        bool ok = true;
        for (int x0 = 0; x0 < G.w && ok; x0 += 4) {
            const int base = G.xofs[x0] & ~3;
            uint32_t offs = 0, apk[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; k++) {
                const int x = std::min(x0 + k, G.w - 1);                 // the pad columns of the last group repeat the last pixel
                const int o = G.xofs[x] - base;
                const int a0 = G.ialpha[2 * x], a1 = G.ialpha[2 * x + 1];
                if (o < 0 || o > 10 || a0 < 0 || a1 < 0 || (G.xofs[x] + 1 > sw - 1 && a1 != 0)) ok = false;
                offs |= (uint32_t)(o & 0xff) << (8 * k);
                apk[k] = (uint32_t)(uint16_t)a0 | ((uint32_t)(uint16_t)a1 << 16);
            }
            lvl.push_back(make_uint4((uint32_t)base, offs, 0u, 0u));
            lvl.push_back(make_uint4(apk[0], apk[1], apk[2], apk[3]));
        }
        if (ok && (g.lv[l - 1].pitch & 3) == 0) { e->xtab_off[l] = xt.size(); xt.insert(xt.end(), lvl.begin(), lvl.end()); }
    }
    alloc((void**)&e->d_xtab, std::max<size_t>(xt.size(), 1) * sizeof(uint4));
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess && !tabs.empty()) ce = cudaMemcpy(e->d_tables, tabs.data(), tabs.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess && !xt.empty()) ce = cudaMemcpy(e->d_xtab, xt.data(), xt.size() * sizeof(uint4), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->d_umax, g.umax.data(), 16 * sizeof(int), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess && !cell_table.empty()) ce = cudaMemcpy(e->d_cells, cell_table.data(), cell_table.size() * sizeof(CellDesc), cudaMemcpyHostToDevice);
    P.cells = e->d_cells;
    if (ce != cudaSuccess) {
        int rc = orbhost::check_cuda(ce, "orbx_create allocations", __FILE__, __LINE__);
        orbx_free(e);
        return rc;
    }
    e->stream = e->own_stream;
    for (int l = 1; l < nlevels; l++) {
        P.lv[l].img_stride = (unsigned long long)g.lv[l].pitch * g.lv[l].h;
        P.base[l] = e->d_levels + level_off[l] * NI;
    }
    // tensor maps of the levels the extractor owns (level 0 belongs to the caller: encoded per call)
    {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (ce != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) { orbx_free(e); ORB_FAIL(ORB_E_CUDA, "orbx_create: cuTensorMapEncodeTiled is not available in this driver"); }
        e->encode_fn = fn;
        for (int l = 1; l < nlevels; l++) {
            if (!P.lv[l].valid) continue;
            const int rc = encode_level_map(e, l, P.base[l], g.lv[l].w, g.lv[l].h, (size_t)g.lv[l].pitch, (size_t)P.lv[l].img_stride, e->max_images);
            if (rc != ORB_OK) { orbx_free(e); return rc; }
        }
    }
    // blurred levels (level 0 included) and the tensor maps around them
    {
        size_t blur_bytes = 0;
        std::vector<size_t> boff(nlevels, 0);
        for (int l = 0; l < nlevels; l++) { boff[l] = blur_bytes; blur_bytes += (size_t)g.lv[l].pitch * g.lv[l].h; }
        ce = cudaMalloc((void**)&e->d_blur, std::max<size_t>(blur_bytes * NI, 16));
        if (ce != cudaSuccess) { int rc = orbhost::check_cuda(ce, "orbx_create blurred levels", __FILE__, __LINE__); orbx_free(e); return rc; }
        P.blur_pitch0 = g.lv[0].pitch;
        P.blur_stride0 = (unsigned long long)g.lv[0].pitch * g.lv[0].h;
        memset(&e->BMh, 0, sizeof(e->BMh));
        for (int l = 0; l < nlevels; l++) {
            P.blur[l] = e->d_blur + boff[l] * NI;
            const size_t pitch = (size_t)g.lv[l].pitch, stride = pitch * g.lv[l].h;
            int rc = encode_map(e, &e->BMh.blr[l], BLR_BOXW, BLR_W, l, P.blur[l], g.lv[l].w, g.lv[l].h, pitch, stride, e->max_images);
            if (rc == ORB_OK && l > 0) rc = encode_map(e, &e->BMh.raw[l], BL_BOXW, BL_BOXH, l, P.base[l], g.lv[l].w, g.lv[l].h, pitch, (size_t)P.lv[l].img_stride, e->max_images);
            if (rc != ORB_OK) { orbx_free(e); return rc; }
        }
    }
    ce = cudaMalloc((void**)&e->d_TM, sizeof(FastMaps));
    if (ce == cudaSuccess) ce = cudaMemcpy(e->d_TM, &e->TM, sizeof(FastMaps), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&e->d_BM, sizeof(BlurMaps));
    if (ce == cudaSuccess) ce = cudaMemcpy(e->d_BM, &e->BMh, sizeof(BlurMaps), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { int rc = orbhost::check_cuda(ce, "orbx_create tensor maps", __FILE__, __LINE__); orbx_free(e); return rc; }
    // FAST shared memory: tile (the largest TMA box) + score + 2 lists
    e->fast_pix_cap = (max_pix + 15) & ~15;
    e->fast_tile_cap = 128;
    for (int l = 0; l < nlevels; l++) e->fast_tile_cap = std::max(e->fast_tile_cap, (e->TM.bw[l] * e->TM.bh[l] + 127) & ~127);
    e->fast_smem = (size_t)e->fast_tile_cap + e->fast_pix_cap + 2 * sizeof(uint16_t) * e->fast_pix_cap;
    e->qt_maxl = (max_quota_l + 8 + 1) & ~1;
    e->qt_smem = (size_t)e->qt_maxl * (2 * sizeof(QtNode) + 10 * sizeof(int) + sizeof(unsigned long long)) + 16;
    if (e->fast_smem > 200 * 1024 || e->qt_smem > 200 * 1024) { orbx_free(e); ORB_FAIL(ORB_E_INVALID, "orbx_create: shared memory need too large (fast %zu, quadtree %zu)", e->fast_smem, e->qt_smem); }
    ce = cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->fast_smem);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(quadtree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->qt_smem);
    if (ce != cudaSuccess) {
        int rc = orbhost::check_cuda(ce, "cudaFuncSetAttribute", __FILE__, __LINE__);
        orbx_free(e);
        return rc;
    }
    *out = e;
    return ORB_OK;
}

void orbx_destroy(orbx_t* e) { orbx_free(e); }

int orbx_get_levels(const orbx_t* e) { return e ? e->geo.nlevels : ORB_E_INVALID; }

int orbx_get_tables(const orbx_t* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* features_per_level, int32_t* umax16) {
    if (!e) ORB_FAIL(ORB_E_INVALID, "orbx_get_tables: NULL handle");
    for (int i = 0; i < e->geo.nlevels; i++) {
        if (scale) scale[i] = e->geo.scale[i];
        if (inv_scale) inv_scale[i] = e->geo.inv_scale[i];
        if (sigma2) sigma2[i] = e->geo.sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = e->geo.inv_sigma2[i];
        if (features_per_level) features_per_level[i] = e->geo.lv[i].quota;
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = e->geo.umax[i];
    return ORB_OK;
}

int orbx_max_keypoints(const orbx_t* e) {
    if (!e) return ORB_E_INVALID;
    int n = 0;
    for (int l = 0; l < e->P.nlevels; l++) if (e->P.lv[l].valid) n += e->P.lv[l].sel_cap;
    return n;
}

int orbx_level_size(const orbx_t* e, int level, int* w, int* h) {
    if (!e || level < 0 || level >= e->P.nlevels) ORB_FAIL(ORB_E_INVALID, "orbx_level_size: bad argument");
    if (w) *w = e->P.lv[level].w;
    if (h) *h = e->P.lv[level].h;
    return ORB_OK;
}

int orbx_set_stream(orbx_t* e, void* s) {
    if (!e) ORB_FAIL(ORB_E_INVALID, "orbx_set_stream: NULL handle");
    e->stream = s ? (cudaStream_t)s : e->own_stream;
    return ORB_OK;
}

int orbx_synchronize(orbx_t* e) {
    if (!e) ORB_FAIL(ORB_E_INVALID, "orbx_synchronize: NULL handle");
    ORB_CUDA(cudaSetDevice(e->device));
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    return ORB_OK;
}

long long orbx_launch_count(const orbx_t* e) { return e ? e->launches : 0; }

int orbx_extract_device(orbx_t* e, const uint8_t* d_imgs, int frames, size_t row_stride, orb_keypoint_t* d_kps, uint8_t* d_desc,
                        int32_t* d_counts, int kp_capacity) {
    if (!e) ORB_FAIL(ORB_E_INVALID, "orbx_extract: NULL handle");
    if (frames == 0) return ORB_OK;
    if (!d_imgs || !d_kps || !d_desc || !d_counts) ORB_FAIL(ORB_E_INVALID, "orbx_extract: NULL buffer");
    if (frames < 0 || frames > e->max_frames) ORB_FAIL(ORB_E_INVALID, "orbx_extract: frames=%d exceeds max_frames=%d", frames, e->max_frames);
    if (row_stride < (size_t)e->W) ORB_FAIL(ORB_E_INVALID, "orbx_extract: row_stride %zu < width %d", row_stride, e->W);
    if (((uintptr_t)d_imgs & 15) || (row_stride & 15)) ORB_FAIL(ORB_E_INVALID, "orbx_extract_device: images must be 16-byte aligned with row_stride %% 16 == 0");
    if (kp_capacity < orbx_max_keypoints(e)) ORB_FAIL(ORB_E_INVALID, "orbx_extract: kp_capacity %d < orbx_max_keypoints() = %d", kp_capacity, orbx_max_keypoints(e));
    ORB_CUDA(cudaSetDevice(e->device));
    const int NI = frames * e->cameras;
    ExtractParams& P = e->P;
    P.n_images = NI;
    P.base[0] = d_imgs;
    P.lv[0].pitch = (int)row_stride;
    P.lv[0].img_stride = (unsigned long long)row_stride * e->H;
    e->last_imgs = d_imgs; e->last_stride = row_stride; e->last_n_images = NI;
    cudaStream_t st = e->stream;
    const int L = P.nlevels;

    ORB_CUDA(cudaMemsetAsync(e->d_cand_count, 0, ((size_t)e->max_images * L + 4) * sizeof(int), st));
    cudaEvent_t* pev = (e->profile && e->prof_used < ORBX_PROF_SETS) ? &e->prof_ev[(size_t)e->prof_used * 5] : nullptr;
    if (pev) ORB_CUDA(cudaEventRecord(pev[0], st));
    for (int l = 1; l < L; l++) {
        const LevelDev& S = P.lv[l - 1];
        const LevelDev& D = P.lv[l];
        dim3 grid((((D.w + 3) / 4) * D.h + 127) / 128, NI);
        const int* T = e->d_tables;
        if (e->xtab_off[l] != (size_t)-1 && (S.pitch & 3) == 0 && ((uintptr_t)P.base[l - 1] & 3) == 0)
            resize4_kernel<<<grid, 128, 0, st>>>(P.base[l - 1], S.h, S.pitch, S.img_stride, const_cast<uint8_t*>(P.base[l]), D.w, D.h, D.pitch, D.img_stride,
                                                 e->d_xtab + e->xtab_off[l], T + e->tab_off[l * 4 + 2], T + e->tab_off[l * 4 + 3]);
        else
            resize_level_kernel<<<grid, 128, 0, st>>>(P.base[l - 1], S.w, S.h, S.pitch, S.img_stride, const_cast<uint8_t*>(P.base[l]), D.w, D.h,
                                                      D.pitch, D.img_stride, T + e->tab_off[l * 4 + 0], T + e->tab_off[l * 4 + 1],
                                                      T + e->tab_off[l * 4 + 2], T + e->tab_off[l * 4 + 3]);
        e->launches++;
    }
    if (pev) ORB_CUDA(cudaEventRecord(pev[1], st));
    int* d_overflow = e->d_cand_count + (size_t)e->max_images * L;
    if (P.cells_per_image > 0) {
        if (P.lv[0].valid && (e->tm0_ptr != d_imgs || e->tm0_stride != row_stride || e->tm0_images != NI)) {   // level 0 is the caller's buffer
            const int rc = encode_level_map(e, 0, d_imgs, e->W, e->H, row_stride, row_stride * e->H, NI);
            if (rc != ORB_OK) return rc;
            const int rc2 = encode_map(e, &e->BMh.raw[0], BL_BOXW, BL_BOXH, 0, d_imgs, e->W, e->H, row_stride, row_stride * e->H, NI);
            if (rc2 != ORB_OK) return rc2;
            e->tm0_ptr = d_imgs; e->tm0_stride = row_stride; e->tm0_images = NI;
            ORB_CUDA(cudaMemcpyAsync(&e->d_TM->lv[0], &e->TM.lv[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
            ORB_CUDA(cudaMemcpyAsync(&e->d_BM->raw[0], &e->BMh.raw[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
        }
        fast_cells_kernel<<<dim3(P.cells_per_image, NI), FAST_THREADS, e->fast_smem, st>>>(P, e->d_TM, e->d_cand, e->d_cand_count, e->fast_tile_cap, e->fast_pix_cap);
        e->launches++;
    }
    if (pev) ORB_CUDA(cudaEventRecord(pev[2], st));
    quadtree_kernel<<<dim3(L, NI), QT_THREADS, e->qt_smem, st>>>(P, e->d_cand, e->d_cand_count, e->d_node_of, e->d_sel, e->d_sel_count, e->qt_maxl, d_overflow);
    e->launches++;
    if (pev) ORB_CUDA(cudaEventRecord(pev[3], st));
    // GaussianBlur of every level that holds key points (src/ORBextractor.cc:1085-1086), then orientation + descriptors
    for (int l = 0; l < L; l++) {
        if (!P.lv[l].valid) continue;
        const int tiles_x = (P.lv[l].w + BL_TW - 1) / BL_TW, tiles_y = (P.lv[l].h + BL_TH - 1) / BL_TH;
        blur_level_kernel<<<dim3(tiles_x * tiles_y, NI), BL_T, 0, st>>>(P, e->d_BM, l, tiles_x);
        e->launches++;
    }
    const int maxkp = orbx_max_keypoints(e);
    describe_kernel<<<dim3((std::max(maxkp, 1) + DESC_WARPS - 1) / DESC_WARPS, NI), DESC_WARPS * 32, 0, st>>>(P, e->d_BM, e->d_sel, e->d_sel_count, d_kps, d_desc, d_counts, kp_capacity, e->d_umax);
    e->launches++;
    if (pev) { ORB_CUDA(cudaEventRecord(pev[4], st)); e->prof_used++; }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orbx_profile(orbx_t* e, int enable) {
    if (!e) ORB_FAIL(ORB_E_INVALID, "orbx_profile: NULL handle");
    ORB_CUDA(cudaSetDevice(e->device));
    if (enable && e->prof_ev.empty()) {
        e->prof_ev.resize((size_t)ORBX_PROF_SETS * 5);
        for (cudaEvent_t& ev : e->prof_ev) ORB_CUDA(cudaEventCreate(&ev));
    }
    e->profile = enable != 0;
    e->prof_used = 0;
    return ORB_OK;
}

int orbx_stage_ms(orbx_t* e, double* ms4, int* calls) {
    if (!e || !ms4) ORB_FAIL(ORB_E_INVALID, "orbx_stage_ms: bad argument");
    ORB_CUDA(cudaSetDevice(e->device));
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < 4; k++) ms4[k] = 0.0;
    for (int i = 0; i < e->prof_used; i++)
        for (int k = 0; k < 4; k++) {
            float ms = 0.f;
            ORB_CUDA(cudaEventElapsedTime(&ms, e->prof_ev[(size_t)i * 5 + k], e->prof_ev[(size_t)i * 5 + k + 1]));
            ms4[k] += ms;
        }
    if (calls) *calls = e->prof_used;
    e->prof_used = 0;
    return ORB_OK;
}

int orbx_extract(orbx_t* e, const uint8_t* imgs, int frames, size_t row_stride, orb_keypoint_t* kps, uint8_t* desc, int32_t* counts, int kp_capacity) {
    if (!e) ORB_FAIL(ORB_E_INVALID, "orbx_extract: NULL handle");
    if (frames == 0) return ORB_OK;   // empty input: silent return like src/ORBextractor.cc:1046-1047
    if (!imgs || !kps || !desc || !counts) ORB_FAIL(ORB_E_INVALID, "orbx_extract: NULL buffer");
    if (frames < 0 || frames > e->max_frames) ORB_FAIL(ORB_E_INVALID, "orbx_extract: frames=%d exceeds max_frames=%d", frames, e->max_frames);
    if (row_stride < (size_t)e->W) ORB_FAIL(ORB_E_INVALID, "orbx_extract: row_stride %zu < width %d", row_stride, e->W);
    if (kp_capacity < orbx_max_keypoints(e)) ORB_FAIL(ORB_E_INVALID, "orbx_extract: kp_capacity %d < orbx_max_keypoints() = %d", kp_capacity, orbx_max_keypoints(e));
    ORB_CUDA(cudaSetDevice(e->device));
    const size_t NI = (size_t)frames * e->cameras, NImax = (size_t)e->max_images;
    if (!e->d_input) {
        e->input_pitch = ((size_t)e->W + 15) & ~(size_t)15;
        ORB_CUDA(cudaMalloc((void**)&e->d_input, e->input_pitch * e->H * NImax));
    }
    if (e->out_capacity < kp_capacity) {
        cudaFree(e->d_kps); cudaFree(e->d_desc); cudaFree(e->d_counts);
        e->d_kps = nullptr; e->d_desc = nullptr; e->d_counts = nullptr; e->out_capacity = 0;
        ORB_CUDA(cudaMalloc((void**)&e->d_kps, NImax * kp_capacity * sizeof(orb_keypoint_t)));
        ORB_CUDA(cudaMalloc((void**)&e->d_desc, NImax * kp_capacity * 32));
        ORB_CUDA(cudaMalloc((void**)&e->d_counts, NImax * sizeof(int)));
        e->out_capacity = kp_capacity;
    }
    cudaStream_t st = e->stream;
    ORB_CUDA(cudaMemcpy2DAsync(e->d_input, e->input_pitch, imgs, row_stride, e->W, (size_t)e->H * NI, cudaMemcpyHostToDevice, st));
    int rc = orbx_extract_device(e, e->d_input, frames, e->input_pitch, e->d_kps, e->d_desc, e->d_counts, kp_capacity);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaMemcpyAsync(counts, e->d_counts, NI * sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(kps, e->d_kps, NI * kp_capacity * sizeof(orb_keypoint_t), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(desc, e->d_desc, NI * kp_capacity * 32, cudaMemcpyDeviceToHost, st));
    int overflow = 0;
    ORB_CUDA(cudaMemcpyAsync(&overflow, e->d_cand_count + NImax * e->P.nlevels, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    if (overflow) ORB_FAIL(ORB_E_OVERFLOW, "orbx_extract: internal capacity exceeded (code %d)", overflow);
    return ORB_OK;
}

// ------------------------------------------------------------------------------------------------ stage taps
int orbx_debug_level(orbx_t* e, int img, int level, uint8_t* out, size_t out_bytes) {
    if (!e || !out || level < 0 || level >= e->P.nlevels || img < 0 || img >= e->last_n_images) ORB_FAIL(ORB_E_INVALID, "orbx_debug_level: bad argument");
    const LevelDev& L = e->P.lv[level];
    if (out_bytes < (size_t)L.w * L.h) ORB_FAIL(ORB_E_INVALID, "orbx_debug_level: buffer too small");
    ORB_CUDA(cudaSetDevice(e->device));
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    ORB_CUDA(cudaMemcpy2D(out, L.w, e->P.base[level] + (size_t)img * L.img_stride, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return ORB_OK;
}

static int debug_packed(orbx_t* e, int img, int level, int32_t* xys, int cap, bool selected, int add) {
    if (!e || !xys || level < 0 || level >= e->P.nlevels || img < 0 || img >= e->last_n_images) ORB_FAIL(ORB_E_INVALID, "orbx_debug: bad argument");
    const LevelDev& L = e->P.lv[level];
    ORB_CUDA(cudaSetDevice(e->device));
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    int n = 0;
    ORB_CUDA(cudaMemcpy(&n, (selected ? e->d_sel_count : e->d_cand_count) + (size_t)img * e->P.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> tmp((size_t)std::max(n, 1));
    const uint32_t* src = selected ? e->d_sel + (size_t)img * e->P.sel_per_image + L.sel_off : e->d_cand + (size_t)img * e->P.cand_per_image + L.cand_off;
    if (n > 0) ORB_CUDA(cudaMemcpy(tmp.data(), src, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; i++) {
        xys[3 * i] = cand_x(tmp[i]) + add; xys[3 * i + 1] = cand_y(tmp[i]) + add; xys[3 * i + 2] = cand_score(tmp[i]);
    }
    return n;
}

int orbx_debug_candidates(orbx_t* e, int img, int level, int32_t* xys, int cap) { return debug_packed(e, img, level, xys, cap, false, 0); }
int orbx_debug_selected(orbx_t* e, int img, int level, int32_t* xys, int cap) { return debug_packed(e, img, level, xys, cap, true, 16); }

}  // extern "C"
