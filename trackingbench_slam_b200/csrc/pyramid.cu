// pyramid.cu — K1: image pyramid by chained cv::resize INTER_LINEAR semantics.
// Replaces Frame::ComputePyramid (reference src/types/Frame.cpp:414-427, cv::resize at :423).
//
// OpenCV's 8-bit INTER_LINEAR is pure integer: 11-bit coefficients (x2048) per output column/row,
// horizontal pass in int32, vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
// The per-column / per-row (offset, coeff) tables involve double->float rounding, so they are
// built once on the host (exactly as OpenCV does) and read by the kernel; the kernel itself is
// integer only and HBM-bound: every source byte is fetched once from DRAM (neighbouring threads
// share rows through L1), every destination byte written once with 4-byte stores.
#include "common.cuh"
#include <cstdlib>
#include "kernels.cuh"
#include <cfloat>
#include <cmath>
#include <algorithm>


static inline int cv_floor_f(float v) { int i = (int)v; return i - (v < (float)i); }

// Host: coefficient table for one axis (OpenCV imgproc/resize.cpp, cv::resize linear branch).
// clamp_coeff = true for x (fx zeroed when the tap is clamped), false for y (rows are clipped,
// coefficients kept) — see oracle/tb_oracle.c:tbo_resize_linear_u8.
void tb_build_resize_tab(int sn, int dn, bool is_x, std::vector<ResizeTab> &tab, std::vector<int32_t> &ofs1)
{
    double inv_scale = (double)dn / sn, scale = 1. / inv_scale;
    tab.resize(dn);
    ofs1.resize(dn);
    for (int d = 0; d < dn; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor_f(f);
        f -= s;
        int s0, s1;
        if (is_x) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= sn - 1) { f = 0; s = sn - 1; }
            s0 = s; s1 = s + 1 < sn ? s + 1 : sn - 1;
        } else {
            s0 = s < 0 ? 0 : (s > sn - 1 ? sn - 1 : s);
            s1 = s + 1 < 0 ? 0 : (s + 1 > sn - 1 ? sn - 1 : s + 1);
        }
        tab[d].ofs = s0;
        ofs1[d] = s1;
        tab[d].c0 = (int16_t)lrintf((1.f - f) * 2048.f);
        tab[d].c1 = (int16_t)lrintf(f * 2048.f);
    }
}

bool tb_resize_is_area2(int sw, int sh, int dw, int dh)
{
    double sx = 1. / ((double)dw / sw), sy = 1. / ((double)dh / sh);
    int ix = (int)lrint(sx), iy = (int)lrint(sy);
    return fabs(sx - ix) < DBL_EPSILON && fabs(sy - iy) < DBL_EPSILON && ix == 2 && iy == 2;
}

// xtab: int4 per output column {ofs0, ofs1, c0, c1}; ytab likewise.
// One CTA produces a PYR_TW x PYR_TH output tile in three phases:
//  (0) the source box it needs (rows ytab[first].x .. ytab[last].y, columns xtab[first].x .. xtab[last].y)
//      is staged in shared memory with 16-byte loads, all issued before any arithmetic;
//  (1) horizontal pass ONCE per source row: H[row][dx] = (S[sx0]*a0 + S[sx1]*a1) >> 4  (fits u16), a thread
//      owns 4 output columns (its 4 table entries stay in registers) and walks the box rows;
//  (2) vertical pass: dst = (((b0*H0) >> 16) + ((b1*H1) >> 16) + 2) >> 2, 4 pixels per 8-byte H read,
//      4-byte stores.
// Every source byte is read from HBM once, every output byte written once; consecutive output rows share
// their source rows through H instead of recomputing the horizontal interpolation.
#define PYR_TW 256
#ifndef PYR_TH
#define PYR_TH 32
#endif
extern __shared__ __align__(128) uint8_t pyr_smem[];

__global__ void __launch_bounds__(256)
pyr_resize_linear_u8(const uint8_t *__restrict__ src, int spitch, size_t sframe,
                     uint8_t *__restrict__ dst, int dpitch, size_t dframe,
                     int dw, int dh, const int4 *__restrict__ xtab, const int4 *__restrict__ ytab, int bp, int bh_max)
{
    const int tid = threadIdx.y * 64 + threadIdx.x;
    const int tx0 = blockIdx.x * PYR_TW, ty0 = blockIdx.y * PYR_TH;
    const int tx1 = min(tx0 + PYR_TW, dw) - 1, ty1 = min(ty0 + PYR_TH, dh) - 1;
    const uint8_t *s = src + (size_t)blockIdx.z * sframe;
    uint8_t *d = dst + (size_t)blockIdx.z * dframe;
    uint8_t *box = pyr_smem;                                                    // bh_max x bp bytes
    uint16_t *Hs = reinterpret_cast<uint16_t *>(pyr_smem + (size_t)bh_max * bp); // bh_max x PYR_TW u16
    // ---- phase 0: source box (tables are monotone)
    const int sx_lo = __ldg(&xtab[tx0]).x & ~15, sx_hi = __ldg(&xtab[tx1]).y;
    const int sy_lo = __ldg(&ytab[ty0]).x, sy_hi = __ldg(&ytab[ty1]).y;
    const int bw16 = (sx_hi - sx_lo + 16) >> 4, bh = sy_hi - sy_lo + 1;
    const uint32_t magic = bw16 > 1 ? 0xFFFFFFFFu / (uint32_t)bw16 + 1u : 0u;
    for (int i = tid; i < bw16 * bh; i += 256) {
        const int r = magic ? (int)__umulhi((uint32_t)i, magic) : i, c = i - r * bw16;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(s + (size_t)(sy_lo + r) * spitch + sx_lo) + c);
        *reinterpret_cast<uint4 *>(box + r * bp + 16 * c) = v;
    }
    // this thread's 4 output columns (phase 1) — fetched while the box is in flight
    const int cx = threadIdx.x * 4;               // column group inside the tile
    int4 xt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { xt[k] = __ldg(&xtab[min(tx0 + cx + k, dw - 1)]); xt[k].x -= sx_lo; xt[k].y -= sx_lo; }
    __syncthreads();
    // ---- phase 1: horizontal pass, rows threadIdx.y, +4, ...
    for (int r = threadIdx.y; r < bh; r += 4) {
        const uint8_t *row = box + r * bp;
        uint32_t h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) h[k] = (uint32_t)((int)row[xt[k].x] * xt[k].z + (int)row[xt[k].y] * xt[k].w) >> 4;
        *reinterpret_cast<uint2 *>(Hs + r * PYR_TW + cx) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
    }
    __syncthreads();
    // ---- phase 2: vertical pass, 4 columns x 8 rows per thread
    const int dx0 = tx0 + cx;
    if (dx0 > tx1) return;
#pragma unroll 2
    for (int kk = 0; kk < PYR_TH / 4; kk++) {
        const int dy = ty0 + threadIdx.y * (PYR_TH / 4) + kk;
        if (dy > ty1) break;
        const int4 yt = __ldg(&ytab[dy]);
        const uint2 a = *reinterpret_cast<const uint2 *>(Hs + (yt.x - sy_lo) * PYR_TW + cx);
        const uint2 b = *reinterpret_cast<const uint2 *>(Hs + (yt.y - sy_lo) * PYR_TW + cx);
        const int h0[4] = {(int)(a.x & 0xFFFF), (int)(a.x >> 16), (int)(a.y & 0xFFFF), (int)(a.y >> 16)};
        const int h1[4] = {(int)(b.x & 0xFFFF), (int)(b.x >> 16), (int)(b.y & 0xFFFF), (int)(b.y >> 16)};
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = (((yt.z * h0[k]) >> 16) + ((yt.w * h1[k]) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xFF) << (8 * k);
        }
        if (dx0 + 3 < dw) {
            *reinterpret_cast<uint32_t *>(d + (size_t)dy * dpitch + dx0) = out;
        } else {
            for (int k = 0; k < 4 && dx0 + k < dw; k++) d[(size_t)dy * dpitch + dx0 + k] = (uint8_t)(out >> (8 * k));
        }
    }
}

// ---- TMA variant: same three phases and arithmetic, but the CTA keeps its output tile and walks PYR_FPC frames. The
// source box of frame f+2 is fetched by the TMA unit (cp.async.bulk.tensor over a 3-D x/y/frame tensor map, two 192-byte
// wide boxes side by side because a box dimension is limited to 256) into a 2-stage mbarrier ring while frames f and f+1
// are interpolated, so no thread issues a global load and the table entries are read once per CTA instead of once per frame.
#define PYR_BOXW 192
#define PYR_FPC 8
#ifndef PYR_ROW_UNROLL
#define PYR_ROW_UNROLL 1   // row loop of the 8-column streaming kernel (2 and 4 measured: see profiles/r02_pyramid_ncu.md)
#endif
constexpr int kPyrRowUnroll = PYR_ROW_UNROLL;
#ifndef PYR_STAGES
#define PYR_STAGES 3   // source boxes in flight (the register file limits the kernel to 3 CTAs/SM, shared memory has room)
#endif
// hi32(a * b) + c
__device__ __forceinline__ uint32_t pyr_madhi(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// lo | hi << 16 as a multiply-add (FMA pipe)
__device__ __forceinline__ uint32_t pyr_pack16(uint32_t lo, uint32_t hi)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pyr_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pyr_tma_box(void *dst, const CUtensorMap *tm, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(pyr_s32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(pyr_s32(bar)) : "memory");
}
__device__ __forceinline__ void pyr_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(pyr_s32(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(256, 3)
pyr_resize_linear_u8_tma(const __grid_constant__ CUtensorMap tmap, uint8_t *__restrict__ dst, int dpitch, size_t dframe,
                         int dw, int dh, const int4 *__restrict__ xtab, const int4 *__restrict__ ytab, int bh_max, int n_frames, int fpc)
{
    __shared__ alignas(8) uint64_t s_full[PYR_STAGES];
    const int tid = threadIdx.y * 64 + threadIdx.x;
    const int tx0 = blockIdx.x * PYR_TW, ty0 = blockIdx.y * PYR_TH;
    const int tx1 = min(tx0 + PYR_TW, dw) - 1, ty1 = min(ty0 + PYR_TH, dh) - 1;
    const int f0 = blockIdx.z * fpc, nf = min(n_frames, f0 + fpc) - f0;
    const int box_bytes = (bh_max * PYR_BOXW + 127) & ~127;                       // one 192-wide box, 128-byte aligned
    uint8_t *boxes = pyr_smem;                                                    // [2 stages][2 boxes][box_bytes]
    uint16_t *Hs = reinterpret_cast<uint16_t *>(pyr_smem + 2 * PYR_STAGES * (size_t)box_bytes); // bh_max x PYR_TW u16
    const int sx_lo = __ldg(&xtab[tx0]).x & ~15;
    const int sy_lo = __ldg(&ytab[ty0]).x, sy_hi = __ldg(&ytab[ty1]).y;
    const int bh = sy_hi - sy_lo + 1;
    if (tid == 0) {
        for (int s = 0; s < PYR_STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pyr_s32(&s_full[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < PYR_STAGES && s < nf; s++) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pyr_s32(&s_full[s])), "r"(2 * bh_max * PYR_BOXW) : "memory");
            pyr_tma_box(boxes + (size_t)(2 * s) * box_bytes, &tmap, sx_lo, sy_lo, f0 + s, &s_full[s]);
            pyr_tma_box(boxes + (size_t)(2 * s + 1) * box_bytes, &tmap, sx_lo + PYR_BOXW, sy_lo, f0 + s, &s_full[s]);
        }
    }
    // this thread's 4 output columns: tap offsets inside a stage (left box, or right box at +box_bytes) and weights
    const int cx = threadIdx.x * 4;
    // The eight taps of a thread's 4 columns span at most 8 source bytes (checked by the launcher): per source row the thread
    // loads the three aligned words that hold them, funnel-shifts them into an 8-byte window that starts at its first tap,
    // and picks each column's tap pair with one PRMT. The 11-bit weights are split into low/high bytes so that
    // a*c0 + b*c1 = DP4A(pair, lo bytes) + 256 * DP4A(pair, hi bytes): shared memory sees 3 word loads instead of 8 byte
    // loads per row (the LSU data path, not the issue rate, bounds this kernel).
    int woff[3];               // stage offsets of the three words
    uint32_t sel01, sel23, cp[4];   // PRMT selectors for the tap pairs of columns (0,1) and (2,3); weights c0 | c1 << 16
    int wsh;
    {
        int4 xt[4];
#pragma unroll
        for (int k = 0; k < 4; k++) xt[k] = __ldg(&xtab[min(tx0 + cx + k, dw - 1)]);
        const int o_first = xt[0].x - sx_lo;
        wsh = 8 * (o_first & 3);
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int x = min(4 * ((o_first >> 2) + j), 2 * PYR_BOXW - 4);
            woff[j] = x < PYR_BOXW ? x : box_bytes + x - PYR_BOXW;
        }
        uint32_t ia[4], ib[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            ia[k] = (uint32_t)(xt[k].x - sx_lo - o_first); ib[k] = (uint32_t)(xt[k].y - sx_lo - o_first);
            cp[k] = ((uint32_t)xt[k].z & 0xFFFF) | ((uint32_t)xt[k].w << 16);
        }
        sel01 = ia[0] | (ib[0] << 4) | (ia[1] << 8) | (ib[1] << 12);
        sel23 = ia[2] | (ib[2] << 4) | (ia[3] << 8) | (ib[3] << 12);
    }
    // this thread's output rows (phase 2): table entries once per CTA
    const int dx0 = tx0 + cx;
    uint32_t yr[PYR_TH / 4], yc[PYR_TH / 4];   // packed: the two H rows (relative to sy_lo) and the two weights
#pragma unroll
    for (int kk = 0; kk < PYR_TH / 4; kk++) {
        const int4 yt = __ldg(&ytab[min(ty0 + threadIdx.y * (PYR_TH / 4) + kk, dh - 1)]);
        yr[kk] = (uint32_t)(yt.x - sy_lo) | ((uint32_t)(yt.y - sy_lo) << 16);
        yc[kk] = (uint32_t)yt.z | ((uint32_t)yt.w << 16);
    }
    __syncthreads();                                       // barriers initialised before anyone waits
    for (int f = 0; f < nf; f++) {
        const int s = f % PYR_STAGES;
        pyr_mbar_wait(&s_full[s], (uint32_t)((f / PYR_STAGES) & 1));
        // ---- phase 1: horizontal pass, rows threadIdx.y, +4, ...
        const uint8_t *stage = boxes + (size_t)(2 * s) * box_bytes;
        for (int r = threadIdx.y; r < bh; r += 4) {
            const uint8_t *row = stage + r * PYR_BOXW;
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(row + woff[0]), w1 = *reinterpret_cast<const uint32_t *>(row + woff[1]),
                           w2 = *reinterpret_cast<const uint32_t *>(row + woff[2]);
            const uint32_t lo = __funnelshift_r(w0, w1, wsh), hi = __funnelshift_r(w1, w2, wsh);
            // a * c0 + b * c1 of one column is ONE dp2a (16-bit weights x 8-bit taps); a PRMT lines up the taps of two columns
            const uint32_t p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);   // (a0, b0, a1, b1), (a2, b2, a3, b3)
            const uint32_t h0 = __dp2a_lo(cp[0], p01, 0u) >> 4, h1 = __dp2a_hi(cp[1], p01, 0u) >> 4;
            const uint32_t h2 = __dp2a_lo(cp[2], p23, 0u) >> 4, h3 = __dp2a_hi(cp[3], p23, 0u) >> 4;
            *reinterpret_cast<uint2 *>(Hs + r * PYR_TW + cx) = make_uint2(pyr_pack16(h0, h1), pyr_pack16(h2, h3));
        }
        __syncthreads();                                   // H complete, stage s drained
        if (tid == 0 && f + PYR_STAGES < nf) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pyr_s32(&s_full[s])), "r"(2 * bh_max * PYR_BOXW) : "memory");
            pyr_tma_box(boxes + (size_t)(2 * s) * box_bytes, &tmap, sx_lo, sy_lo, f0 + f + PYR_STAGES, &s_full[s]);
            pyr_tma_box(boxes + (size_t)(2 * s + 1) * box_bytes, &tmap, sx_lo + PYR_BOXW, sy_lo, f0 + f + PYR_STAGES, &s_full[s]);
        }
        // ---- phase 2: vertical pass, 4 columns x 8 rows per thread
        if (dx0 <= tx1) {
            const int dy0 = ty0 + threadIdx.y * (PYR_TH / 4);
            uint8_t *dp = dst + (size_t)(f0 + f) * dframe + (size_t)dy0 * dpitch + dx0;   // advances by the pitch
            const bool wordst = dx0 + 3 < dw;
#pragma unroll
            for (int kk = 0; kk < PYR_TH / 4; kk++) {
                if (dy0 + kk > ty1) break;
                const uint2 a = *reinterpret_cast<const uint2 *>(Hs + (yr[kk] & 0xFFFF) * PYR_TW + cx);
                const uint2 b = *reinterpret_cast<const uint2 *>(Hs + (yr[kk] >> 16) * PYR_TW + cx);
                // (w * h) >> 16 = high word of (w << 16) * h: one multiply-high per tap on the FMA pipe (the kernel is bound by
                // the integer ALU pipe), the second one adds the first and the rounding constant
                const uint32_t W0 = yc[kk] << 16, W1 = yc[kk] & 0xFFFF0000u;
                const uint32_t h0[4] = {a.x & 0xFFFFu, a.x >> 16, a.y & 0xFFFFu, a.y >> 16};
                const uint32_t h1[4] = {b.x & 0xFFFFu, b.x >> 16, b.y & 0xFFFFu, b.y >> 16};
                uint32_t v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) v[k] = pyr_madhi(W1, h1[k], pyr_madhi(W0, h0[k], 2u)) >> 2;   // <= 255
                const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
                if (wordst) *reinterpret_cast<uint32_t *>(dp) = out;
                else for (int k = 0; k < 4 && dx0 + k < dw; k++) dp[k] = (uint8_t)(out >> (8 * k));
                dp += dpitch;
            }
        }
        __syncthreads();                                   // H free for the next frame
    }
}

// ---- streaming variant (default): no H tile and no block barrier.  A thread owns 4 output columns x 8 output rows and walks
// down its rows: the horizontal interpolation of a source row lives in 4 registers, the vertical taps of consecutive output
// rows are (r, r+1) then (r+1, r+2) or (r+2, r+3), so the lower row of one output is reused as the upper row of the next and
// every source row is interpolated once per 8-row group (1.2 rows per output row + one at the group's start).  Source boxes
// arrive by TMA exactly as above; a stage is handed back through an mbarrier the eight warps arrive on, so the warps of a
// CTA drift apart by up to PYR_STAGES frames instead of meeting at two barriers per frame.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
pyr_resize_linear_u8_stream(const __grid_constant__ CUtensorMap tmap, uint8_t *__restrict__ dst, int dpitch, size_t dframe,
                            int dw, int dh, const int4 *__restrict__ xtab, const int4 *__restrict__ ytab, int bh_max, int n_frames, int fpc)
{
    __shared__ alignas(8) uint64_t s_full[PYR_STAGES], s_empty[PYR_STAGES];
    const int tid = threadIdx.y * 64 + threadIdx.x;
    const int tx0 = blockIdx.x * PYR_TW, ty0 = blockIdx.y * PYR_TH;
    const int tx1 = min(tx0 + PYR_TW, dw) - 1, ty1 = min(ty0 + PYR_TH, dh) - 1;
    const int f0 = blockIdx.z * fpc, nf = min(n_frames, f0 + fpc) - f0;
    const int box_bytes = (bh_max * PYR_BOXW + 127) & ~127;                       // one 192-wide box, 128-byte aligned
    uint8_t *boxes = pyr_smem;                                                    // [PYR_STAGES][2 boxes][box_bytes]
    const int sx_lo = __ldg(&xtab[tx0]).x & ~15;
    const int sy_lo = __ldg(&ytab[ty0]).x;
    auto fetch = [&](int s, int f) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pyr_s32(&s_full[s])), "r"(2 * bh_max * PYR_BOXW) : "memory");
        pyr_tma_box(boxes + (size_t)(2 * s) * box_bytes, &tmap, sx_lo, sy_lo, f, &s_full[s]);
        pyr_tma_box(boxes + (size_t)(2 * s + 1) * box_bytes, &tmap, sx_lo + PYR_BOXW, sy_lo, f, &s_full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < PYR_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pyr_s32(&s_full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pyr_s32(&s_empty[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < PYR_STAGES && s < nf; s++) fetch(s, f0 + s);
    }
    // this thread's 4 output columns: word offsets inside a stage, PRMT selectors of the tap pairs, weights c0 | c1 << 16
    const int cx = threadIdx.x * 4;
    int woff[3];
    uint32_t sel01, sel23, cp[4];
    int wsh;
    {
        int4 xt[4];
#pragma unroll
        for (int k = 0; k < 4; k++) xt[k] = __ldg(&xtab[min(tx0 + cx + k, dw - 1)]);
        const int o_first = xt[0].x - sx_lo;
        wsh = 8 * (o_first & 3);
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int x = min(4 * ((o_first >> 2) + j), 2 * PYR_BOXW - 4);
            woff[j] = x < PYR_BOXW ? x : box_bytes + x - PYR_BOXW;
        }
        uint32_t ia[4], ib[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            ia[k] = (uint32_t)(xt[k].x - sx_lo - o_first); ib[k] = (uint32_t)(xt[k].y - sx_lo - o_first);
            cp[k] = ((uint32_t)xt[k].z & 0xFFFF) | ((uint32_t)xt[k].w << 16);
        }
        sel01 = ia[0] | (ib[0] << 4) | (ia[1] << 8) | (ib[1] << 12);
        sel23 = ia[2] | (ib[2] << 4) | (ia[3] << 8) | (ib[3] << 12);
    }
    // this thread's output rows: source rows (relative to sy_lo, premultiplied by the box pitch) and weights, once per CTA
    const int dx0 = tx0 + cx, dy0 = ty0 + threadIdx.y * (PYR_TH / 4);
    // the tile's output rows: w0 (12 bits) | w1 << 12 | upper source row << 24 | (lower row = upper row + 1) << 31, once per CTA
    __shared__ uint32_t s_yq[PYR_TH];
    if (tid < PYR_TH) {
        const int4 yt = __ldg(&ytab[min(ty0 + tid, dh - 1)]);
        s_yq[tid] = ((uint32_t)yt.z & 0xFFF) | (((uint32_t)yt.w & 0xFFF) << 12) | ((uint32_t)(yt.x - sy_lo) << 24) | (yt.y != yt.x ? 0x80000000u : 0u);
    }
    const uint32_t *yq = s_yq + threadIdx.y * (PYR_TH / 4);
    const bool active = dx0 <= tx1 && dy0 <= ty1;
    __syncthreads();                                       // barriers initialised before anyone waits
    for (int f = 0; f < nf; f++) {
        const int s = f % PYR_STAGES;
        const uint32_t ph = (uint32_t)((f / PYR_STAGES) & 1);
        pyr_mbar_wait(&s_full[s], ph);
        if (active) {
            const uint8_t *stage = boxes + (size_t)(2 * s) * box_bytes;
            // horizontal interpolation of source row r for the 4 columns: 3 word loads, a funnel shift to the first tap, one
            // PRMT per column pair, one dp2a (16-bit weights x 8-bit taps) per column
            auto hrow = [&](int r, uint32_t (&h)[4]) {
                const uint8_t *row = stage + r * PYR_BOXW;
                const uint32_t w0 = *reinterpret_cast<const uint32_t *>(row + woff[0]), w1 = *reinterpret_cast<const uint32_t *>(row + woff[1]),
                               w2 = *reinterpret_cast<const uint32_t *>(row + woff[2]);
                const uint32_t lo = __funnelshift_r(w0, w1, wsh), hi = __funnelshift_r(w1, w2, wsh);
                const uint32_t p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);
                h[0] = __dp2a_lo(cp[0], p01, 0u) >> 4; h[1] = __dp2a_hi(cp[1], p01, 0u) >> 4;
                h[2] = __dp2a_lo(cp[2], p23, 0u) >> 4; h[3] = __dp2a_hi(cp[3], p23, 0u) >> 4;
            };
            uint8_t *dp = dst + (size_t)(f0 + f) * dframe + (size_t)dy0 * dpitch + dx0;   // advances by the pitch
            // two row buffers; an output row needs source rows (r0, r1 = r0 or r0 + 1) and the next one starts at r1 or r1 + 1:
            // the buffer that already holds r0 stays, the other one is refilled, and the WEIGHTS follow the buffers (the
            // vertical blend is symmetric), so no register is copied
            uint32_t hx[4], hy[4];
            int rx = -1, ry = -1;                          // source rows held in hx / hy
            const int nrows = min(PYR_TH / 4, ty1 - dy0 + 1);
#pragma unroll 1
            for (int kk = 0; kk < nrows; kk++) {
                const uint32_t q = yq[kk];
                const int r0 = (int)((q >> 24) & 0x7F), r1 = r0 + (int)(q >> 31);
                const uint32_t W0 = (q & 0xFFFu) << 16, W1 = (q << 4) & 0x0FFF0000u;
                uint32_t wx, wy;
                if (ry == r0) { if (r1 != r0) { hrow(r1, hx); rx = r1; wx = W1; wy = W0; } else { wx = 0u; wy = W0 + W1; } }
                else if (rx == r0) { if (r1 != r0) { hrow(r1, hy); ry = r1; wx = W0; wy = W1; } else { wx = W0 + W1; wy = 0u; } }
                else { hrow(r0, hx); rx = r0; wx = W0; if (r1 != r0) { hrow(r1, hy); ry = r1; wy = W1; } else { wx = W0 + W1; wy = 0u; } }
                // (w * h) >> 16 = high word of (w << 16) * h, one multiply-high per tap (FMA pipe)
                uint32_t v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) v[k] = (__umulhi(wx, hx[k]) + __umulhi(wy, hy[k]) + 2u) >> 2;   // <= 255
                // the row pitch is a multiple of 64 bytes: a whole word fits even where fewer than 4 columns are inside the level
                *reinterpret_cast<uint32_t *>(dp) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
                dp += dpitch;
            }
        }
        // hand the stage back: every warp arrives once; thread 0 refills it when all eight have
        __syncwarp();
        if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pyr_s32(&s_empty[s])) : "memory");
        if (tid == 0 && f + PYR_STAGES < nf) {
            pyr_mbar_wait(&s_empty[s], ph);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fetch(s, f0 + f + PYR_STAGES);
        }
    }
}

// ---- 8 columns per thread (default where the taps allow it): same streaming scheme with blockDim (32, 8) — a thread owns 8
// output columns x 4 output rows.  Per-row bookkeeping (row table entry, buffer choice, store, loop) is paid once per 8 pixels
// instead of once per 4: ncu showed the 4-column kernel issuing on 69 % of the cycles with ~94 instructions per 4-pixel row.
// The taps of columns 0..3 lie in bytes [0, 8) and those of columns 4..7 in bytes [4, 12) of a 12-byte window that starts at
// the thread's first tap (tb_resize_span8_ok checks this on the host).
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
pyr_resize_linear_u8_stream8(const __grid_constant__ CUtensorMap tmap, uint8_t *__restrict__ dst, int dpitch, size_t dframe,
                             int dw, int dh, const int4 *__restrict__ xtab, const int4 *__restrict__ ytab, int bh_max, int n_frames, int fpc)
{
    __shared__ alignas(8) uint64_t s_full[PYR_STAGES], s_empty[PYR_STAGES];
    __shared__ uint32_t s_yq[PYR_TH];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int tx0 = blockIdx.x * PYR_TW, ty0 = blockIdx.y * PYR_TH;
    const int tx1 = min(tx0 + PYR_TW, dw) - 1, ty1 = min(ty0 + PYR_TH, dh) - 1;
    const int f0 = blockIdx.z * fpc, nf = min(n_frames, f0 + fpc) - f0;
    const int box_bytes = (bh_max * PYR_BOXW + 127) & ~127;
    uint8_t *boxes = pyr_smem;                                                    // [PYR_STAGES][2 boxes][box_bytes]
    const int sx_lo = __ldg(&xtab[tx0]).x & ~15;
    const int sy_lo = __ldg(&ytab[ty0]).x;
    auto fetch = [&](int s, int f) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pyr_s32(&s_full[s])), "r"(2 * bh_max * PYR_BOXW) : "memory");
        pyr_tma_box(boxes + (size_t)(2 * s) * box_bytes, &tmap, sx_lo, sy_lo, f, &s_full[s]);
        pyr_tma_box(boxes + (size_t)(2 * s + 1) * box_bytes, &tmap, sx_lo + PYR_BOXW, sy_lo, f, &s_full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < PYR_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pyr_s32(&s_full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pyr_s32(&s_empty[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < PYR_STAGES && s < nf; s++) fetch(s, f0 + s);
    }
    if (tid < PYR_TH) {   // the tile's output rows: w0 (12 bits) | w1 << 12 | upper source row << 24 | (lower row = upper + 1) << 31
        const int4 yt = __ldg(&ytab[min(ty0 + tid, dh - 1)]);
        s_yq[tid] = ((uint32_t)yt.z & 0xFFF) | (((uint32_t)yt.w & 0xFFF) << 12) | ((uint32_t)(yt.x - sy_lo) << 24) | (yt.y != yt.x ? 0x80000000u : 0u);
    }
    // this thread's 8 output columns: four word offsets inside a stage, PRMT selectors of the four tap pairs, weights c0 | c1 << 16
    const int cx = threadIdx.x * 8;
    int woff[4];
    uint32_t sel[4], cp[8];
    int wsh;
    {
        int4 xt[8];
#pragma unroll
        for (int k = 0; k < 8; k++) xt[k] = __ldg(&xtab[min(tx0 + cx + k, dw - 1)]);
        const int o_first = xt[0].x - sx_lo;
        wsh = 8 * (o_first & 3);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int x = min(4 * ((o_first >> 2) + j), 2 * PYR_BOXW - 4);
            woff[j] = x < PYR_BOXW ? x : box_bytes + x - PYR_BOXW;
        }
        uint32_t ia[8], ib[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int base = o_first + (k < 4 ? 0 : 4);       // columns 4..7 select from the window's bytes 4..11
            ia[k] = (uint32_t)(xt[k].x - sx_lo - base) & 7u; ib[k] = (uint32_t)(xt[k].y - sx_lo - base) & 7u;
            cp[k] = ((uint32_t)xt[k].z & 0xFFFF) | ((uint32_t)xt[k].w << 16);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) sel[j] = ia[2 * j] | (ib[2 * j] << 4) | (ia[2 * j + 1] << 8) | (ib[2 * j + 1] << 12);
    }
    const int dx0 = tx0 + cx, dy0 = ty0 + threadIdx.y * (PYR_TH / 8);
    const uint32_t *yq = s_yq + threadIdx.y * (PYR_TH / 8);
    const bool active = dx0 <= tx1 && dy0 <= ty1;
    const bool wide = dx0 + 4 <= tx1;                      // the second word holds at least one column of the level
    __syncthreads();                                       // barriers and the row table initialised before anyone goes on
    for (int f = 0; f < nf; f++) {
        const int s = f % PYR_STAGES;
        const uint32_t ph = (uint32_t)((f / PYR_STAGES) & 1);
        pyr_mbar_wait(&s_full[s], ph);
        if (active) {
            const uint8_t *stage = boxes + (size_t)(2 * s) * box_bytes;
            auto hrow = [&](int r, uint32_t (&h)[8]) {
                const uint8_t *row = stage + r * PYR_BOXW;
                const uint32_t w0 = *reinterpret_cast<const uint32_t *>(row + woff[0]), w1 = *reinterpret_cast<const uint32_t *>(row + woff[1]),
                               w2 = *reinterpret_cast<const uint32_t *>(row + woff[2]), w3 = *reinterpret_cast<const uint32_t *>(row + woff[3]);
                const uint32_t b0 = __funnelshift_r(w0, w1, wsh), b1 = __funnelshift_r(w1, w2, wsh), b2 = __funnelshift_r(w2, w3, wsh);
                const uint32_t p01 = __byte_perm(b0, b1, sel[0]), p23 = __byte_perm(b0, b1, sel[1]);
                const uint32_t p45 = __byte_perm(b1, b2, sel[2]), p67 = __byte_perm(b1, b2, sel[3]);
                h[0] = __dp2a_lo(cp[0], p01, 0u) >> 4; h[1] = __dp2a_hi(cp[1], p01, 0u) >> 4;
                h[2] = __dp2a_lo(cp[2], p23, 0u) >> 4; h[3] = __dp2a_hi(cp[3], p23, 0u) >> 4;
                h[4] = __dp2a_lo(cp[4], p45, 0u) >> 4; h[5] = __dp2a_hi(cp[5], p45, 0u) >> 4;
                h[6] = __dp2a_lo(cp[6], p67, 0u) >> 4; h[7] = __dp2a_hi(cp[7], p67, 0u) >> 4;
            };
            uint8_t *dp = dst + (size_t)(f0 + f) * dframe + (size_t)dy0 * dpitch + dx0;   // advances by the pitch
            uint32_t hx[8], hy[8];
            int rx = -1, ry = -1;                          // source rows held in hx / hy
            const int nrows = min(PYR_TH / 8, ty1 - dy0 + 1);
#pragma unroll kPyrRowUnroll
            for (int kk = 0; kk < nrows; kk++) {
                const uint32_t q = yq[kk];
                const int r0 = (int)((q >> 24) & 0x7F), r1 = r0 + (int)(q >> 31);
                const uint32_t W0 = (q & 0xFFFu) << 16, W1 = (q << 4) & 0x0FFF0000u;
                uint32_t wx, wy;
                if (ry == r0) { if (r1 != r0) { hrow(r1, hx); rx = r1; wx = W1; wy = W0; } else { wx = 0u; wy = W0 + W1; } }
                else if (rx == r0) { if (r1 != r0) { hrow(r1, hy); ry = r1; wx = W0; wy = W1; } else { wx = W0 + W1; wy = 0u; } }
                else { hrow(r0, hx); rx = r0; wx = W0; if (r1 != r0) { hrow(r1, hy); ry = r1; wy = W1; } else { wx = W0 + W1; wy = 0u; } }
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = (__umulhi(wx, hx[k]) + __umulhi(wy, hy[k]) + 2u) >> 2;   // <= 255
                const uint32_t o0 = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
                const uint32_t o1 = __byte_perm(__byte_perm(v[4], v[5], 0x0040), __byte_perm(v[6], v[7], 0x0040), 0x5410);
                // the row pitch is a multiple of 64 bytes: whole words fit even where fewer columns are inside the level
                if (wide) *reinterpret_cast<uint2 *>(dp) = make_uint2(o0, o1);
                else *reinterpret_cast<uint32_t *>(dp) = o0;
                dp += dpitch;
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pyr_s32(&s_empty[s])) : "memory");
        if (tid == 0 && f + PYR_STAGES < nf) {
            pyr_mbar_wait(&s_empty[s], ph);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fetch(s, f0 + f + PYR_STAGES);
        }
    }
}

// Host: shared-memory box (pitch, rows) needed by the worst tile of a level.
void tb_resize_box(const std::vector<ResizeTab> &tx, const std::vector<int32_t> &x1, const std::vector<ResizeTab> &ty,
                   const std::vector<int32_t> &y1, int *bp, int *bh)
{
    int mw = 0, mh = 0;
    const int dw = (int)tx.size(), dh = (int)ty.size();
    for (int a = 0; a < dw; a += PYR_TW) { int b = std::min(a + PYR_TW, dw) - 1; mw = std::max(mw, x1[b] - (tx[a].ofs & ~15) + 16); }
    for (int a = 0; a < dh; a += PYR_TH) { int b = std::min(a + PYR_TH, dh) - 1; mh = std::max(mh, y1[b] - ty[a].ofs + 1); }
    *bp = ((mw + 15) & ~15) + 16;
    *bh = mh;
}

// Host: the TMA kernel needs the 8 taps of every thread's 4 adjacent output columns inside 8 source bytes.
bool tb_resize_span_ok(const std::vector<ResizeTab> &tx, const std::vector<int32_t> &x1)
{
    const int dw = (int)tx.size();
    for (int c = 0; c < dw; c += 4) {
        const int last = std::min(c + 3, dw - 1);
        for (int k = c; k <= last; k++) if (tx[k].ofs < tx[c].ofs || x1[k] - tx[c].ofs > 7) return false;
    }
    return true;
}

// Host: the 8-column kernel needs the taps of columns 0..3 of every group of 8 inside bytes [0, 8) and those of columns 4..7
// inside bytes [4, 12) of the window that starts at the group's first tap.
bool tb_resize_span8_ok(const std::vector<ResizeTab> &tx, const std::vector<int32_t> &x1)
{
    const int dw = (int)tx.size();
    for (int c = 0; c < dw; c += 8) {
        const int last = std::min(c + 7, dw - 1), o = tx[c].ofs;
        for (int k = c; k <= last; k++) {
            const int lo = k - c < 4 ? 0 : 4;
            if (tx[k].ofs - o < lo || x1[k] - o < lo || tx[k].ofs - o > lo + 7 || x1[k] - o > lo + 7) return false;
        }
    }
    return true;
}

// Exact 2x: OpenCV switches INTER_LINEAR to INTER_AREA (2x2 mean with rounding).
__global__ void __launch_bounds__(256)
pyr_resize_area2_u8(const uint8_t *__restrict__ src, int spitch, size_t sframe,
                    uint8_t *__restrict__ dst, int dpitch, size_t dframe, int dw, int dh)
{
    const int dx = blockIdx.x * 64 + threadIdx.x;
    const int dy = blockIdx.y * 4 + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sframe + (size_t)(2 * dy) * spitch + 2 * dx;
    int v = (s[0] + s[1] + s[spitch] + s[spitch + 1] + 2) >> 2;
    dst[(size_t)blockIdx.z * dframe + (size_t)dy * dpitch + dx] = (uint8_t)v;
}

int tb_launch_resize(tb_ctx *ctx, const uint8_t *src, int spitch, size_t sframe, uint8_t *dst, int dpitch,
                     size_t dframe, int dw, int dh, const int4 *xtab, const int4 *ytab, bool area2, int n_frames, int bp, int bh,
                     const CUtensorMap *tmap, bool span8)
{
    ProfScope prof__(ctx, "pyr_resize");
    static const bool use_tma = !(getenv("TB_PYR_TMA") && atoi(getenv("TB_PYR_TMA")) == 0);   // A/B knob: 0 forces the LDG-staged kernel
    if (!area2 && tmap && use_tma && bp <= 2 * PYR_BOXW && bh <= 256) {
        const size_t box_bytes = ((size_t)bh * PYR_BOXW + 127) & ~(size_t)127;
        const size_t smem = 2 * PYR_STAGES * box_bytes + (size_t)bh * PYR_TW * sizeof(uint16_t);
        static const bool use_stream = !(getenv("TB_PYR_STREAM") && atoi(getenv("TB_PYR_STREAM")) == 0);   // A/B knob: 0 = H tile + barriers
        if (use_stream) {
            // frames per CTA: at most TB_PYR_FPC (8), fewer on the small levels so that a launch still has TB_PYR_WAVES (6) waves of
            // 4 CTAs per SM (the last levels of a 1080p batch have ~30 tiles per frame)
            static const int fpc_max = getenv("TB_PYR_FPC") ? std::max(1, atoi(getenv("TB_PYR_FPC"))) : PYR_FPC;
            static const int waves = getenv("TB_PYR_WAVES") ? std::max(1, atoi(getenv("TB_PYR_WAVES"))) : 6;
            const long tiles = (long)div_up(dw, PYR_TW) * div_up(dh, PYR_TH);
            const int fpc = (int)std::min<long>(fpc_max, std::max<long>(1, tiles * n_frames / (148L * 4 * waves)));
            dim3 grid(div_up(dw, PYR_TW), div_up(dh, PYR_TH), div_up(n_frames, fpc)), block(64, 4);
            static const int occ = getenv("TB_PYR_OCC") ? atoi(getenv("TB_PYR_OCC")) : 4;   // resident CTAs per SM the register budget allows (tuning knob)
            static const bool want8 = !(getenv("TB_PYR_COLS") && atoi(getenv("TB_PYR_COLS")) == 4);   // A/B knob: 4 = four columns per thread
            const size_t sm = 2 * PYR_STAGES * box_bytes;
            if (want8 && span8) {
                dim3 block8(32, 8);
                if (occ >= 4) pyr_resize_linear_u8_stream8<4><<<grid, block8, sm, ctx->stream>>>(*tmap, dst, dpitch, dframe, dw, dh, xtab, ytab, bh, n_frames, fpc);
                else pyr_resize_linear_u8_stream8<3><<<grid, block8, sm, ctx->stream>>>(*tmap, dst, dpitch, dframe, dw, dh, xtab, ytab, bh, n_frames, fpc);
            } else if (occ >= 4) pyr_resize_linear_u8_stream<4><<<grid, block, sm, ctx->stream>>>(*tmap, dst, dpitch, dframe, dw, dh, xtab, ytab, bh, n_frames, fpc);
            else pyr_resize_linear_u8_stream<3><<<grid, block, sm, ctx->stream>>>(*tmap, dst, dpitch, dframe, dw, dh, xtab, ytab, bh, n_frames, fpc);
            TB_LAUNCH_CHECK(ctx);
            return TB_OK;
        }
        if (smem <= 200 * 1024) {
            static const int fpc = getenv("TB_PYR_FPC") ? std::max(1, atoi(getenv("TB_PYR_FPC"))) : PYR_FPC;   // frames per CTA (tuning knob)
            dim3 grid(div_up(dw, PYR_TW), div_up(dh, PYR_TH), div_up(n_frames, fpc)), block(64, 4);
            pyr_resize_linear_u8_tma<<<grid, block, smem, ctx->stream>>>(*tmap, dst, dpitch, dframe, dw, dh, xtab, ytab, bh, n_frames, fpc);
            TB_LAUNCH_CHECK(ctx);
            return TB_OK;
        }
    }
    if (area2) {
        dim3 grid(div_up(dw, 64), div_up(dh, 4), n_frames), block(64, 4);
        pyr_resize_area2_u8<<<grid, block, 0, ctx->stream>>>(src, spitch, sframe, dst, dpitch, dframe, dw, dh);
    } else {
        const size_t smem = (size_t)bp * bh + (size_t)bh * PYR_TW * sizeof(uint16_t);
        if (smem > 200 * 1024) return tb_fail(ctx, TB_E_UNSUPPORTED, "resize: scale factor too small for the shared-memory tile");
        dim3 grid(div_up(dw, PYR_TW), div_up(dh, PYR_TH), n_frames), block(64, 4);
        pyr_resize_linear_u8<<<grid, block, smem, ctx->stream>>>(src, spitch, sframe, dst, dpitch, dframe, dw, dh, xtab, ytab, bp, bh);
    }
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

void tb_pyramid_init_device()
{
    cudaFuncSetAttribute(pyr_resize_linear_u8, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pyr_resize_linear_u8_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pyr_resize_linear_u8_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pyr_resize_linear_u8_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pyr_resize_linear_u8_stream8<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pyr_resize_linear_u8_stream8<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
// tensor map over a level's frames for the TMA resize (box = PYR_BOXW x box_rows); false if the layout does not qualify
bool tb_resize_tensor_map(CUtensorMap *tm, const uint8_t *src, int w, int h, int n_frames, int pitch, size_t frame_stride, int box_rows)
{
    return tb_encode_tiled_u8_3d(tm, src, w, h, n_frames, (size_t)pitch, frame_stride, PYR_BOXW, box_rows);
}
