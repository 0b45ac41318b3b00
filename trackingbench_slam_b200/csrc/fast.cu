// fast.cu — K2: FAST-9/16 detection per ORB grid cell with per-cell threshold fallback and
// non-maximum suppression.  Replaces the cell loop of ORBExtractor::ComputeKeyPointsOctTree
// (reference src/extractors/ORBextractor.cpp:743-804) and the cv::FAST calls at :785-792.
//
// Facts used (each verified against cv2 4.13 by the oracle tests):
//  * cv::FAST's score is threshold independent: m = max(A,B), A = max over the 16 arcs of the min of
//    d_k = I(p) - I(ring_k) over 9 contiguous k, B likewise for -d_k; corner at T  <=>  m > T;
//    response = m - 1.
//  * NMS keeps strict 8-neighbourhood maxima of the score buffer, in which non-corners and pixels
//    outside the ROI interior are 0.  A non-corner neighbour can never suppress a corner
//    (its m-1 < T <= score), so NMS only needs m of pixels that are corners at minTh.
//  * Cell ROIs are [ini, ini+cell+6) with a 3-px untested border, so cell INTERIORS tile the level
//    exactly with period (wCell, hCell); a cell is self-contained: tile = interior + 3-px ring halo.
//  * A cell uses initTh if any NMS survivor has m > initTh, else minTh.
// One CTA (four warps) handles G consecutive cells of one cell row: the u8 tile (+halo) is staged once in
// shared memory by cp.async; a byte-SIMD prefilter thins the pixels into per-warp queues so that the exact
// arc score runs on dense lanes; corners go to per-warp lists, the warp that found a corner does its NMS,
// survivors are appended to the (frame, level) candidate array with one global atomic per warp and pass.
// Candidate ORDER in memory is arbitrary: every consumer derives the reference's order (cell row, cell col,
// y, x) from the coordinates themselves.
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>


__device__ __forceinline__ int min3i(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ int max3i(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// ring pixel k = 0..15 at (dx,dy): (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
// (same order as OpenCV's makeOffsets(pixel, step, 16))
#define RING_LOAD(p, tp) { \
    r[0] = (p)[3 * (tp)];       r[1] = (p)[3 * (tp) + 1];   r[2] = (p)[2 * (tp) + 2];   r[3] = (p)[(tp) + 3];      \
    r[4] = (p)[3];              r[5] = (p)[-(tp) + 3];      r[6] = (p)[-2 * (tp) + 2];  r[7] = (p)[-3 * (tp) + 1]; \
    r[8] = (p)[-3 * (tp)];      r[9] = (p)[-3 * (tp) - 1];  r[10] = (p)[-2 * (tp) - 2]; r[11] = (p)[-(tp) - 3];    \
    r[12] = (p)[-3];            r[13] = (p)[(tp) - 3];      r[14] = (p)[2 * (tp) - 2];  r[15] = (p)[3 * (tp) - 1]; }

// Same m, both polarities at once in 16-bit lanes: lane 0 = v - ring + 256, lane 1 = ring - v + 256 (both in
// [1,511], so one 32-bit subtraction per ring pixel forms both without borrows), min/max networks on
// VIMNMX3.U16x2.  m = max(lane0, lane1) - 256.
__device__ __forceinline__ int fast9_score_packed(const uint8_t *p, int tp)
{
    int r[16];
    RING_LOAD(p, tp);
    const uint32_t v = p[0];
    // P[k] = (v - ring + 256) | (ring - v + 256) << 16 as ONE multiply-add per ring pixel (FMA pipe; the kernel is bound by
    // the half-rate integer ALU pipe): ring * 0xFFFF adds ring << 16 and subtracts ring from the low lane
    const uint32_t C = (v | ((255u - v) << 16)) + 0x01000100u - 0x00FF0000u;
    uint32_t P[18];
#pragma unroll
    for (int k = 0; k < 16; k++) P[k] = (uint32_t)r[k] * 0x0000FFFFu + C;
    P[16] = P[0]; P[17] = P[1];
    // minimum of every 9-arc = min3 of three overlapping triple minima; 16 + 16 + 8 three-input ops
    uint32_t m3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m3[k] = __vimin3_u16x2(P[k], P[k + 1], P[k + 2]);
    uint32_t best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t w0 = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const uint32_t w1 = __vimin3_u16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_u16x2(best, w0, w1);
    }
    return max((int)(best & 0xFFFFu), (int)(best >> 16)) - 256;
}

extern __shared__ __align__(16) uint8_t fast_smem[];

// a + b issued as a multiply-add: the kernel is bound by the half-rate integer ALU pipe (LOP3 / VIMNMX / IADD), the FMA
// pipe that executes IMAD is mostly idle
__device__ __forceinline__ uint32_t add_fma_pipe(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// the same with the multiplier 1 read from a kernel argument: ptxas turns `mad x, 1, uniform` back into an ALU-pipe VIADD
__device__ __forceinline__ uint32_t add_fma_pipe1(uint32_t a, uint32_t b, uint32_t one)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}
// |ring - centre| > T per byte, result in bit 7 of each byte (T <= 127; K = (127-T)*0x01010101)
__device__ __forceinline__ uint32_t absdiff_gt(uint32_t ring, uint32_t c, uint32_t K)
{
    const uint32_t ad = __vabsdiffu4(ring, c);
    return add_fma_pipe(ad & 0x7f7f7f7fu, K) | ad;
}
// compass test for one word of 4 centres: some N/S pixel AND some E/W pixel differ from the centre by more than T.  A 9-arc
// always contains two adjacent compass pixels of ONE polarity; the polarity-free form lets through about a tenth more pixels
// (18.0 % instead of 16.4 % at T = 20 on the bench stream) at less than half the instructions (19 instead of 42 per word), and
// the arc score decides exactly either way.  (Round-2 builds up to the third session carried the signed form; measured slower
// in both passes once pass 2 walked only the groups it needs.)
__device__ __forceinline__ uint32_t compass_unsigned(uint32_t n, uint32_t s_, uint32_t e, uint32_t w_, uint32_t c, uint32_t K, uint32_t one)
{
    const uint32_t dN = __vabsdiffu4(n, c), dS = __vabsdiffu4(s_, c), dE = __vabsdiffu4(e, c), dW = __vabsdiffu4(w_, c);
    const uint32_t tN = add_fma_pipe1(dN & 0x7f7f7f7fu, K, one), tS = add_fma_pipe1(dS & 0x7f7f7f7fu, K, one);
    const uint32_t tE = add_fma_pipe1(dE & 0x7f7f7f7fu, K, one), tW = add_fma_pipe1(dW & 0x7f7f7f7fu, K, one);
    return (tN | dN | tS | dS) & (tE | dE | tW | dW);
}

// i / n for small i via multiply-high; magic = 2^32/n rounded up (n == 1 has no 32-bit magic)
__device__ __forceinline__ uint32_t div_magic(int n) { return n > 1 ? 0xFFFFFFFFu / (uint32_t)n + 1u : 0u; }
__device__ __forceinline__ int fast_div(int i, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)i, magic) : i; }

// Pipeline of one CTA (G cells of one cell row), every stage on dense lanes:
//   0  stage the u8 tile (+3-px halo) by cp.async (16- or 8-byte units)
//   then, at T = initTh for every cell and again at T = minTh only for the cells that produced no
//   keypoint (the reference's two cv::FAST calls, ORBextractor.cpp:785-792) — stages 1-4 are fused per
//   warp (private survivor queue, no block barrier in between):
//   1+2  SWAR prefilter, 16 px per thread: a FAST-9 corner at T needs two ADJACENT compass pixels
//      (N,E,S,W at radius 3) that differ from the centre by more than T; per 4 px and compass pixel:
//      VABSDIFF4 + 3 ops for |d| > T (the polarity is left to the arc score)
//   3+4  arc score m (both polarities packed in 16-bit lanes, VIMNMX3) of every survivor; corner <=> m > T -> m map,
//      corner appended to the warp's own list (warp-uniform counter in a register: no atomics)
//   5  NMS inside the cell interior by the warp that found the corner, survivors collected in the warp's (dead)
//      queue and written to the candidate array with one global atomic per warp and pass
// The tile pitch is a compile-time constant: queue and list entries are byte offsets into the tile, the m map has the
// tile's pitch (plus a zero row above and below), so ring and neighbour reads are LDS with immediate offsets from one
// register.
__device__ uint2 g_bitpos[256];   // set-bit positions of a byte (one byte each, lowest first), filled by tb_fast_init_device
#define FAST_NT TB_FAST_NT   // threads per CTA
#define FAST_NW (FAST_NT / 32)
static_assert(TB_FAST_TPITCH % 16 == 0 && TB_FAST_TPITCH / 16 <= 32, "the pass-2 group list is built by one warp");
#define FAST_QCAP 544   // per-warp survivor queue: up to 31 waiting + 512 new per iteration
#define FAST_TP TB_FAST_TPITCH
struct FastSmem {
    uint8_t *tile, *scp, *colInfo, *need;   // scp[o] = m of the pixel at tile offset o (0 = not a corner)
    uint16_t *queue, *listB;   // per-warp queues (FAST_NW x FAST_QCAP), per-warp corner lists (FAST_NW x capW entries)
};

// NMS of one corner (tile offset o) + emission bookkeeping; returns keep and the packed candidate
__device__ __forceinline__ bool fast_nms_one(const FastArgs &a, const FastSmem &S, const int o, const int T, const int hEff,
                                             const int xbase, const int cand_y0, uint8_t *s_has, uint32_t &val)
{
    constexpr int tp = FAST_TP;
    const uint8_t *q = S.scp + o;
    const int m = q[0];
    const int trow = (int)__umulhi((uint32_t)o, 0xFFFFFFFFu / tp + 1u);   // o / tp (o < 65536)
    const int col = o - trow * tp;
    const int info = S.colInfo[col];          // cell | has-left-neighbour << 6 | has-right-neighbour << 7 (inside the cell's ROI)
    const int lmask = (info & 0x40) ? 0xFF : 0, rmask = (info & 0x80) ? 0xFF : 0;
    // OpenCV's buffer holds m-1 for corners and 0 elsewhere: keep iff m-1 > max(m2-1, 0) for every neighbour inside the
    // cell's ROI interior, i.e. m > (largest in-range neighbour) and, if there is any neighbour, m > 1. Rows above and
    // below the interior are zero rows of the map.
    const int n0 = q[-tp - 1] & lmask, n1 = q[-tp], n2 = q[-tp + 1] & rmask;
    const int n3 = q[-1] & lmask, n4 = q[1] & rmask;
    const int n5 = q[tp - 1] & lmask, n6 = q[tp], n7 = q[tp + 1] & rmask;
    const int mx = max3i(max3i(n0, n1, n2), max3i(n3, n4, n5), max(n6, n7));
    bool keep = m > mx;
    if (T < 1) keep = keep && (m > 1 || !(trow > 3 || trow + 1 < hEff + 3 || (info & 0xC0)));
    if (keep) {
        s_has[info & 63] = 1;
        val = cand_pack(xbase + col, cand_y0 + trow, m);   // coordinates relative to (16,16): ROI-local (cc+3, rr+3) + cell offset
    }
    return keep;
}

// NMS survivors of a warp are collected in its (by then dead) survivor queue and written out with one global atomic per
// warp and pass (the round trip of an atomic per NMS round showed up as the second largest stall of the kernel)
#define FAST_STAGE_CAP (FAST_QCAP / 2)     // u32 entries that fit a warp's queue
struct FastEmit {
    uint32_t *stage; int n;                // warp-uniform fill
    __device__ __forceinline__ void flush(const FastArgs &a, const int f)
    {
        if (n == 0) return;
        const int lane = threadIdx.x & 31;
        int base = 0;
        __syncwarp();
        if (lane == 0) base = atomicAdd(&a.n_cand[(size_t)f * a.n_cand_stride], n);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < n; i += 32) {
            if (base + i < a.cand_cap) a.cand[(size_t)f * a.cand_cap + base + i] = stage[i];
            else atomicExch(a.overflow, 1);
        }
        __syncwarp();
        n = 0;
    }
    __device__ __forceinline__ void push(const FastArgs &a, const int f, const bool keep, const uint32_t val)
    {
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (bal == 0) return;
        const int cnt = __popc(bal);
        if (n + cnt > FAST_STAGE_CAP) flush(a, f);
        if (keep) stage[n + __popc(bal & ((1u << (threadIdx.x & 31)) - 1))] = val;
        n += cnt;
    }
};

__device__ __forceinline__ void fast_pass(const FastArgs &a, const FastSmem &S, const int T, const int f,
                                          const int hEff, const int wEff, const int cs, const int capW,
                                          const int xbase, const int cand_y0, uint8_t *s_has, int *s_ovf,
                                          const int col_lo, const int col_hi,   // interior columns [col_lo, col_hi) to test
                                          const uint8_t *glist, const int n_glist,  // pass 2: the 16-pixel groups that touch a cell to redo
                                          const bool last_pass)                     // nothing follows: no block barrier at the end
{
    constexpr int tp = FAST_TP;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint8_t *tile = S.tile;
    uint8_t *scp = S.scp;
    uint16_t *myList = S.listB + warp * capW;
    int nCorn = 0;                                      // warp-uniform: corners this warp found in this pass
    const unsigned ltmask = (1u << lane) - 1;
    // ---- stages 1-4, fused per warp (no block barrier, no block-wide list): every warp prefilters 512 pixels
    //      per iteration (need[] holds 0x80 for the bytes this pass must test), pushes the survivors into its own
    //      shared-memory queue and, whenever 32 are waiting, scores them on dense lanes.
    {
        uint16_t *q = S.queue + warp * FAST_QCAP;
        int qn = 0;                                     // warp-uniform queue fill
        const bool swar = T <= 127;                     // larger thresholds skip the byte-SIMD test: everything is scored
        const uint32_t K = (uint32_t)(127 - (swar ? T : 0)) * 0x01010101u;
        const uint32_t one = a.one;
        const int g0 = (cs + col_lo) >> 4, nG = glist ? n_glist : ((cs + col_hi + 15) >> 4) - g0;      // 16-pixel groups
        const uint32_t magic = (!glist && col_lo == 0 && col_hi == a.scw && a.magic_groups != 0xFFFFFFFFu) ? a.magic_groups : div_magic(nG);
        const int items = hEff * nG;
        // score 32 (or, for the final flush, fewer) queued pixels; corners go to the warp's corner list
        auto score_batch = [&](int o, bool valid) {
            bool corner = false;
            if (valid) {
                const int m = fast9_score_packed(tile + o, tp);
                corner = m > T;
                if (corner) scp[o] = (uint8_t)m;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, corner);
            const int pos = nCorn + __popc(bal & ltmask);
            if (corner && pos < capW) myList[pos] = (uint16_t)o;
            nCorn += __popc(bal);
        };
        for (int base = 0; base < items; base += FAST_NT) {
            const int it = base + tid;
            uint32_t c[4] = {0, 0, 0, 0};
            int e0 = 0;
            if (it < items) {
                const int r = fast_div(it, magic);
                const int o = 16 * (glist ? (int)glist[it - r * nG] : it - r * nG + g0);
                e0 = (r + 3) * tp + o;                   // tile offset of the group's first pixel
                const uint4 nd = *reinterpret_cast<const uint4 *>(S.need + o);
                if (nd.x | nd.y | nd.z | nd.w) {
                    if (swar) {
                        const uint8_t *row = tile + e0;
                        const uint4 cen = *reinterpret_cast<const uint4 *>(row);
                        const uint4 up = *reinterpret_cast<const uint4 *>(row - 3 * tp);
                        const uint4 dn = *reinterpret_cast<const uint4 *>(row + 3 * tp);
                        const uint32_t L = *reinterpret_cast<const uint32_t *>(row - 4);
                        const uint32_t R = *reinterpret_cast<const uint32_t *>(row + 16);
                        // west = pixel x-3, east = pixel x+3 of every byte lane
                        c[0] = compass_unsigned(dn.x, up.x, __byte_perm(cen.x, cen.y, 0x6543), __byte_perm(L, cen.x, 0x4321), cen.x, K, one) & nd.x;
                        c[1] = compass_unsigned(dn.y, up.y, __byte_perm(cen.y, cen.z, 0x6543), __byte_perm(cen.x, cen.y, 0x4321), cen.y, K, one) & nd.y;
                        c[2] = compass_unsigned(dn.z, up.z, __byte_perm(cen.z, cen.w, 0x6543), __byte_perm(cen.y, cen.z, 0x4321), cen.z, K, one) & nd.z;
                        c[3] = compass_unsigned(dn.w, up.w, __byte_perm(cen.w, R, 0x6543), __byte_perm(cen.z, cen.w, 0x4321), cen.w, K, one) & nd.w;
                    } else { c[0] = nd.x; c[1] = nd.y; c[2] = nd.z; c[3] = nd.w; }
                }
            }
            // bit 7 of every byte -> one flag bit per pixel (the multiply moves bits 7,15,23,31 to 28..31 without carries)
            // (c[] only has bit 7 of each byte left: need[] holds 0x80 or 0)
            const uint32_t m16 = (c[0] * 0x00204081u >> 28) | ((c[1] * 0x00204081u >> 28) << 4) |
                                 ((c[2] * 0x00204081u >> 28) << 8) | ((c[3] * 0x00204081u >> 28) << 12);
            const int cnt = __popc(m16);
            int inc = cnt;                               // inclusive warp scan of the survivor counts
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            const int total = __shfl_sync(0xffffffffu, inc, 31);
            if (total == 0) continue;                    // nothing in this warp's 512 pixels
            if (cnt) {
                // set-bit positions of each flag byte from a 256-entry table (eight 3-bit positions per entry)
                uint16_t *qp = q + qn + inc - cnt;
                const uint32_t lo = m16 & 0xFF, hi = m16 >> 8;
                const uint2 tl = __ldg(&g_bitpos[lo]), th = __ldg(&g_bitpos[hi]);   // eight byte-wide positions each (2 KB, L1 resident)
                const int nl = __popc(lo), nh = cnt - nl;
                const uint32_t e8 = (uint32_t)(e0 + 8);
                // PRMT pulls position j out zero-extended, the add rides the FMA pipe (the integer ALU pipe is the bottleneck)
#pragma unroll
                for (int j = 0; j < 4; j++) if (j < nl) qp[j] = (uint16_t)add_fma_pipe(__byte_perm(tl.x, 0u, 0x4440 + j), (uint32_t)e0);
#pragma unroll
                for (int j = 4; j < 8; j++) if (j < nl) qp[j] = (uint16_t)add_fma_pipe(__byte_perm(tl.y, 0u, 0x4440 + j - 4), (uint32_t)e0);
                qp += nl;
#pragma unroll
                for (int j = 0; j < 4; j++) if (j < nh) qp[j] = (uint16_t)add_fma_pipe(__byte_perm(th.x, 0u, 0x4440 + j), e8);
#pragma unroll
                for (int j = 4; j < 8; j++) if (j < nh) qp[j] = (uint16_t)add_fma_pipe(__byte_perm(th.y, 0u, 0x4440 + j - 4), e8);
            }
            qn += total;
            __syncwarp();
            while (qn >= 32) {
                qn -= 32;
                score_batch(q[qn + lane], true);
            }
            __syncwarp();                                // queue reads above happen before the next pushes
        }
        if (qn > 0) score_batch(lane < qn ? q[lane] : 0, lane < qn);
        if (nCorn > capW && lane == 0) *s_ovf = 1;
    }
    __syncthreads();
    // ---- stage 5: NMS inside the cell interior. Every warp walks its own corner list (capW entries: an eighth of a
    // quarter of the tile's pixels). If a warp ever finds more corners, the whole block scans the m map instead: slower,
    // same result.
    FastEmit em;
    em.stage = reinterpret_cast<uint32_t *>(S.queue + warp * FAST_QCAP);
    em.n = 0;
    if (*s_ovf == 0) {
        for (int base = 0; base < nCorn; base += 32) {
            const int i = base + lane;
            bool keep = false;
            uint32_t val = 0;
            if (i < nCorn) keep = fast_nms_one(a, S, myList[i], T, hEff, xbase, cand_y0, s_has, val);
            em.push(a, f, keep, val);
        }
    } else {
        const int n_iter = hEff * tp;
        for (int base = 0; base < n_iter; base += FAST_NT) {
            const int i = base + tid;
            bool keep = false;
            uint32_t val = 0;
            if (i < n_iter) {
                // overflow path: pixel i of the m map; only columns this pass tested (pass 2 leaves pass-1 cells alone)
                const int o = i + 3 * tp;
                const int col = i % tp;
                if (col >= cs && col < cs + wEff && S.need[col] != 0 && scp[o] != 0)
                    keep = fast_nms_one(a, S, o, T, hEff, xbase, cand_y0, s_has, val);
            }
            em.push(a, f, keep, val);
        }
    }
    em.flush(a, f);
    if (!last_pass) __syncthreads();
}

__global__ void __launch_bounds__(FAST_NT, TB_FAST_CTAS)
fast9_cells(FastArgs a)
{
    constexpr int tp = FAST_TP;
    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int ci = blockIdx.y;          // cell row
    const int cj0 = blockIdx.x * a.G;   // first cell column
    const int y0 = 16 + ci * a.hCell;   // ROI origin (iniY)
    const int x0 = 16 + cj0 * a.wCell;  // ROI origin (iniX of first cell)
    const int hEff = min(a.hCell, a.maxBY - 3 - (y0 + 3));   // interior rows of this CTA
    const int wEff = min(a.scw, a.maxBX - 3 - (x0 + 3));     // interior columns of this CTA
    if (hEff <= 0 || wEff <= 0) return;
    const int xa = x0 & ~7;             // 8-byte aligned tile origin
    const int ox = x0 - xa;
    const int cs = ox + 3, ce = cs + wEff;   // interior byte range inside a tile row
    const int rows = hEff + 6;
    const int capW = a.list_cap / FAST_NW;

    FastSmem S;
    S.tile = fast_smem;                                              // (hCell+6) * tp
    uint8_t *sc = S.tile + (size_t)(a.hCell + 6) * tp;               // (hCell+2) * tp : m map with a zero row above and below
    S.scp = sc - 2 * tp;                                             // indexed by tile offset: interior row r is tile row r+3, map row r+1
    S.colInfo = sc + (size_t)(a.hCell + 2) * tp;                     // tp : per tile column, cell | lf << 6 | rt << 7
    S.need = S.colInfo + tp;                                         // tp : 0x80 where this pass tests the byte column
    S.queue = reinterpret_cast<uint16_t *>(S.need + tp);
    S.listB = S.queue + FAST_NW * FAST_QCAP;
    __shared__ int s_ovf, s_ng;
    __shared__ uint8_t s_has[64], s_glist[32];

    const uint8_t *img = a.img + (size_t)f * a.frame_stride;
    // ---- stage 0: rows [y0, y0+rows), bytes [xa, x0+wEff+6) copied asynchronously (no registers, every copy of the thread
    //      in flight at once): 16-byte units when the 8-byte aligned tile origin xa is 16-byte aligned, else 8-byte units.
    //      The wait sits behind the zeroing and the column tables below.
    {
        const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(S.tile);
        const uint8_t *src0 = img + (size_t)y0 * a.pitch + xa;
        if ((xa & 15) == 0) {
            const int w16 = (ox + wEff + 6 + 15) >> 4;
            const uint32_t magic16 = div_magic(w16);
            const int total = rows * w16;
            for (int i = tid; i < total; i += FAST_NT) {
                const int rr = fast_div(i, magic16), ww = i - rr * w16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(tile_s + rr * tp + 16 * ww), "l"(src0 + (size_t)rr * a.pitch + 16 * ww) : "memory");
            }
        } else {
            const int w8 = (ox + wEff + 6 + 7) >> 3;
            const uint32_t magic8 = div_magic(w8);
            const int total = rows * w8;
            for (int i = tid; i < total; i += FAST_NT) {
                const int rr = fast_div(i, magic8), ww = i - rr * w8;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(tile_s + rr * tp + 8 * ww), "l"(src0 + (size_t)rr * a.pitch + 8 * ww) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int i = tid; i < (a.hCell + 2) * (tp / 16); i += FAST_NT) reinterpret_cast<uint4 *>(sc)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < tp; i += FAST_NT) {
        const bool in = i >= cs && i < ce;
        S.need[i] = in ? 0x80 : 0;
        int info = 0;
        if (in) {
            const int cc = i - cs;
            const int cell = fast_div(cc, a.magic_wcell);
            const int cl = cell * a.wCell, chi = min(cl + a.wCell, wEff);   // [cl, chi) = this cell's interior columns
            info = cell | (cc > cl ? 0x40 : 0) | (cc + 1 < chi ? 0x80 : 0);
        }
        S.colInfo[i] = (uint8_t)info;
    }
    if (tid == 0) s_ovf = 0;
    for (int i = tid; i < 64; i += FAST_NT) s_has[i] = 0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int xbase = cj0 * a.wCell - cs + 3, cand_y0 = ci * a.hCell;   // candidate x = xbase + tile column, y = cand_y0 + tile row
    // pass 1: every cell at initTh (ORBextractor.cpp:785)
    fast_pass(a, S, a.initTh, f, hEff, wEff, cs, capW, xbase, cand_y0, s_has, &s_ovf, 0, wEff, nullptr, 0, a.minTh == a.initTh);
    // pass 2: cells without a keypoint again at minTh (:788-792)
    if (a.minTh == a.initTh) return;
    const int nCells = (wEff + a.wCell - 1) / a.wCell;     // <= 32 (the host caps G)
    // (pass 1 ended with a block barrier: s_has is final; every warp forms the mask itself, no broadcast and no barrier)
    const unsigned mk = __ballot_sync(0xffffffffu, (tid & 31) < nCells && !s_has[tid & 31]);
    if (mk == 0) return;
    if (tid == 0) s_ovf = 0;
    for (int i = tid; i < tp; i += FAST_NT) S.need[i] = (i >= cs && i < ce && ((mk >> (S.colInfo[i] & 63)) & 1)) ? 0x80 : 0;
    // the 16-pixel groups that hold a column to redo, in order (tp / 16 <= 32 groups: one warp's ballot): a group is
    // needed when one of the cells between its first and its last interior column is in the mask
    if (tid < 32) {
        bool any = false;
        const int lo = max(16 * tid, cs), hi = min(16 * tid + 15, ce - 1);
        if (lo <= hi) {
            const int c0 = S.colInfo[lo] & 63, c1 = S.colInfo[hi] & 63;
            any = ((mk >> c0) & (0xFFFFFFFFu >> (31 - (c1 - c0)))) != 0;
        }
        const unsigned gm = __ballot_sync(0xffffffffu, any);
        if (any) s_glist[__popc(gm & ((1u << tid) - 1))] = (uint8_t)tid;
        if (tid == 0) s_ng = __popc(gm);
    }
    __syncthreads();
    const int n_glist = s_ng;
    fast_pass(a, S, a.minTh, f, hEff, wEff, cs, capW, xbase, cand_y0, s_has, &s_ovf, 0, wEff, s_glist, n_glist, true);
}

size_t tb_fast_smem_bytes(int hCell, int scw, int tpitch, int list_cap)
{
    (void)scw;
    // tile + m map (zero row above and below) + column info + need + per-warp queues + per-warp corner lists
    return (size_t)(hCell + 6) * tpitch + (size_t)(hCell + 2) * tpitch + 2 * (size_t)tpitch +
           sizeof(uint16_t) * FAST_NW * FAST_QCAP + sizeof(uint16_t) * FAST_NW * (size_t)(list_cap / FAST_NW) + 64;
}

int tb_launch_fast(tb_ctx *ctx, const FastArgs &a, int nColsRun, int nRowsRun, int n_frames, size_t smem)
{
    ProfScope prof__(ctx, "fast9_cells");
    dim3 grid(div_up(nColsRun, a.G), nRowsRun, n_frames);
    fast9_cells<<<grid, FAST_NT, smem, ctx->stream>>>(a);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

void tb_fast_init_device()
{
    cudaFuncSetAttribute(fast9_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(fast9_cells, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    uint2 lut[256];
    for (int m = 0; m < 256; m++) {
        unsigned long long v = 0;
        int n = 0;
        for (int b = 0; b < 8; b++) if (m & (1 << b)) v |= (unsigned long long)b << (8 * n++);
        lut[m] = make_uint2((unsigned)(v & 0xFFFFFFFFu), (unsigned)(v >> 32));
    }
    cudaMemcpyToSymbol(g_bitpos, lut, sizeof lut);
}
