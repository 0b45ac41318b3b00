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
// One CTA handles FAST_G consecutive cells of one cell row: the u8 tile (+halo) is staged once in
// shared memory with 4-byte coalesced loads; corners are compacted to a shared work list so the
// expensive score runs on dense lanes; survivors are appended to the (frame, level) candidate
// array with one global atomic per CTA.  Candidate ORDER in memory is arbitrary: every consumer
// derives the reference's order (cell row, cell col, y, x) from the coordinates themselves.
#include "common.cuh"
#include "kernels.cuh"


__device__ __forceinline__ int min3i(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ int max3i(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// ring offsets (dx,dy) k = 0..15, same order as OpenCV's makeOffsets(pixel, step, 16)
__device__ __constant__ int8_t c_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__device__ __constant__ int8_t c_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

#define RING_LOAD(p, tp) { \
    r[0] = (p)[3 * (tp)];       r[1] = (p)[3 * (tp) + 1];   r[2] = (p)[2 * (tp) + 2];   r[3] = (p)[(tp) + 3];      \
    r[4] = (p)[3];              r[5] = (p)[-(tp) + 3];      r[6] = (p)[-2 * (tp) + 2];  r[7] = (p)[-3 * (tp) + 1]; \
    r[8] = (p)[-3 * (tp)];      r[9] = (p)[-3 * (tp) - 1];  r[10] = (p)[-2 * (tp) - 2]; r[11] = (p)[-(tp) - 3];    \
    r[12] = (p)[-3];            r[13] = (p)[(tp) - 3];      r[14] = (p)[2 * (tp) - 2];  r[15] = (p)[3 * (tp) - 1]; }

// 9 contiguous set bits in a 16-bit circular mask
__device__ __forceinline__ bool arc9(uint32_t m16)
{
    uint32_t x = m16 | (m16 << 16);
    uint32_t t = x & (x >> 1);
    t &= t >> 2;
    t &= t >> 4;     // runs >= 8
    t &= x >> 8;     // runs >= 9
    return t != 0;
}

// m = max(A,B) for a pixel known to be a corner (OpenCV cornerScore<16> without the threshold floor)
__device__ __forceinline__ int fast9_score_m(const uint8_t *p, int tp)
{
    int r[16];
    RING_LOAD(p, tp);
    const int v = p[0];
    int d[25];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - r[k];
#pragma unroll
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int A = -256, Bm = 256; // Bm = min over arcs of max
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int a = min3i(min3i(d[k + 1], d[k + 2], d[k + 3]), min3i(d[k + 4], d[k + 5], d[k + 6]), min(d[k + 7], d[k + 8]));
        A = max3i(A, min(a, d[k]), min(a, d[k + 9]));
        int b = max3i(max3i(d[k + 1], d[k + 2], d[k + 3]), max3i(d[k + 4], d[k + 5], d[k + 6]), max(d[k + 7], d[k + 8]));
        Bm = min3i(Bm, max(b, d[k]), max(b, d[k + 9]));
    }
    return max(A, -Bm);
}

extern __shared__ __align__(16) uint8_t fast_smem[];

__global__ void __launch_bounds__(256)
fast9_cells(FastArgs a)
{
    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int ci = blockIdx.y;          // cell row
    const int cj0 = blockIdx.x * a.G;   // first cell column
    const int y0 = 16 + ci * a.hCell;   // ROI origin (iniY)
    const int x0 = 16 + cj0 * a.wCell;  // ROI origin (iniX of first cell)
    // interior extents handled by this CTA
    const int hEff = min(a.hCell, a.maxBY - 3 - (y0 + 3));
    const int wEff = min(a.scw, a.maxBX - 3 - (x0 + 3));
    if (hEff <= 0 || wEff <= 0) return;
    const int xa = x0 & ~3;             // aligned tile origin
    const int ox = x0 - xa;
    const int rows = hEff + 6;
    const int tp = a.tpitch;

    uint8_t *tile = fast_smem;                                   // (hCell+6) * tp
    uint8_t *sc = tile + (size_t)(a.hCell + 6) * tp;             // hCell * scw  (m map, 0 = not a corner)
    uint16_t *list = reinterpret_cast<uint16_t *>(sc + (size_t)((a.hCell * a.scw + 3) & ~3)); // hCell*scw
    __shared__ int s_n, s_emit, s_base;
    __shared__ int s_has_init[64];

    const uint8_t *img = a.img + (size_t)f * a.frame_stride;
    // ---- stage tile: 4-byte loads, rows [y0, y0+rows), bytes [xa, x0+wEff+6)
    const int words = (ox + wEff + 6 + 3) >> 2;
    for (int i = tid; i < rows * words; i += 256) {
        int rr = i / words, ww = i - rr * words;
        uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(img + (size_t)(y0 + rr) * a.pitch + xa) + ww);
        *reinterpret_cast<uint32_t *>(tile + rr * tp + 4 * ww) = v;
    }
    for (int i = tid; i < (a.hCell * a.scw + 3) / 4; i += 256) reinterpret_cast<uint32_t *>(sc)[i] = 0;
    if (tid == 0) { s_n = 0; s_emit = 0; }
    if (tid < 64) s_has_init[tid] = 0;
    __syncthreads();

    // ---- phase 1a: corner test at minTh, compact corners to the work list
    const int npx = hEff * wEff;
    const int T = a.minTh;
    for (int base = 0; base < npx; base += 256) {
        int i = base + tid;
        bool corner = false;
        int rr = 0, cc = 0;
        if (i < npx) {
            rr = i / wEff; cc = i - rr * wEff;
            const uint8_t *p = tile + (rr + 3) * tp + (cc + 3 + ox);
            const int v = p[0], lo = v - T, hi = v + T;
            // antipodal pairs: every 9-arc contains one pixel of each pair
            int p0 = p[3 * tp], p8 = p[-3 * tp];
            if ((p0 < lo || p0 > hi || p8 < lo || p8 > hi)) {
                int p4 = p[3], p12 = p[-3];
                if ((p4 < lo || p4 > hi || p12 < lo || p12 > hi)) {
                    int r[16];
                    RING_LOAD(p, tp);
                    uint32_t mb = 0, md = 0;
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        mb |= (r[k] > hi ? 1u : 0u) << k;   // ring brighter  (d < -T)
                        md |= (r[k] < lo ? 1u : 0u) << k;   // ring darker    (d >  T)
                    }
                    corner = arc9(mb) || arc9(md);
                }
            }
        }
        unsigned bal = __ballot_sync(0xffffffffu, corner);
        if (bal) {
            int lane = tid & 31, wbase = 0;
            if (lane == 0) wbase = atomicAdd(&s_n, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (corner) list[wbase + __popc(bal & ((1u << lane) - 1))] = (uint16_t)(rr * a.scw + cc);
        }
    }
    __syncthreads();
    const int nc = s_n;
    // ---- phase 1b: score of the corners (dense lanes)
    for (int i = tid; i < nc; i += 256) {
        int idx = list[i];
        int rr = idx / a.scw, cc = idx - rr * a.scw;
        const uint8_t *p = tile + (rr + 3) * tp + (cc + 3 + ox);
        sc[idx] = (uint8_t)fast9_score_m(p, tp);
    }
    __syncthreads();
    // ---- phase 2: NMS inside the cell interior; per-cell "has a survivor above initTh"
    for (int i = tid; i < nc; i += 256) {
        int idx = list[i];
        int rr = idx / a.scw, cc = idx - rr * a.scw;
        int m = sc[idx];
        int cell = cc / a.wCell;
        int cl = cell * a.wCell, chi = min(cl + a.wCell, wEff); // [cl, chi) = this cell's interior columns
        bool keep = true;
#pragma unroll
        for (int dr = -1; dr <= 1; dr++) {
            int r2 = rr + dr;
            if (r2 < 0 || r2 >= hEff) continue;
#pragma unroll
            for (int dc = -1; dc <= 1; dc++) {
                if (!dr && !dc) continue;
                int c2 = cc + dc;
                if (c2 < cl || c2 >= chi) continue;
                int m2 = sc[r2 * a.scw + c2];
                // compare scores (m-1), non-corners are 0 in OpenCV's buffer: m-1 > max(m2-1, 0)
                if (!(m - 1 > max(m2 - 1, 0))) keep = false;
            }
        }
        if (keep) {
            if (m > a.initTh) s_has_init[cell] = 1;
        } else {
            list[i] = 0xFFFF;
        }
    }
    __syncthreads();
    // ---- phase 3: apply the cell threshold; stage survivors in shared memory (the tile is dead now),
    //      reserve space in the (frame, level) candidate array with ONE global atomic, copy out.
    uint32_t *stage = reinterpret_cast<uint32_t *>(tile);
    for (int base = 0; base < nc; base += 256) {
        int i = base + tid;
        bool emit = false;
        uint32_t val = 0;
        if (i < nc && list[i] != 0xFFFF) {
            int idx = list[i];
            int rr = idx / a.scw, cc = idx - rr * a.scw;
            int m = sc[idx];
            int cell = cc / a.wCell;
            int Tc = s_has_init[cell] ? a.initTh : a.minTh;
            if (m > Tc) {
                emit = true;
                // coordinates relative to (16,16): ROI-local (cc+3, rr+3) + cell offset
                val = cand_pack(cj0 * a.wCell + cc + 3, ci * a.hCell + rr + 3, m);
            }
        }
        unsigned bal = __ballot_sync(0xffffffffu, emit);
        if (bal) {
            int lane = tid & 31, wbase = 0;
            if (lane == 0) wbase = atomicAdd(&s_emit, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (emit) stage[wbase + __popc(bal & ((1u << lane) - 1))] = val;
        }
    }
    __syncthreads();
    const int cnt = s_emit;
    if (cnt == 0) return;
    if (tid == 0) s_base = atomicAdd(&a.n_cand[(size_t)f * a.n_cand_stride], cnt);
    __syncthreads();
    const int gbase = s_base;
    if (gbase + cnt > a.cand_cap) { if (tid == 0) atomicExch(a.overflow, 1); }
    for (int i = tid; i < cnt; i += 256)
        if (gbase + i < a.cand_cap) a.cand[(size_t)f * a.cand_cap + gbase + i] = stage[i];
}

int tb_launch_fast(tb_ctx *ctx, const FastArgs &a, int nColsRun, int nRowsRun, int n_frames, size_t smem)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(fast9_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set = true;
    }
    ProfScope prof__(ctx, "fast9_cells");
    dim3 grid(div_up(nColsRun, a.G), nRowsRun, n_frames);
    fast9_cells<<<grid, 256, smem, ctx->stream>>>(a);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}
