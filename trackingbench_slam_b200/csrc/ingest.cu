// ingest.cu — the two image operations in front of the pyramid:
//   PinholeCamera::UndistortImage -> cv::remap(raw, rectified, map1, map2, INTER_LINEAR)   src/camera/CameraModel.cpp:96-102
//   PinholeCamera ctor            -> cv::initUndistortRectifyMap(K, D, I, K, size, CV_32F)  src/camera/CameraModel.cpp:15-18
//   Frame::Equalize               -> cv::createCLAHE(3.0, Size(8,8))->apply                 src/types/Frame.cpp:453-458
// OpenCV's arithmetic is restated bit-exactly (oracle/tb_oracle.c, pinned against cv2 in tests/test_oracle_cv2.py):
//   remap: map*32 rounded half-even to 5 fractional bits, integer parts saturated to int16, bilinear weights
//          (32-fy)(32-fx)*32 out of 2^15, (sum + 2^14) >> 15, BORDER_CONSTANT 0 per tap;
//   CLAHE: per-tile histogram -> clip -> redistribute -> prefix sum * (255/area) -> LUT; per pixel a float bilinear
//          blend of the four surrounding tiles' LUT values, round-half-even.
// Both are byte-in/byte-out image passes: the float maps are converted once into an 8-byte fixed-point record per pixel
// that stays L2-resident while a batch of frames streams through; a remap CTA reads its records once, applies them to up to
// 16 frames, and gets the source pixels as TMA tensor tiles (zero fill outside the image = the constant border).
#include "common.cuh"
#include "kernels.cuh"
#include <cuda.h>
#include <climits>
#include <cstdlib>
#include <cmath>

struct tb_undistort {
    tb_ctx *ctx;
    int w, h;          // destination size (= map size)
    uint2 *fix;        // [h][w]: .x = (ix & 0xFFFF) | (iy << 16) (int16 each), .y = fx | fy << 5
};

int tb_undistort_width(const tb_undistort *u) { return u->w; }
int tb_undistort_height(const tb_undistort *u) { return u->h; }

// cvRound(float) as cvtss2si: NaN and out-of-range values give INT_MIN
__device__ __forceinline__ int cv_round_indef(float v)
{
    if (!(v >= -2147483648.f && v < 2147483648.f)) return INT_MIN;
    return __float2int_rn(v);
}

__global__ void __launch_bounds__(256)
undistort_fix_maps(const float *__restrict__ m1, const float *__restrict__ m2, long long n, uint2 *__restrict__ fix)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int sx = cv_round_indef(__fmul_rn(m1[i], 32.f)), sy = cv_round_indef(__fmul_rn(m2[i], 32.f));
    const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
    fix[i] = make_uint2(((uint32_t)ix & 0xFFFFu) | ((uint32_t)iy << 16), (uint32_t)(sx & 31) | ((uint32_t)(sy & 31) << 5));
}

// one bilinear sample with BORDER_CONSTANT 0 taps straight from global memory (general path)
__device__ __forceinline__ uint32_t remap_gather_px(const uint8_t *__restrict__ s, size_t spitch, int sw, int sh, int ix, int iy, int fx, int fy)
{
    int v00 = 0, v01 = 0, v10 = 0, v11 = 0;
    if ((unsigned)ix < (unsigned)(sw - 1) && (unsigned)iy < (unsigned)(sh - 1)) {   // all four taps inside
        const uint8_t *p = s + (size_t)iy * spitch + ix;
        v00 = __ldg(p); v01 = __ldg(p + 1); v10 = __ldg(p + spitch); v11 = __ldg(p + spitch + 1);
    } else {
        const bool x0in = (unsigned)ix < (unsigned)sw, x1in = (unsigned)(ix + 1) < (unsigned)sw;
        const bool y0in = (unsigned)iy < (unsigned)sh, y1in = (unsigned)(iy + 1) < (unsigned)sh;
        if (y0in) { const uint8_t *p = s + (size_t)iy * spitch; if (x0in) v00 = __ldg(p + ix); if (x1in) v01 = __ldg(p + ix + 1); }
        if (y1in) { const uint8_t *p = s + (size_t)(iy + 1) * spitch; if (x0in) v10 = __ldg(p + ix); if (x1in) v11 = __ldg(p + ix + 1); }
    }
    // weights (32-fy)(32-fx)*32 ... of 2^15; dropping the common factor 32: (sum + 2^9) >> 10, never above 255
    const int sum = (v00 * (32 - fx) + v01 * fx) * (32 - fy) + (v10 * (32 - fx) + v11 * fx) * fy;
    return (uint32_t)((sum + 512) >> 10);
}

#define REMAP_FRAMES 8   // frames that share one read (and decode) of the map records
// General path (any pitch, any map): a thread owns 4 adjacent destination pixels and gathers their taps from global memory.
__global__ void __launch_bounds__(256)
remap_linear_u8(const uint2 *__restrict__ fix, int dw, int dh, const uint8_t *__restrict__ src, int sw, int sh, size_t spitch,
                size_t sframe, uint8_t *__restrict__ dst, size_t dpitch, size_t dframe, int n_frames)
{
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= dw) return;
    const int f0 = blockIdx.z * REMAP_FRAMES, f1 = min(n_frames, f0 + REMAP_FRAMES);
    const int npx = min(4, dw - x0);
    const uint2 *rp = fix + (size_t)y * dw + x0;
    int ixs[4], iys[4], fxs[4], fys[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint2 rec = k < npx ? __ldg(rp + k) : make_uint2(0x80008000u, 0);   // (-32768,-32768): outside
        ixs[k] = (int)(short)(rec.x & 0xFFFFu); iys[k] = (int)(short)(rec.x >> 16);
        fxs[k] = (int)(rec.y & 31u); fys[k] = (int)((rec.y >> 5) & 31u);
    }
    uint8_t *drow = dst + (size_t)y * dpitch + x0;
    const bool vec_store = npx == 4 && ((reinterpret_cast<uintptr_t>(drow) | dframe) & 3) == 0;
    for (int f = f0; f < f1; f++) {
        const uint8_t *s = src + (size_t)f * sframe;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) out |= remap_gather_px(s, spitch, sw, sh, ixs[k], iys[k], fxs[k], fys[k]) << (8 * k);
        uint8_t *d = drow + (size_t)f * dframe;
        if (vec_store) *reinterpret_cast<uint32_t *>(d) = out;
        else for (int k = 0; k < npx; k++) d[k] = (uint8_t)(out >> (8 * k));
    }
}

// ---- TMA path (source pitch and frame stride multiples of 16 bytes).
// A CTA owns a 128 x 16 destination tile for up to RM_FPC frames. It reduces the bounding box of the source taps its map
// records point at; if the box fits RM_BOXW x RM_BOXH (true for lens-undistortion maps), one thread streams that box of
// every frame into shared memory with cp.async.bulk.tensor (3-D tensor map x, y, frame) through an RM_STAGES-deep mbarrier
// ring. Out-of-image parts of the box are zero-filled by the TMA unit, which IS cv::remap's BORDER_CONSTANT 0, so the
// gather needs no border branches and no global load instruction in the inner loop. (The box origin must be 16-byte
// aligned in x — an unaligned origin faults with "illegal instruction" — so it is rounded down and the box is 192 wide.)
#define RM_BOXW 192
#define RM_BOXH 24
#define RM_STAGES 4
#define RM_FPC 16
#define RM_COPY (RM_BOXW * RM_BOXH)
__device__ __forceinline__ uint32_t rm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rm_tma_box(void *dst, const CUtensorMap *tm, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(rm_smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(rm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rm_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(rm_smem_u32(bar)), "r"(parity) : "memory");
}

// out-of-line general path for one thread's 2 x 4 pixels of a tile (keeps the TMA kernel's register count low)
__device__ __noinline__ void remap_tile_general(const uint2 *__restrict__ fix, int dw, int dh, const uint8_t *__restrict__ src, int sw, int sh,
                                                size_t spitch, size_t sframe, uint8_t *__restrict__ dst, size_t dpitch, size_t dframe,
                                                int x0, int yb, int f0, int nf)
{
    for (int h = 0; h < 2; h++) {
        const int y = yb + 8 * h;
        if (y >= dh || x0 >= dw) continue;
        const int npx = min(4, dw - x0);
        int ixs[4], iys[4], fxs[4], fys[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint2 rec = k < npx ? __ldg(fix + (size_t)y * dw + x0 + k) : make_uint2(0x80008000u, 0);
            ixs[k] = (int)(short)(rec.x & 0xFFFFu); iys[k] = (int)(short)(rec.x >> 16);
            fxs[k] = (int)(rec.y & 31u); fys[k] = (int)((rec.y >> 5) & 31u);
        }
        uint8_t *drow = dst + (size_t)y * dpitch + x0;
        const bool vec_store = npx == 4 && ((reinterpret_cast<uintptr_t>(drow) | dframe) & 3) == 0;
        for (int f = f0; f < f0 + nf; f++) {
            const uint8_t *s = src + (size_t)f * sframe;
            uint32_t out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) out |= remap_gather_px(s, spitch, sw, sh, ixs[k], iys[k], fxs[k], fys[k]) << (8 * k);
            uint8_t *d = drow + (size_t)f * dframe;
            if (vec_store) *reinterpret_cast<uint32_t *>(d) = out;
            else for (int k = 0; k < npx; k++) d[k] = (uint8_t)(out >> (8 * k));
        }
    }
}

__global__ void __launch_bounds__(256, 3)
remap_linear_u8_tma(const __grid_constant__ CUtensorMap tmap, const uint2 *__restrict__ fix, int dw, int dh, const uint8_t *__restrict__ src,
                    int sw, int sh, size_t spitch, size_t sframe, uint8_t *__restrict__ dst, size_t dpitch, size_t dframe, int n_frames)
{
    __shared__ alignas(128) uint8_t s_box[RM_STAGES][RM_COPY];
    __shared__ alignas(8) uint64_t s_full[RM_STAGES];
    __shared__ int s_bb[4];
    const int tid = threadIdx.x, lane = tid & 31, wrow = tid >> 5;
    const int x0 = blockIdx.x * 128 + lane * 4, yb = blockIdx.y * 16 + wrow;
    const int f0 = blockIdx.z * RM_FPC, nf = min(n_frames, f0 + RM_FPC) - f0;
    if (tid == 0) {
        s_bb[0] = INT_MAX; s_bb[1] = INT_MAX; s_bb[2] = INT_MIN; s_bb[3] = INT_MIN;
        for (int s = 0; s < RM_STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rm_smem_u32(&s_full[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // pixel k: row yb + 8*(k>>2), column x0 + (k&3)
    int ixs[8], iys[8];
    uint32_t fr[8];        // fx | fy << 5 | live << 10 | sampled << 11
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int x = x0 + (k & 3), y = yb + 8 * (k >> 2);
        const bool live = x < dw && y < dh;
        const uint2 rec = live ? __ldg(fix + (size_t)y * dw + x) : make_uint2(0x80008000u, 0);
        ixs[k] = (int)(short)(rec.x & 0xFFFFu); iys[k] = (int)(short)(rec.x >> 16);
        const bool sampled = live && ixs[k] + 1 >= 0 && ixs[k] < sw && iys[k] + 1 >= 0 && iys[k] < sh;   // at least one tap inside
        fr[k] = (rec.y & 1023u) | (live ? 1024u : 0u) | (sampled ? 2048u : 0u);
        if (sampled) { mnx = min(mnx, ixs[k]); mxx = max(mxx, ixs[k]); mny = min(mny, iys[k]); mxy = max(mxy, iys[k]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0) { atomicMin(&s_bb[0], mnx); atomicMin(&s_bb[1], mny); atomicMax(&s_bb[2], mxx); atomicMax(&s_bb[3], mxy); }
    __syncthreads();
    const int bx0 = s_bb[0] & ~15, by0 = s_bb[1];      // floor to a multiple of 16 (also for negative origins)
    const bool any = s_bb[2] >= s_bb[0];
    const bool fits = any && (s_bb[2] - bx0 <= RM_BOXW - 2) && (s_bb[3] - by0 <= RM_BOXH - 2);
    uint8_t *dbase = dst + (size_t)f0 * dframe + (size_t)yb * dpitch + x0;
    const bool vec_store = x0 + 3 < dw && ((reinterpret_cast<uintptr_t>(dst) | dframe | dpitch) & 3) == 0;
    if (!fits) {   // no source pixel needed at all, or a map too wild for the box: general gathers
        remap_tile_general(fix, dw, dh, src, sw, sh, spitch, sframe, dst, dpitch, dframe, x0, yb, f0, nf);
        return;
    }
    // shared-memory byte offset (within a stage) of the top-left tap of every pixel
    int soff[8];
#pragma unroll
    for (int k = 0; k < 8; k++) soff[k] = (fr[k] & 2048u) ? (iys[k] - by0) * RM_BOXW + (ixs[k] - bx0) : 0;
    if (tid == 0) {
        for (int s = 0; s < RM_STAGES && s < nf; s++) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rm_smem_u32(&s_full[s])), "r"(RM_COPY) : "memory");
            rm_tma_box(&s_box[s][0], &tmap, bx0, by0, f0 + s, &s_full[s]);
        }
    }
    for (int f = 0; f < nf; f++) {
        const int s = f % RM_STAGES;
        rm_mbar_wait(&s_full[s], (uint32_t)((f / RM_STAGES) & 1));
        const uint8_t *box = &s_box[s][0];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t out = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = 4 * h + j;
                const uint8_t *p = box + soff[k];
                const int fx = fr[k] & 31, fy = (fr[k] >> 5) & 31;
                const int sum = ((int)p[0] * (32 - fx) + (int)p[1] * fx) * (32 - fy) + ((int)p[RM_BOXW] * (32 - fx) + (int)p[RM_BOXW + 1] * fx) * fy;
                out |= ((fr[k] & 2048u) ? (uint32_t)((sum + 512) >> 10) : 0u) << (8 * j);
            }
            uint8_t *d = dbase + (size_t)f * dframe + (size_t)(8 * h) * dpitch;
            if (vec_store && (fr[4 * h] & 1024u)) *reinterpret_cast<uint32_t *>(d) = out;
            else for (int j = 0; j < 4; j++) if (fr[4 * h + j] & 1024u) d[j] = (uint8_t)(out >> (8 * j));
        }
        __syncthreads();                                  // every warp is done with stage s
        if (tid == 0 && f + RM_STAGES < nf) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rm_smem_u32(&s_full[s])), "r"(RM_COPY) : "memory");
            rm_tma_box(&s_box[s][0], &tmap, bx0, by0, f0 + f + RM_STAGES, &s_full[s]);
        }
    }
}

// ---- CLAHE
struct ClaheGeom {
    int w, h, tiles_x, tiles_y, tw, th, clip;
    float lutScale, inv_tw, inv_th;
};
__device__ __forceinline__ int reflect101(int p, int len) { return p < len ? p : 2 * len - 2 - p; }   // right/bottom extension only

// one CTA per (tile, frame): histogram of the (reflect-extended) tile, clip + redistribute, LUT
__global__ void __launch_bounds__(256)
clahe_tile_lut(ClaheGeom g, const uint8_t *__restrict__ src, size_t pitch, size_t frame, uint8_t *__restrict__ lut)
{
    __shared__ int s_hist[8][256];
    __shared__ int s_warp[8];
    __shared__ int s_clipped;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, tx = tile % g.tiles_x, ty = tile / g.tiles_x;
    const uint8_t *s = src + (size_t)blockIdx.y * frame;
    for (int i = tid; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int x0 = tx * g.tw, y0 = ty * g.th;
    const bool inside = (x0 + g.tw <= g.w) && (y0 + g.th <= g.h);
    if (inside) {
        // rows of tw bytes: aligned 4-byte words in the middle, bytes at the ragged ends
        for (int r = warp; r < g.th; r += 8) {
            const uint8_t *row = s + (size_t)(y0 + r) * pitch + x0;
            const int head = min(g.tw, (int)((4 - (reinterpret_cast<uintptr_t>(row) & 3)) & 3));
            const int nwords = (g.tw - head) >> 2, tail0 = head + 4 * nwords;
            if (lane < head) atomicAdd(&s_hist[warp][row[lane]], 1);
            for (int wd = lane; wd < nwords; wd += 32) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(row + head) + wd);
                atomicAdd(&s_hist[warp][v & 0xFF], 1); atomicAdd(&s_hist[warp][(v >> 8) & 0xFF], 1);
                atomicAdd(&s_hist[warp][(v >> 16) & 0xFF], 1); atomicAdd(&s_hist[warp][v >> 24], 1);
            }
            if (tail0 + lane < g.tw) atomicAdd(&s_hist[warp][row[tail0 + lane]], 1);
        }
    } else {
        for (int r = warp; r < g.th; r += 8) {
            const uint8_t *row = s + (size_t)reflect101(y0 + r, g.h) * pitch;
            for (int c = lane; c < g.tw; c += 32) atomicAdd(&s_hist[warp][row[reflect101(x0 + c, g.w)]], 1);
        }
    }
    __syncthreads();
    int hv = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) hv += s_hist[k][tid];
    // clip (clahe.cpp: CLAHE_CalcLut_Body)
    int excess = 0;
    if (g.clip > 0 && hv > g.clip) { excess = hv - g.clip; hv = g.clip; }
    int e = excess;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) s_warp[warp] = e;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int k = 0; k < 8; k++) t += s_warp[k]; s_clipped = t; }
    __syncthreads();
    if (g.clip > 0) {
        const int clipped = s_clipped, batch = clipped / 256, residual = clipped - batch * 256;
        hv += batch;
        if (residual != 0) {
            const int step = max(256 / residual, 1);
            if (tid % step == 0 && tid / step < residual) hv++;
        }
    }
    __syncthreads();
    // inclusive prefix sum over the 256 bins
    int inc = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < warp; k++) base += s_warp[k];
    const int sum = base + inc;
    const int r = __float2int_rn(__fmul_rn((float)sum, g.lutScale));
    lut[((size_t)blockIdx.y * g.tiles_x * g.tiles_y + tile) * 256 + tid] = (uint8_t)max(0, min(255, r));
}

#define CLAHE_ROWS 16
// exact byte -> float and round-half-even float -> int for 0 <= value < 2^22 on the full-rate pipes (I2F / F2I issue at a
// quarter of that rate): PRMT drops the byte into the mantissa of 2^23, one FADD removes the bias
__device__ __forceinline__ float byte_to_float(uint32_t q, uint32_t sel) { return __fsub_rn(__uint_as_float(__byte_perm(q, 0x4B000000u, sel)), 8388608.0f); }
__device__ __forceinline__ int rint_small(float v) { return (int)(__float_as_uint(__fadd_rn(v, 12582912.0f)) & 0x3FFFFFu); }

// The four LUT bytes a pixel of value v blends — tiles (ty1,tx1), (ty1,tx2), (ty2,tx1), (ty2,tx2) — packed into one 32-bit
// word for every (row pair, column pair) of the tile grid, clamped pairs included: quad[frame][tiles_y+1][tiles_x+1][256].
__global__ void __launch_bounds__(256)
clahe_quad_lut(ClaheGeom g, const uint8_t *__restrict__ lut, uint32_t *__restrict__ quad)
{
    const int P = g.tiles_x + 1, pc = blockIdx.x % P, rc = blockIdx.x / P, v = threadIdx.x;
    const int tA = max(rc - 1, 0), tB = min(rc, g.tiles_y - 1), xA = max(pc - 1, 0), xB = min(pc, g.tiles_x - 1);
    const uint8_t *fl = lut + (size_t)blockIdx.y * g.tiles_y * g.tiles_x * 256 + v;
    quad[((size_t)blockIdx.y * (g.tiles_y + 1) * P + blockIdx.x) * 256 + v] =
        (uint32_t)fl[(tA * g.tiles_x + xA) * 256] | ((uint32_t)fl[(tA * g.tiles_x + xB) * 256] << 8) |
        ((uint32_t)fl[(tB * g.tiles_x + xA) * 256] << 16) | ((uint32_t)fl[(tB * g.tiles_x + xB) * 256] << 24);
}

// One CTA per CLAHE_ROWS image rows of one frame. The rows of a CTA blend at most two (tile row, tile row + 1)
// combinations; their two rows of the packed table sit in shared memory: one LDS per pixel instead of four. A thread owns 4 adjacent columns,
// computes their column blend factors once and walks down the rows with all rows' pixel loads in flight.
__global__ void __launch_bounds__(256)
clahe_apply(ClaheGeom g, const uint8_t *__restrict__ src, size_t spitch, size_t sframe, const uint32_t *__restrict__ quad,
            uint8_t *__restrict__ dst, size_t dpitch, size_t dframe, int rows_per_cta)
{
    extern __shared__ uint32_t s_quad[];   // [combo 2][pair tiles_x+1][256]
    __shared__ float s_ya[CLAHE_ROWS], s_ya1[CLAHE_ROWS];
    __shared__ int s_rb[CLAHE_ROWS];
    const int tid = threadIdx.x;
    const int y0 = blockIdx.x * rows_per_cta, y1 = min(g.h, y0 + rows_per_cta), nrows = y1 - y0;
    const int P = g.tiles_x + 1;
    const int ty0 = (int)floorf(__fsub_rn(__fmul_rn((float)y0, g.inv_th), 0.5f));     // unclamped: -1 .. tiles_y-1
    for (int c = 0; c < 2; c++) {              // two rows of the packed table, straight 16-byte copies
        const int rc = min(ty0 + 1 + c, g.tiles_y);
        const uint4 *qs = reinterpret_cast<const uint4 *>(quad + ((size_t)blockIdx.y * (g.tiles_y + 1) + rc) * P * 256);
        uint4 *qd = reinterpret_cast<uint4 *>(s_quad + c * P * 256);
        for (int i = tid; i < P * 64; i += 256) qd[i] = __ldg(qs + i);
    }
    if (tid < nrows) {
        const float tyf = __fsub_rn(__fmul_rn((float)(y0 + tid), g.inv_th), 0.5f);
        const int ty1 = (int)floorf(tyf);
        const float ya = __fsub_rn(tyf, (float)ty1);
        s_ya[tid] = ya; s_ya1[tid] = __fsub_rn(1.0f, ya);
        s_rb[tid] = (ty1 - ty0) * P * 256;      // 0 or 1 combos down (rows_per_cta <= tile height)
    }
    __syncthreads();
    const uint8_t *sbase = src + (size_t)blockIdx.y * sframe + (size_t)y0 * spitch;
    uint8_t *dbase = dst + (size_t)blockIdx.y * dframe + (size_t)y0 * dpitch;
    for (int x4 = tid * 4; x4 < g.w; x4 += 1024) {
        float xa[4], xa1[4];
        int po[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float txf = __fsub_rn(__fmul_rn((float)(x4 + k), g.inv_tw), 0.5f);
            const int tx1 = (int)floorf(txf);
            xa[k] = __fsub_rn(txf, (float)tx1); xa1[k] = __fsub_rn(1.0f, xa[k]);
            po[k] = min(tx1 + 1, g.tiles_x) * 256;      // pair (tx1, tx1 + 1) with both ends clamped into the grid
        }
        const int npx = min(4, g.w - x4);
        uint32_t v4[CLAHE_ROWS];
#pragma unroll
        for (int r = 0; r < CLAHE_ROWS; r++) {       // all rows' pixels in flight before the first blend
            v4[r] = 0;
            if (r < nrows) {
                const uint8_t *sp = sbase + (size_t)r * spitch + x4;
                if (npx == 4 && (reinterpret_cast<uintptr_t>(sp) & 3) == 0) v4[r] = __ldg(reinterpret_cast<const uint32_t *>(sp));
                else for (int k = 0; k < npx; k++) v4[r] |= (uint32_t)sp[k] << (8 * k);
            }
        }
#pragma unroll
        for (int r = 0; r < CLAHE_ROWS; r++) {
            if (r >= nrows) break;
            uint8_t *dp = dbase + (size_t)r * dpitch + x4;
            const uint32_t *q0 = s_quad + s_rb[r];
            const float ya = s_ya[r], ya1 = s_ya1[r];
            uint32_t out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t q = q0[po[k] + ((v4[r] >> (8 * k)) & 0xFF)];
                const float a = byte_to_float(q, 0x7650), b = byte_to_float(q, 0x7651), c = byte_to_float(q, 0x7652), d = byte_to_float(q, 0x7653);
                const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(a, xa1[k]), __fmul_rn(b, xa[k])), ya1),
                                            __fmul_rn(__fadd_rn(__fmul_rn(c, xa1[k]), __fmul_rn(d, xa[k])), ya));
                out |= (uint32_t)min(255, rint_small(res)) << (8 * k);
            }
            if (npx == 4 && (reinterpret_cast<uintptr_t>(dp) & 3) == 0) *reinterpret_cast<uint32_t *>(dp) = out;
            else for (int k = 0; k < npx; k++) dp[k] = (uint8_t)(out >> (8 * k));
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
// cv::initUndistortRectifyMap with R = I and newCameraMatrix = K: double arithmetic, inverse by cofactors (cv::invert 3x3)
extern "C" int tb_init_undistort_rectify_map(const tb_pinhole *cam, float *map1, float *map2)
{
    if (!cam || !map1 || !map2 || cam->width <= 0 || cam->height <= 0) return TB_E_INVALID;
    const double A[3][3] = {{cam->fx, 0, cam->cx}, {0, cam->fy, cam->cy}, {0, 0, 1}};
    double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                 A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    if (det == 0) return TB_E_INVALID;
    det = 1.0 / det;
    const double i00 = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) * det, i01 = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * det,
                 i02 = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * det, i10 = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) * det,
                 i11 = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * det, i12 = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * det,
                 i20 = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) * det, i21 = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * det,
                 i22 = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * det;
    const double fx = cam->fx, fy = cam->fy, u0 = cam->cx, v0 = cam->cy;
    const double k1 = cam->d[0], k2 = cam->d[1], p1 = cam->d[2], p2 = cam->d[3], k3 = cam->d[4];
    const double k4 = 0, k5 = 0, k6 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    const int w = cam->width, h = cam->height;
    for (int i = 0; i < h; i++) {
        double _x = i * i01 + i02, _y = i * i11 + i12, _w = i * i21 + i22;
        float *r1 = map1 + (size_t)i * w, *r2m = map2 + (size_t)i * w;
        for (int j = 0; j < w; j++, _x += i00, _y += i10, _w += i20) {
            const double iw = 1. / _w, x = _x * iw, y = _y * iw;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
            const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2);
            const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2);
            r1[j] = (float)(fx * 1.0 * xd + u0);     // tilt model is the identity: invProj = 1
            r2m[j] = (float)(fy * 1.0 * yd + v0);
        }
    }
    return TB_OK;
}

extern "C" int tb_undistort_create(tb_ctx *ctx, int dst_w, int dst_h, const float *map1, const float *map2, tb_undistort **out)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !out || dst_w <= 0 || dst_h <= 0 || !map1 || !map2) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_undistort_create: bad arguments") : TB_E_INVALID;
    const size_t n = (size_t)dst_w * dst_h;
    float *d1 = nullptr, *d2 = nullptr;
    uint2 *fix = nullptr;
    if (cudaMalloc(&d1, n * 4) != cudaSuccess || cudaMalloc(&d2, n * 4) != cudaSuccess || cudaMalloc(&fix, n * 8) != cudaSuccess) {
        cudaFree(d1); cudaFree(d2); cudaFree(fix);
        return tb_fail(ctx, TB_E_NOMEM, "tb_undistort_create: cudaMalloc failed");
    }
    cudaError_t e = cudaMemcpyAsync(d1, map1, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d2, map2, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        undistort_fix_maps<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d1, d2, (long long)n, fix);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d1); cudaFree(d2);
    if (e != cudaSuccess) { cudaFree(fix); return tb_fail(ctx, TB_E_CUDA, "tb_undistort_create: %s", cudaGetErrorString(e)); }
    tb_undistort *u = new tb_undistort();
    u->ctx = ctx; u->w = dst_w; u->h = dst_h; u->fix = fix;
    *out = u;
    return TB_OK;
}

extern "C" void tb_undistort_destroy(tb_undistort *u)
{
    if (!u) return;
    cudaFree(u->fix);
    delete u;
}

extern "C" int tb_undistort_image_device(tb_ctx *ctx, const tb_undistort *u, const void *src, int src_w, int src_h, size_t src_pitch,
                                         size_t src_frame_stride, void *dst, size_t dst_pitch, size_t dst_frame_stride, int n_frames)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !u || !src || !dst || src_w <= 0 || src_h <= 0 || n_frames < 0 || src_pitch < (size_t)src_w || dst_pitch < (size_t)u->w)
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_undistort_image_device: bad arguments") : TB_E_INVALID;
    if (src_w > 32767 || src_h > 32767) return tb_fail(ctx, TB_E_UNSUPPORTED, "remap: source larger than 32767 (int16 coordinates, as cv::remap)");
    if (src == dst) return tb_fail(ctx, TB_E_INVALID, "remap cannot run in place");
    if (n_frames == 0) return TB_OK;
    // TMA path: 3-D tensor map (x, y, frame) over the source; needs 16-byte aligned base, pitch and frame stride
    const size_t fstride = n_frames > 1 ? src_frame_stride : src_pitch * (size_t)src_h;
    static const bool use_tma = !(getenv("TB_REMAP_TMA") && atoi(getenv("TB_REMAP_TMA")) == 0);   // A/B knob: 0 forces the gather kernel
    if (use_tma && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (src_pitch & 15) == 0 && (fstride & 15) == 0 && fstride >= src_pitch * (size_t)src_h) {
        CUtensorMap tm;
        if (tb_encode_tiled_u8_3d(&tm, src, src_w, src_h, n_frames, src_pitch, fstride, RM_BOXW, RM_BOXH)) {
            {
                ProfScope prof__(ctx, "remap_linear_u8_tma");
                dim3 grid(div_up(u->w, 128), div_up(u->h, 16), div_up(n_frames, RM_FPC));
                remap_linear_u8_tma<<<grid, 256, 0, ctx->stream>>>(tm, u->fix, u->w, u->h, (const uint8_t *)src, src_w, src_h, src_pitch, fstride,
                                                                   (uint8_t *)dst, dst_pitch, dst_frame_stride, n_frames);
                TB_LAUNCH_CHECK(ctx);
                return TB_OK;
            }
        }
    }
    ProfScope prof__(ctx, "remap_linear_u8");
    dim3 grid(div_up(div_up(u->w, 4), 256), u->h, div_up(n_frames, REMAP_FRAMES));
    remap_linear_u8<<<grid, 256, 0, ctx->stream>>>(u->fix, u->w, u->h, (const uint8_t *)src, src_w, src_h, src_pitch, src_frame_stride,
                                                   (uint8_t *)dst, dst_pitch, dst_frame_stride, n_frames);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

extern "C" int tb_undistort_image_host(tb_ctx *ctx, const tb_undistort *u, const uint8_t *src, int src_w, int src_h, size_t src_stride,
                                       uint8_t *dst, size_t dst_stride)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !u || !src || !dst || src_w <= 0 || src_h <= 0) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_undistort_image_host: bad arguments") : TB_E_INVALID;
    ArenaScope buf(ctx);
    const size_t sp = align_up((size_t)src_w, 64);
    uint8_t *ds = buf.alloc<uint8_t>(sp * src_h), *dd = buf.alloc<uint8_t>((size_t)u->w * u->h);
    if (!ds || !dd) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    TB_CUDA(ctx, cudaMemcpy2DAsync(ds, sp, src, src_stride, src_w, src_h, cudaMemcpyHostToDevice, ctx->stream));
    int rc = tb_undistort_image_device(ctx, u, ds, src_w, src_h, sp, 0, dd, u->w, 0, 1);
    if (rc) return rc;
    TB_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_stride, dd, u->w, u->w, u->h, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

struct IngestWs { uint8_t *lut = nullptr; uint32_t *quad = nullptr; size_t bytes = 0, quad_bytes = 0; };
void tb_ingest_ws_free(tb_ctx *ctx)
{
    if (!ctx->iws) return;
    cudaFree(ctx->iws->lut); cudaFree(ctx->iws->quad);
    delete ctx->iws;
    ctx->iws = nullptr;
}

extern "C" int tb_clahe_device(tb_ctx *ctx, const void *src, int w, int h, size_t src_pitch, size_t src_frame_stride, void *dst,
                               size_t dst_pitch, size_t dst_frame_stride, int n_frames, double clip_limit, int tiles_x, int tiles_y)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !src || !dst || w <= 0 || h <= 0 || n_frames < 0 || tiles_x < 1 || tiles_y < 1 || src_pitch < (size_t)w || dst_pitch < (size_t)w)
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_clahe_device: bad arguments") : TB_E_INVALID;
    if (n_frames == 0) return TB_OK;
    ClaheGeom g;
    g.w = w; g.h = h; g.tiles_x = tiles_x; g.tiles_y = tiles_y;
    const bool even = (w % tiles_x == 0) && (h % tiles_y == 0);
    const int ew = even ? w : w + (tiles_x - w % tiles_x), eh = even ? h : h + (tiles_y - h % tiles_y);   // clahe.cpp: copyMakeBorder on both axes
    g.tw = ew / tiles_x; g.th = eh / tiles_y;
    if (ew - w >= w || eh - h >= h) return tb_fail(ctx, TB_E_UNSUPPORTED, "CLAHE: image smaller than the tile grid padding");
    const int area = g.tw * g.th;
    volatile float lutScale = (float)255 / (float)area, inv_tw = 1.0f / (float)g.tw, inv_th = 1.0f / (float)g.th;
    g.lutScale = lutScale; g.inv_tw = inv_tw; g.inv_th = inv_th;
    g.clip = 0;
    if (clip_limit > 0.0) { g.clip = (int)(clip_limit * area / 256); if (g.clip < 1) g.clip = 1; }
    if (!ctx->iws) ctx->iws = new IngestWs();
    IngestWs &ws = *ctx->iws;
    const size_t need = (size_t)n_frames * tiles_x * tiles_y * 256;
    const size_t need_quad = (size_t)n_frames * (tiles_x + 1) * (tiles_y + 1) * 256 * sizeof(uint32_t);
    if (need > ws.bytes || need_quad > ws.quad_bytes) {
        TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ws.lut); cudaFree(ws.quad); ws.lut = nullptr; ws.quad = nullptr; ws.bytes = 0; ws.quad_bytes = 0;
        TB_CUDA(ctx, cudaMalloc(&ws.lut, need));
        TB_CUDA(ctx, cudaMalloc(&ws.quad, need_quad));
        ws.bytes = need; ws.quad_bytes = need_quad;
    }
    {
        ProfScope prof__(ctx, "clahe_tile_lut");
        clahe_tile_lut<<<dim3(tiles_x * tiles_y, n_frames), 256, 0, ctx->stream>>>(g, (const uint8_t *)src, src_pitch, src_frame_stride, ws.lut);
        TB_LAUNCH_CHECK(ctx);
        clahe_quad_lut<<<dim3((tiles_x + 1) * (tiles_y + 1), n_frames), 256, 0, ctx->stream>>>(g, ws.lut, ws.quad);
        TB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope prof__(ctx, "clahe_apply");
        const int rows = g.th >= CLAHE_ROWS ? CLAHE_ROWS : 1;     // rows of one CTA see at most two (tile row, tile row + 1) combinations
        const size_t smem = (size_t)2 * (tiles_x + 1) * 256 * sizeof(uint32_t);
        if (smem > 200 * 1024) return tb_fail(ctx, TB_E_UNSUPPORTED, "CLAHE: too many tile columns for the shared-memory LUT window");
        if (smem > 48 * 1024) TB_CUDA(ctx, cudaFuncSetAttribute(clahe_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        clahe_apply<<<dim3(div_up(h, rows), n_frames), 256, smem, ctx->stream>>>(g, (const uint8_t *)src, src_pitch, src_frame_stride, ws.quad,
                                                                                 (uint8_t *)dst, dst_pitch, dst_frame_stride, rows);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

extern "C" int tb_clahe_host(tb_ctx *ctx, const uint8_t *src, int w, int h, size_t src_stride, uint8_t *dst, size_t dst_stride,
                             double clip_limit, int tiles_x, int tiles_y)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !src || !dst || w <= 0 || h <= 0) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_clahe_host: bad arguments") : TB_E_INVALID;
    ArenaScope buf(ctx);
    uint8_t *ds = buf.alloc<uint8_t>((size_t)w * h), *dd = buf.alloc<uint8_t>((size_t)w * h);
    if (!ds || !dd) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    TB_CUDA(ctx, cudaMemcpy2DAsync(ds, w, src, src_stride, w, h, cudaMemcpyHostToDevice, ctx->stream));
    int rc = tb_clahe_device(ctx, ds, w, h, w, 0, dd, w, 0, 1, clip_limit, tiles_x, tiles_y);
    if (rc) return rc;
    TB_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_stride, dd, w, w, h, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// PinholeCamera::UndistortPoint / Frame::UnDistortPoints (src/camera/CameraModel.cpp:104-121, src/types/Frame.cpp:165-171)
// = cv::undistortPoints(src, dst, K, D, Mat(), K): 5 fixed-point iterations of the radtan model in double, re-projection
// with K. One thread per point; every operation is a single IEEE double operation in OpenCV's order (the library is built
// with --fmad=false), bit-exact vs cv2 4.13 through the oracle.
__global__ void __launch_bounds__(256)
undistort_points_kernel(const uint8_t *__restrict__ pts, size_t stride, size_t slot_stride, const int *__restrict__ counts, int cap,
                        tb_pinhole cam, float2 *__restrict__ out)
{
    const int i = blockIdx.x * 256 + threadIdx.x, slot = blockIdx.y;
    const int n = counts ? min(counts[slot], cap) : cap;
    if (i >= cap) return;
    float2 *o = out + (size_t)slot * cap + i;
    if (i >= n) { *o = make_float2(0.f, 0.f); return; }
    const float *p = reinterpret_cast<const float *>(pts + (size_t)slot * slot_stride + (size_t)i * stride);
    const float pu = __ldg(p), pv = __ldg(p + 1);
    if (!(fabsf(cam.d[0]) > 1e-7f)) { *o = make_float2(pu, pv); return; }     // mbDistortion false (CameraModel.cpp:9)
    const double fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
    const double k0 = cam.d[0], k1 = cam.d[1], k2 = cam.d[2], k3 = cam.d[3], k4 = cam.d[4];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double u = pu, v = pv;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
        if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
        const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
        const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + 0. * r2 + 0. * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
    *o = make_float2((float)(xx * ww), (float)(yy * ww));
}

int tb_undistort_points_launch(tb_ctx *ctx, const tb_pinhole *cam, const uint8_t *pts, size_t stride, size_t slot_stride, const int *counts,
                               int cap, int n_slots, float2 *out)
{
    if (cap <= 0 || n_slots <= 0) return TB_OK;
    ProfScope prof__(ctx, "undistort_points");
    undistort_points_kernel<<<dim3(div_up(cap, 256), n_slots), 256, 0, ctx->stream>>>(pts, stride, slot_stride, counts, cap, *cam, out);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

extern "C" int tb_undistort_points_device(tb_ctx *ctx, const tb_pinhole *cam, const void *pts_dptr, size_t stride, long long n, void *out_dptr)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !cam || n < 0 || n > 0x7fffffff || (n && (!pts_dptr || !out_dptr)) || stride < 8)
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_undistort_points_device: bad arguments") : TB_E_INVALID;
    return tb_undistort_points_launch(ctx, cam, (const uint8_t *)pts_dptr, stride, 0, nullptr, (int)n, 1, (float2 *)out_dptr);
}

extern "C" int tb_undistort_points(tb_ctx *ctx, const tb_pinhole *cam, const float *xy_in, int n, float *xy_out)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !cam || n < 0 || (n && (!xy_in || !xy_out))) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_undistort_points: bad arguments") : TB_E_INVALID;
    if (n == 0) return TB_OK;
    ArenaScope buf(ctx);
    float2 *d_in = buf.alloc<float2>((size_t)n), *d_out = buf.alloc<float2>((size_t)n);
    if (!d_in || !d_out) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    TB_CUDA(ctx, cudaMemcpyAsync(d_in, xy_in, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = tb_undistort_points_launch(ctx, cam, (const uint8_t *)d_in, sizeof(float2), 0, nullptr, n, 1, d_out);
    if (rc) return rc;
    TB_CUDA(ctx, cudaMemcpyAsync(xy_out, d_out, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}
