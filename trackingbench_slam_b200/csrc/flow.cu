// flow.cu — pyramidal Lucas-Kanade point tracking: the cv::calcOpticalFlowPyrLK(img2, img1, keys, pts, status, err,
// Size(21,21), 3) call of Matcher::searchByOPFlow (src/matchers/matcher.cpp:745), SURVEY.md 8f rank 3 (second half).
//
// OpenCV's algorithm (video/lkpyramid.cpp), restated in oracle/tb_oracle.c (tbo_lk_flow) and here:
//   pyramid   level l = cv::pyrDown(level l-1) (1-4-6-4-1, (v+128)>>8, REFLECT_101), built while both sides stay > win;
//             image reads outside a level = REFLECT_101 (the border buildOpticalFlowPyramid pads with);
//   Scharr    dx, dy of the previous image (3-10-3), REFLECT_101 inside the level, ZERO outside it (BORDER_CONSTANT pad);
//   per point, coarse to fine: 14-bit fixed-point bilinear patch of I, dIx, dIy (int16), the 2x2 system, then up to
//             max_iter Newton steps on the bilinear patch of J; stop on |delta|^2 <= eps^2 or on an oscillation.
// The sums of the 2x2 system and of the right-hand side are exact integers here and in the oracle (OpenCV's SIMD build
// accumulates them in float lanes), every float operation is a single IEEE operation in the oracle's order (the library
// is built with --fmad=false), so kernel and oracle agree bit for bit; the oracle agrees with cv2 within a tolerance.
//
// Work decomposition: the dense part is only the pyrDown chain (3 small levels per frame). Derivatives are NOT built as
// images (that would be 4 bytes per pixel per level, 11 MB per 1080p frame, for ~2000 x 4 windows of 22x22): each warp
// stages the (win+3)^2 neighbourhood of its point in shared memory and takes the Scharr derivatives it needs from there.
// One warp per point walks all levels, so a track never leaves the SM between levels.
#include "common.cuh"
#include "kernels.cuh"

#define LK_MAX_LEVELS 8
#define LK_MAX_WIN 21
#define LK_WARPS 8
#define LK_REG 32      // side of the staged region of the next image: a (win+1)^2 window can move (32-22)/2 = 5 px each way inside it

struct LkLevel { const uint8_t *p; int w, h, pitch; };
struct LkPair { LkLevel I[LK_MAX_LEVELS], J[LK_MAX_LEVELS]; const uint8_t *pts; const int *count; };

__device__ __forceinline__ int lk_reflect(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// cv::pyrDown: a thread owns one output column of a 16-row strip and slides down it, keeping the horizontal 1-4-6-4-1 sums of the
// last three source rows: 10 byte loads per output instead of 25 (the two new source rows), no shared memory.
#define LK_PD_ROWS 16
__global__ void __launch_bounds__(128)
lk_pyrdown_u8(const uint8_t *__restrict__ src, int sw, int sh, size_t spitch, size_t sframe,
              uint8_t *__restrict__ dst, int dw, int dh, size_t dpitch, size_t dframe)
{
    const int x = blockIdx.x * 128 + threadIdx.x, y0 = blockIdx.y * LK_PD_ROWS;
    if (x >= dw) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sframe;
    uint8_t *d = dst + (size_t)blockIdx.z * dframe + x;
    const int x0 = lk_reflect(2 * x - 2, sw), x1 = lk_reflect(2 * x - 1, sw), x2 = lk_reflect(2 * x, sw), x3 = lk_reflect(2 * x + 1, sw),
              x4 = lk_reflect(2 * x + 2, sw);
    auto hsum = [&](int sy) {
        const uint8_t *r = s + (size_t)lk_reflect(sy, sh) * spitch;
        return __ldg(r + x2) * 6 + (__ldg(r + x1) + __ldg(r + x3)) * 4 + __ldg(r + x0) + __ldg(r + x4);
    };
    int h0 = hsum(2 * y0 - 2), h1 = hsum(2 * y0 - 1), h2 = hsum(2 * y0);
    const int y1 = min(y0 + LK_PD_ROWS, dh);
    for (int y = y0; y < y1; y++) {
        const int h3 = hsum(2 * y + 1), h4 = hsum(2 * y + 2);
        d[(size_t)y * dpitch] = (uint8_t)((h2 * 6 + (h1 + h3) * 4 + h0 + h4 + 128) >> 8);
        h0 = h2; h1 = h3; h2 = h4;
    }
}

// cvFloor as x86 computes it (cvttss2si): NaN and out-of-range values give INT_MIN, which every range check rejects
__device__ __forceinline__ int lk_floor(float v)
{
    return (v >= -2147483648.f && v < 2147483648.f) ? (int)floorf(v) : (-2147483647 - 1);
}
struct LkWeights { int w00, w01, w10, w11; };
__device__ __forceinline__ LkWeights lk_weights(float a, float b)
{
    LkWeights k;
    const float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    k.w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), 16384.f));
    k.w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), 16384.f));
    k.w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), 16384.f));
    k.w11 = 16384 - k.w00 - k.w01 - k.w10;
    return k;
}
// exact warp sum of 32-bit partials: the low 16 bits and the (signed) high part go through redux.sync separately, so
// neither 32-lane sum can overflow; 2 REDUX instead of 5 rounds of 64-bit shuffle + add
__device__ __forceinline__ long long lk_warp_sum(int v)
{
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return (long long)hi * 65536 + (long long)lo;
}

struct LkWarpSmem {
    union {
        uint8_t nb[(LK_MAX_WIN + 3) * (LK_MAX_WIN + 3) + 4];   // (win+3)^2 bytes of the previous image
        uint16_t nb16[LK_REG * LK_REG];                        // 32x32 region of the next image as overlapping pairs: px(x) | px(x+1) << 8
    };
    short2 d[(LK_MAX_WIN + 1) * (LK_MAX_WIN + 1)];             // Scharr (dx, dy) on the (win+1)^2 grid
    int2 dI[LK_MAX_WIN * LK_MAX_WIN];                          // patch of I: (ixval, iyval) and ival, widened to 32 bits once so the
    int I[LK_MAX_WIN * LK_MAX_WIN];                            // iterations need no sign extension (9.3 KB per warp: 3 CTAs per SM)
};

// stage the n x n neighbourhood with origin (x0, y0) of a level (REFLECT_101 outside) into nb
// (n is a compile-time constant in the WIN = 21 instantiation: the index split is a multiply-shift and the loop unrolls)
__device__ __forceinline__ void lk_stage(const LkLevel &L, int x0, int y0, const int n, uint8_t *nb, int lane)
{
    if (x0 >= 0 && y0 >= 0 && x0 + n <= L.w && y0 + n <= L.h) {   // warp-uniform: window inside the level, no reflection
        const uint8_t *b = L.p + (size_t)y0 * L.pitch + x0;
#pragma unroll 6
        for (int i = lane; i < n * n; i += 32) {
            const int yy = i / n, xx = i - yy * n;
            nb[i] = __ldg(b + yy * L.pitch + xx);
        }
        return;
    }
    for (int i = lane; i < n * n; i += 32) {
        const int yy = i / n, xx = i - yy * n;
        nb[i] = __ldg(L.p + (size_t)lk_reflect(y0 + yy, L.h) * L.pitch + lk_reflect(x0 + xx, L.w));
    }
}

// stage the LK_REG x LK_REG region with origin (x0, y0) as overlapping byte pairs (px(x), px(x+1)): a bilinear tap row is
// then one 16-bit load and one dp2a. lane = column, one coalesced row read per step, the right neighbour by shuffle.
__device__ __forceinline__ void lk_stage_region(const LkLevel &L, int x0, int y0, uint16_t *nb16, int lane)
{
    if (x0 >= 0 && y0 >= 0 && x0 + LK_REG + 1 <= L.w && y0 + LK_REG <= L.h) {
        const uint8_t *b = L.p + (size_t)y0 * L.pitch + x0 + lane;
#pragma unroll 8
        for (int r = 0; r < LK_REG; r++) {
            const unsigned v = __ldg(b + r * L.pitch);
            unsigned nx = __shfl_down_sync(0xffffffffu, v, 1);
            if (lane == 31) nx = __ldg(b + r * L.pitch + 1);
            nb16[r * LK_REG + lane] = (uint16_t)(v | (nx << 8));
        }
        return;
    }
    const int xa = lk_reflect(x0 + lane, L.w), xb = lk_reflect(x0 + lane + 1, L.w);
    for (int r = 0; r < LK_REG; r++) {
        const uint8_t *row = L.p + (size_t)lk_reflect(y0 + r, L.h) * L.pitch;
        nb16[r * LK_REG + lane] = (uint16_t)(__ldg(row + xa) | ((unsigned)__ldg(row + xb) << 8));
    }
}
// w_lo * px(x) + w_hi * px(x+1) + acc: signed 16-bit weights (w11 can be -1), unsigned bytes
__device__ __forceinline__ int lk_dp2a(int w_pair, unsigned px_pair, int acc)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w_pair), "r"(px_pair), "r"(acc));
    return d;
}

template <int WIN>   // WIN = 21: the reference's window as a compile-time constant; WIN = 0: p.win at run time
__global__ void __launch_bounds__(LK_WARPS * 32)
lk_track(const LkPair *__restrict__ pairs, tb_lk_params p, int max_level, size_t pts_stride, int pts_cap, int frame_w, int frame_h,
         int check_in_frame, float2 *__restrict__ out_pts, uint8_t *__restrict__ out_status, float *__restrict__ out_err)
{
    extern __shared__ __align__(16) unsigned char lk_smem[];
    LkWarpSmem &S = reinterpret_cast<LkWarpSmem *>(lk_smem)[threadIdx.x >> 5];
    const LkPair &pr = pairs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * LK_WARPS + (threadIdx.x >> 5);
    const int n = min(__ldg(pr.count), pts_cap);
    if (i >= pts_cap) return;
    const size_t o = (size_t)blockIdx.y * pts_cap + i;
    if (i >= n) {
        if (lane == 0) { out_pts[o] = make_float2(0.f, 0.f); out_status[o] = 0; if (out_err) out_err[o] = 0.f; }
        return;
    }
    const int win = WIN ? WIN : p.win, win1 = win + 1, win3 = win + 3;
    const float halfWin = __fmul_rn((float)(win - 1), 0.5f);
    const float FLT_SCALE = 1.f / (1 << 20);
    int maxCount = min(max(p.max_iter, 0), 100);
    double epsilon = fmin(fmax((double)p.eps, 0.), 10.);
    epsilon *= epsilon;
    const float ptx = __ldg(reinterpret_cast<const float *>(pr.pts + (size_t)i * pts_stride));
    const float pty = __ldg(reinterpret_cast<const float *>(pr.pts + (size_t)i * pts_stride) + 1);
    float outx = 0.f, outy = 0.f, errv = 0.f;
    int status = 1;

    for (int level = max_level; level >= 0; level--) {
        const LkLevel &Il = pr.I[level], &Jl = pr.J[level];
        const float sc = (float)(1. / (double)(1 << level));
        float px = __fmul_rn(ptx, sc), py = __fmul_rn(pty, sc), nx, ny;
        if (level == max_level) { nx = px; ny = py; }
        else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;
        px = __fsub_rn(px, halfWin); py = __fsub_rn(py, halfWin);
        const int ipx = lk_floor(px), ipy = lk_floor(py);
        if (ipx < -win || ipx >= Il.w || ipy < -win || ipy >= Il.h) {
            if (level == 0) { status = 0; errv = 0.f; }
            continue;
        }
        // ---- patch of I with its Scharr derivatives
        __syncwarp();
        lk_stage(Il, ipx - 1, ipy - 1, win3, S.nb, lane);
        __syncwarp();
        // Scharr on the (win+1)^2 grid: a lane owns half a row and slides along x, so a position costs 3 new bytes, not 8
        {
            const int halfw = (win1 + 1) >> 1;
#pragma unroll
            for (int item = lane; item < 2 * win1; item += 32) {
                const int yy = item >> 1, xs = (item & 1) ? halfw : 0, xe = (item & 1) ? win1 : halfw;
                const bool yin = (unsigned)(ipy + yy) < (unsigned)Il.h;
                const uint8_t *r = S.nb + yy * win3 + xs;              // rows yy..yy+2, columns xx..xx+2 of the staged bytes
                int a = r[0], b = r[win3], c = r[2 * win3];
                int t0l = (a + c) * 3 + b * 10, t1l = c - a;
                a = r[1]; b = r[win3 + 1]; c = r[2 * win3 + 1];
                int t0m = (a + c) * 3 + b * 10, t1m = c - a;
                short2 *dst = S.d + yy * win1 + xs;
                for (int xx = xs; xx < xe; xx++, r++, dst++) {
                    a = r[2]; b = r[win3 + 2]; c = r[2 * win3 + 2];
                    const int t0r = (a + c) * 3 + b * 10, t1r = c - a;
                    short2 dv = make_short2(0, 0);                         // zero border of the derivative image
                    if (yin && (unsigned)(ipx + xx) < (unsigned)Il.w) dv = make_short2((short)(t0r - t0l), (short)((t1r + t1l) * 3 + t1m * 10));
                    *dst = dv;
                    t0l = t0m; t1l = t1m; t0m = t0r; t1m = t1r;
                }
            }
        }
        __syncwarp();
        LkWeights wt = lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy));
        int a11 = 0, a12 = 0, a22 = 0;
        for (int k = lane; k < win * win; k += 32) {
            const int yy = k / win, xx = k - yy * win;
            const uint8_t *c = S.nb + (yy + 1) * win3 + (xx + 1);
            const short2 *dd = S.d + yy * win1 + xx;
            const int ival = (c[0] * wt.w00 + c[1] * wt.w01 + c[win3] * wt.w10 + c[win3 + 1] * wt.w11 + (1 << 8)) >> 9;
            const short2 d00 = dd[0], d01 = dd[1], d10 = dd[win1], d11 = dd[win1 + 1];
            const int ix = (d00.x * wt.w00 + d01.x * wt.w01 + d10.x * wt.w10 + d11.x * wt.w11 + (1 << 13)) >> 14;
            const int iy = (d00.y * wt.w00 + d01.y * wt.w01 + d10.y * wt.w10 + d11.y * wt.w11 + (1 << 13)) >> 14;
            S.I[k] = (short)ival;
            S.dI[k] = make_int2((short)ix, (short)iy);
            a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
        }
        const float A11 = __fmul_rn(__ll2float_rn(lk_warp_sum(a11)), FLT_SCALE);
        const float A12 = __fmul_rn(__ll2float_rn(lk_warp_sum(a12)), FLT_SCALE);
        const float A22 = __fmul_rn(__ll2float_rn(lk_warp_sum(a22)), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dif = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
                                       (float)(2 * win * win));
        if ((double)minEig < (double)p.min_eig || D < 1.1920929e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, halfWin); ny = __fsub_rn(ny, halfWin);
        float pdx = 0.f, pdy = 0.f;
        int rx = 0, ry = 0;
        bool staged = false;   // S.nb holds the region [rx, rx+32) x [ry, ry+32) of J (the I neighbourhood was there before)
        for (int j = 0; j < maxCount; j++) {
            const int inx = lk_floor(nx), iny = lk_floor(ny);
            if (inx < -win || inx >= Jl.w || iny < -win || iny >= Jl.h) {
                if (level == 0) status = 0;
                break;
            }
            if (!staged || inx < rx || iny < ry || inx + win1 > rx + LK_REG || iny + win1 > ry + LK_REG) {
                rx = inx - (LK_REG - win1) / 2; ry = iny - (LK_REG - win1) / 2;
                __syncwarp();
                lk_stage_region(Jl, rx, ry, S.nb16, lane);
                __syncwarp();
                staged = true;
            }
            const uint16_t *c0 = S.nb16 + (iny - ry) * LK_REG + (inx - rx);
            wt = lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny));
            const int wtop = (wt.w00 & 0xffff) | (wt.w01 << 16), wbot = (wt.w10 & 0xffff) | (wt.w11 << 16);
            int b1i = 0, b2i = 0;
#pragma unroll
            for (int k = lane; k < win * win; k += 32) {
                const int yy = k / win, xx = k - yy * win;
                const uint16_t *c = c0 + yy * LK_REG + xx;
                const int2 g = S.dI[k];
                const int diff = (lk_dp2a(wbot, c[LK_REG], lk_dp2a(wtop, c[0], 1 << 8)) >> 9) - S.I[k];
                b1i += diff * g.x; b2i += diff * g.y;
            }
            const float b1 = __fmul_rn(__ll2float_rn(lk_warp_sum(b1i)), FLT_SCALE), b2 = __fmul_rn(__ll2float_rn(lk_warp_sum(b2i)), FLT_SCALE);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, ddx); ny = __fadd_rn(ny, ddy);
            outx = __fadd_rn(nx, halfWin); outy = __fadd_rn(ny, halfWin);
            if (__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)) <= epsilon) break;
            if (j > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 && fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(ddx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx; pdy = ddy;
        }
        if (status && level == 0) {   // final window check + error measure (the reference passes `err`)
            const float fx = __fsub_rn(outx, halfWin), fy = __fsub_rn(outy, halfWin);
            const int inx = lk_floor(fx), iny = lk_floor(fy);
            if (inx < -win || inx >= Jl.w || iny < -win || iny >= Jl.h) { status = 0; continue; }
            if (!staged || inx < rx || iny < ry || inx + win1 > rx + LK_REG || iny + win1 > ry + LK_REG) {
                rx = inx - (LK_REG - win1) / 2; ry = iny - (LK_REG - win1) / 2;
                __syncwarp();
                lk_stage_region(Jl, rx, ry, S.nb16, lane);
                __syncwarp();
            }
            const uint16_t *c0 = S.nb16 + (iny - ry) * LK_REG + (inx - rx);
            wt = lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny));
            const int wtop = (wt.w00 & 0xffff) | (wt.w01 << 16), wbot = (wt.w10 & 0xffff) | (wt.w11 << 16);
            int e = 0;
#pragma unroll
            for (int k = lane; k < win * win; k += 32) {
                const int yy = k / win, xx = k - yy * win;
                const uint16_t *c = c0 + yy * LK_REG + xx;
                e += abs((lk_dp2a(wbot, c[LK_REG], lk_dp2a(wtop, c[0], 1 << 8)) >> 9) - S.I[k]);
            }
            // sum of |diff| < 2^24: the reference's running float sum is exact, so the integer sum converts to the same float
            errv = __fdiv_rn(__fmul_rn(__ll2float_rn(lk_warp_sum(e)), 1.f), (float)(32 * win * win));
        }
    }
    if (lane == 0) {
        if (check_in_frame && status) {   // matcher.cpp:747-749: IsInFrame(Vector2i(pt.x, pt.y)) — float -> int truncation
            const int xi = __float2int_rz(outx), yi = __float2int_rz(outy);
            if (!(xi >= 0 && xi < frame_w && yi >= 0 && yi < frame_h) || isnan(outx) || isnan(outy)) status = 0;
        }
        out_pts[o] = make_float2(outx, outy);
        out_status[o] = (uint8_t)status;
        if (out_err) out_err[o] = errv;
    }
}

// ------------------------------------------------------------------------------------------------------------ host side
struct FlowWs {
    uint8_t *pyr = nullptr; size_t pyr_bytes = 0;     // LK pyramid levels 1.. of the previous and next image stacks
    LkPair *pairs = nullptr; int pairs_cap = 0;
    uint8_t *eq = nullptr; size_t eq_bytes = 0;       // CLAHE'd copies of level 0 (searchByOPFlow's `equalized`)
};
void tb_flow_ws_free(tb_ctx *ctx)
{
    if (!ctx->fws) return;
    cudaFree(ctx->fws->pyr); cudaFree(ctx->fws->pairs); cudaFree(ctx->fws->eq);
    delete ctx->fws;
    ctx->fws = nullptr;
}

struct LkGeom { int levels; int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS], pitch[LK_MAX_LEVELS]; size_t off[LK_MAX_LEVELS], frame_bytes; };
static void lk_geometry(int w, int h, int win, int max_level, LkGeom &g)
{
    g.levels = 0; g.w[0] = w; g.h[0] = h; g.pitch[0] = 0; g.off[0] = 0; g.frame_bytes = 0;
    for (int l = 1; l <= max_level; l++) {   // buildOpticalFlowPyramid's stop rule
        const int lw = (g.w[l - 1] + 1) / 2, lh = (g.h[l - 1] + 1) / 2;
        if (lw <= win || lh <= win) break;
        if (l >= LK_MAX_LEVELS) { g.levels = -1; return; }   // OpenCV would build more levels than the kernel holds
        g.w[l] = lw; g.h[l] = lh; g.pitch[l] = (int)align_up((size_t)lw, 16);
        g.off[l] = g.frame_bytes;
        g.frame_bytes += align_up((size_t)g.pitch[l] * lh, 256);
        g.levels = l;
    }
}

static int lk_build_pyramid(tb_ctx *ctx, const LkGeom &g, const uint8_t *base, size_t pitch, size_t frame_stride, int n_frames, uint8_t *ws)
{
    for (int l = 1; l <= g.levels; l++) {
        const uint8_t *src = l == 1 ? base : ws + g.off[l - 1];
        const size_t sp = l == 1 ? pitch : (size_t)g.pitch[l - 1], sf = l == 1 ? frame_stride : g.frame_bytes;
        ProfScope prof__(ctx, "lk_pyrdown_u8");
        lk_pyrdown_u8<<<dim3(div_up(g.w[l], 128), div_up(g.h[l], LK_PD_ROWS), n_frames), 128, 0, ctx->stream>>>(src, g.w[l - 1], g.h[l - 1], sp, sf, ws + g.off[l], g.w[l],
                                                                                                   g.h[l], (size_t)g.pitch[l], g.frame_bytes);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

int tb_flow_batched(tb_ctx *ctx, const tb_lk_params *p, const uint8_t *prev, size_t prev_pitch, size_t prev_frame_stride, int n_prev,
                    const uint8_t *next, size_t next_pitch, size_t next_frame_stride, int n_next, int w, int h,
                    int n_pairs, const int32_t *prev_idx, const int32_t *next_idx,
                    const uint8_t *pts, size_t pts_stride, size_t pts_slot_stride, const int32_t *pts_slot, const int *counts, int pts_cap,
                    int check_in_frame, float2 *out_pts, uint8_t *out_status, float *out_err)
{
    if (n_pairs <= 0 || pts_cap <= 0) return TB_OK;
    if (p->win < 3 || p->win > LK_MAX_WIN || p->max_level < 0) return tb_fail(ctx, TB_E_UNSUPPORTED, "LK: window 3..21, max_level >= 0");
    if (w < 2 || h < 2) return tb_fail(ctx, TB_E_UNSUPPORTED, "LK: image must be at least 2x2");
    LkGeom g;
    lk_geometry(w, h, p->win, p->max_level, g);
    if (g.levels < 0) return tb_fail(ctx, TB_E_UNSUPPORTED, "LK: more than 7 pyramid levels above level 0 (image too large for this max_level)");
    if (!ctx->fws) ctx->fws = new FlowWs();
    FlowWs &W = *ctx->fws;
    const bool same = prev == next && prev_pitch == next_pitch && prev_frame_stride == next_frame_stride;
    const int n_first = same ? (n_prev > n_next ? n_prev : n_next) : n_prev;   // one shared pyramid stack covers both index ranges
    const size_t need = g.frame_bytes * ((size_t)n_first + (same ? 0 : (size_t)n_next)) + 256;
    if (need > W.pyr_bytes) {
        cudaFree(W.pyr); W.pyr = nullptr; W.pyr_bytes = 0;
        TB_CUDA(ctx, cudaMalloc(&W.pyr, need));
        W.pyr_bytes = need;
    }
    if (n_pairs > W.pairs_cap) {
        cudaFree(W.pairs); W.pairs = nullptr; W.pairs_cap = 0;
        TB_CUDA(ctx, cudaMalloc(&W.pairs, sizeof(LkPair) * (size_t)n_pairs));
        W.pairs_cap = n_pairs;
    }
    uint8_t *ws_prev = W.pyr, *ws_next = same ? W.pyr : W.pyr + g.frame_bytes * (size_t)n_first;
    int rc = lk_build_pyramid(ctx, g, prev, prev_pitch, prev_frame_stride, n_first, ws_prev);
    if (rc) return rc;
    if (!same) { rc = lk_build_pyramid(ctx, g, next, next_pitch, next_frame_stride, n_next, ws_next); if (rc) return rc; }
    std::vector<LkPair> hp(n_pairs);
    memset(hp.data(), 0, sizeof(LkPair) * n_pairs);
    for (int i = 0; i < n_pairs; i++) {
        const int a = prev_idx[i], b = next_idx[i];
        if (a < 0 || a >= n_prev || b < 0 || b >= n_next) return tb_fail(ctx, TB_E_INVALID, "LK: frame index out of range");
        for (int l = 0; l <= g.levels; l++) {
            LkLevel &I = hp[i].I[l], &J = hp[i].J[l];
            I.w = J.w = g.w[l]; I.h = J.h = g.h[l];
            if (l == 0) { I.p = prev + (size_t)a * prev_frame_stride; I.pitch = (int)prev_pitch; J.p = next + (size_t)b * next_frame_stride; J.pitch = (int)next_pitch; }
            else { I.p = ws_prev + (size_t)a * g.frame_bytes + g.off[l]; J.p = ws_next + (size_t)b * g.frame_bytes + g.off[l]; I.pitch = J.pitch = g.pitch[l]; }
        }
        hp[i].pts = pts + (size_t)pts_slot[i] * pts_slot_stride;
        hp[i].count = counts + pts_slot[i];
    }
    if ((rc = tb_stage_upload(ctx, W.pairs, hp.data(), sizeof(LkPair) * n_pairs))) return rc;   // through the pinned ring: no host wait
    const size_t smem = sizeof(LkWarpSmem) * LK_WARPS;
    {
        ProfScope prof__(ctx, "lk_track");
        const dim3 grid(div_up(pts_cap, LK_WARPS), n_pairs);
        if (p->win == 21) lk_track<21><<<grid, LK_WARPS * 32, smem, ctx->stream>>>(W.pairs, *p, g.levels, pts_stride, pts_cap, w, h, check_in_frame, out_pts, out_status, out_err);
        else lk_track<0><<<grid, LK_WARPS * 32, smem, ctx->stream>>>(W.pairs, *p, g.levels, pts_stride, pts_cap, w, h, check_in_frame, out_pts, out_status, out_err);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

void tb_flow_init_device()   // per device, from tb_ctx_create
{
    cudaFuncSetAttribute(lk_track<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(LkWarpSmem) * LK_WARPS));
    cudaFuncSetAttribute(lk_track<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(LkWarpSmem) * LK_WARPS));
}

extern "C" int tb_lk_flow_device(tb_ctx *ctx, const tb_lk_params *p, const void *prev, size_t prev_pitch, size_t prev_frame_stride, int n_prev,
                                 const void *next, size_t next_pitch, size_t next_frame_stride, int n_next, int w, int h,
                                 int n_pairs, const int32_t *prev_idx, const int32_t *next_idx, const void *pts_f32x2, const void *counts_i32,
                                 int pts_cap, void *next_pts, void *status, void *err)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || !prev || !next || n_pairs < 0 || (n_pairs && (!prev_idx || !next_idx)) || !pts_f32x2 || !counts_i32 || !next_pts || !status ||
        prev_pitch < (size_t)w || next_pitch < (size_t)w || pts_cap < 0)
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_lk_flow_device: bad arguments") : TB_E_INVALID;
    std::vector<int32_t> slot(n_pairs);
    for (int i = 0; i < n_pairs; i++) slot[i] = i;
    return tb_flow_batched(ctx, p, (const uint8_t *)prev, prev_pitch, prev_frame_stride, n_prev, (const uint8_t *)next, next_pitch, next_frame_stride, n_next,
                           w, h, n_pairs, prev_idx, next_idx, (const uint8_t *)pts_f32x2, sizeof(float2), sizeof(float2) * (size_t)pts_cap, slot.data(),
                           (const int *)counts_i32, pts_cap, 0, (float2 *)next_pts, (uint8_t *)status, (float *)err);
}

extern "C" int tb_lk_flow_host(tb_ctx *ctx, const tb_lk_params *p, const uint8_t *prev, const uint8_t *next, int w, int h, size_t stride,
                               const float *prev_xy, int n, float *next_xy, uint8_t *status, float *err)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || !prev || !next || w <= 0 || h <= 0 || stride < (size_t)w || n < 0 || (n && (!prev_xy || !next_xy || !status)))
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_lk_flow_host: bad arguments") : TB_E_INVALID;
    if (n == 0) return TB_OK;
    ArenaScope buf(ctx);
    const size_t pitch = align_up((size_t)w, 16), fs = align_up(pitch * h, 256);
    uint8_t *d_img = buf.alloc<uint8_t>(2 * fs);
    float2 *d_in = buf.alloc<float2>(n), *d_out = buf.alloc<float2>(n);
    uint8_t *d_st = buf.alloc<uint8_t>(n);
    float *d_err = buf.alloc<float>(n);
    int *d_cnt = buf.alloc<int>(1);
    if (!d_img || !d_in || !d_out || !d_st || !d_err || !d_cnt) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    TB_CUDA(ctx, cudaMemcpy2DAsync(d_img, pitch, prev, stride, w, h, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(ctx, cudaMemcpy2DAsync(d_img + fs, pitch, next, stride, w, h, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(ctx, cudaMemcpyAsync(d_in, prev_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(ctx, cudaMemcpyAsync(d_cnt, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const int32_t a = 0, b = 1, slot = 0;
    int rc = tb_flow_batched(ctx, p, d_img, pitch, fs, 2, d_img, pitch, fs, 2, w, h, 1, &a, &b, (const uint8_t *)d_in, sizeof(float2), 0, &slot, d_cnt, n, 0,
                             d_out, d_st, d_err);
    if (rc) return rc;
    TB_CUDA(ctx, cudaMemcpyAsync(next_xy, d_out, sizeof(float2) * n, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaMemcpyAsync(status, d_st, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (err) TB_CUDA(ctx, cudaMemcpyAsync(err, d_err, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

// scratch for the CLAHE'd copies of the frontend's level 0 (searchByOPFlow's `equalized` = F1->Equalize())
int tb_flow_eq_reserve(tb_ctx *ctx, size_t bytes, uint8_t **out)
{
    if (!ctx->fws) ctx->fws = new FlowWs();
    FlowWs &W = *ctx->fws;
    if (bytes > W.eq_bytes) {
        cudaFree(W.eq); W.eq = nullptr; W.eq_bytes = 0;
        TB_CUDA(ctx, cudaMalloc(&W.eq, bytes));
        W.eq_bytes = bytes;
    }
    *out = W.eq;
    return TB_OK;
}

// LocalBA::AddMapPointsByStereo's live body (src/mapping/LocalBA.cpp:46-68) after the tracks: Depth[i] = bf / |pt_i.x - kp_i.x|
// for every valid track, -1 elsewhere (one IEEE division).
__global__ void __launch_bounds__(256)
lk_stereo_depth(const LkPair *__restrict__ pairs, size_t pts_stride, int pts_cap, float bf, const float2 *__restrict__ tracked,
                const uint8_t *__restrict__ status, float *__restrict__ depth)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= pts_cap) return;
    const size_t o = (size_t)blockIdx.y * pts_cap + i;
    float z = -1.0f;
    if (status[o]) {
        const float x0 = __ldg(reinterpret_cast<const float *>(pairs[blockIdx.y].pts + (size_t)i * pts_stride));
        z = __fdiv_rn(bf, fabsf(__fsub_rn(tracked[o].x, x0)));
    }
    depth[o] = z;
}

int tb_flow_stereo_depth(tb_ctx *ctx, int n_pairs, size_t pts_stride, int pts_cap, float bf, const float2 *tracked, const uint8_t *status, float *depth)
{
    if (n_pairs <= 0 || pts_cap <= 0) return TB_OK;
    ProfScope prof__(ctx, "lk_stereo_depth");   // pair descriptors are the ones tb_flow_batched has just uploaded
    lk_stereo_depth<<<dim3(div_up(pts_cap, 256), n_pairs), 256, 0, ctx->stream>>>(ctx->fws->pairs, pts_stride, pts_cap, bf, tracked, status, depth);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}
