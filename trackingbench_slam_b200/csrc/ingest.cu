// ingest.cu — the two image operations in front of the pyramid:
//   PinholeCamera::UndistortImage -> cv::remap(raw, rectified, map1, map2, INTER_LINEAR)   src/camera/CameraModel.cpp:96-102
//   PinholeCamera ctor            -> cv::initUndistortRectifyMap(K, D, I, K, size, CV_32F)  src/camera/CameraModel.cpp:15-18
//   Frame::Equalize               -> cv::createCLAHE(3.0, Size(8,8))->apply                 src/types/Frame.cpp:453-458
// OpenCV's arithmetic is restated bit-exactly (oracle/tb_oracle.c, pinned against cv2 in tests/test_oracle_cv2.py):
//   remap: map*32 rounded half-even to 5 fractional bits, integer parts saturated to int16, bilinear weights
//          (32-fy)(32-fx)*32 out of 2^15, (sum + 2^14) >> 15, BORDER_CONSTANT 0 per tap;
//   CLAHE: per-tile histogram -> clip -> redistribute -> prefix sum * (255/area) -> LUT; per pixel a float bilinear
//          blend of the four surrounding tiles' LUT values, round-half-even.
// Both are byte-in/byte-out, HBM-bound image passes: the float maps are converted once into an 8-byte fixed-point
// record per pixel that stays L2-resident while a batch of frames streams through; each remap thread reads its
// records once and applies them to up to 4 frames.
#include "common.cuh"
#include "kernels.cuh"
#include <climits>
#include <cmath>

struct tb_undistort {
    tb_ctx *ctx;
    int w, h;          // destination size (= map size)
    uint2 *fix;        // [h][w]: .x = (ix & 0xFFFF) | (iy << 16) (int16 each), .y = fx | fy << 5
};

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

#define REMAP_FRAMES 4   // frames that share one read of the map records
__global__ void __launch_bounds__(256)
remap_linear_u8(const uint2 *__restrict__ fix, int dw, int dh, const uint8_t *__restrict__ src, int sw, int sh, size_t spitch,
                size_t sframe, uint8_t *__restrict__ dst, size_t dpitch, size_t dframe, int n_frames)
{
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= dw) return;
    const int f0 = blockIdx.z * REMAP_FRAMES, f1 = min(n_frames, f0 + REMAP_FRAMES);
    const int npx = min(4, dw - x0);
    uint2 rec[4];
    const uint2 *rp = fix + (size_t)y * dw + x0;
#pragma unroll
    for (int k = 0; k < 4; k++) rec[k] = k < npx ? __ldg(rp + k) : make_uint2(0x80008000u, 0);   // (-32768,-32768): outside
    for (int f = f0; f < f1; f++) {
        const uint8_t *s = src + (size_t)f * sframe;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ix = (int)(short)(rec[k].x & 0xFFFFu), iy = (int)(short)(rec[k].x >> 16);
            const int fx = (int)(rec[k].y & 31u), fy = (int)((rec[k].y >> 5) & 31u);
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
            out |= (uint32_t)((sum + 512) >> 10) << (8 * k);
        }
        uint8_t *d = dst + (size_t)f * dframe + (size_t)y * dpitch + x0;
        if (npx == 4 && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) *reinterpret_cast<uint32_t *>(d) = out;
        else for (int k = 0; k < npx; k++) d[k] = (uint8_t)(out >> (8 * k));
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

#define CLAHE_ROWS 4
// one CTA per CLAHE_ROWS image rows of one frame: the (at most 3) tile rows of LUTs those rows blend sit in shared memory
__global__ void __launch_bounds__(256)
clahe_apply(ClaheGeom g, const uint8_t *__restrict__ src, size_t spitch, size_t sframe, const uint8_t *__restrict__ lut,
            uint8_t *__restrict__ dst, size_t dpitch, size_t dframe, int rows_per_cta)
{
    extern __shared__ uint8_t s_lut[];   // [n_tile_rows][tiles_x][256]
    const int tid = threadIdx.x;
    const int y0 = blockIdx.x * rows_per_cta, y1 = min(g.h, y0 + rows_per_cta);
    const int tlo = max(0, (int)floorf(__fsub_rn(__fmul_rn((float)y0, g.inv_th), 0.5f)));
    const int thi = min(g.tiles_y - 1, (int)floorf(__fsub_rn(__fmul_rn((float)(y1 - 1), g.inv_th), 0.5f)) + 1);
    const int row_bytes = g.tiles_x * 256;
    const uint8_t *fl = lut + (size_t)blockIdx.y * g.tiles_y * row_bytes + (size_t)tlo * row_bytes;
    for (int i = tid; i < (thi - tlo + 1) * row_bytes / 16; i += 256) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(fl) + i);
    __syncthreads();
    for (int y = y0; y < y1; y++) {
        const float tyf = __fsub_rn(__fmul_rn((float)y, g.inv_th), 0.5f);
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
        ty1 = max(ty1, 0); ty2 = min(ty2, g.tiles_y - 1);
        const uint8_t *l1 = s_lut + (ty1 - tlo) * row_bytes, *l2 = s_lut + (ty2 - tlo) * row_bytes;
        const uint8_t *srow = src + (size_t)blockIdx.y * sframe + (size_t)y * spitch;
        uint8_t *drow = dst + (size_t)blockIdx.y * dframe + (size_t)y * dpitch;
        for (int x = tid; x < g.w; x += 256) {
            const float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
            tx1 = max(tx1, 0); tx2 = min(tx2, g.tiles_x - 1);
            const int v = srow[x];
            const float a = (float)l1[tx1 * 256 + v], b = (float)l1[tx2 * 256 + v];
            const float c = (float)l2[tx1 * 256 + v], d = (float)l2[tx2 * 256 + v];
            const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(a, xa1), __fmul_rn(b, xa)), ya1),
                                        __fmul_rn(__fadd_rn(__fmul_rn(c, xa1), __fmul_rn(d, xa)), ya));
            drow[x] = (uint8_t)max(0, min(255, __float2int_rn(res)));
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
    if (!ctx || !u || !src || !dst || src_w <= 0 || src_h <= 0 || n_frames < 0 || src_pitch < (size_t)src_w || dst_pitch < (size_t)u->w)
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_undistort_image_device: bad arguments") : TB_E_INVALID;
    if (src_w > 32767 || src_h > 32767) return tb_fail(ctx, TB_E_UNSUPPORTED, "remap: source larger than 32767 (int16 coordinates, as cv::remap)");
    if (src == dst) return tb_fail(ctx, TB_E_INVALID, "remap cannot run in place");
    if (n_frames == 0) return TB_OK;
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
    if (!ctx || !u || !src || !dst || src_w <= 0 || src_h <= 0) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_undistort_image_host: bad arguments") : TB_E_INVALID;
    uint8_t *ds = nullptr, *dd = nullptr;
    if (cudaMalloc(&ds, (size_t)src_w * src_h) != cudaSuccess || cudaMalloc(&dd, (size_t)u->w * u->h) != cudaSuccess) {
        cudaFree(ds); cudaFree(dd);
        return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    }
    int rc = TB_OK;
    cudaError_t e = cudaMemcpy2DAsync(ds, src_w, src, src_stride, src_w, src_h, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = tb_undistort_image_device(ctx, u, ds, src_w, src_h, src_w, 0, dd, u->w, 0, 1);
    if (e == cudaSuccess && rc == TB_OK) e = cudaMemcpy2DAsync(dst, dst_stride, dd, u->w, u->w, u->h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == TB_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(ds); cudaFree(dd);
    if (e != cudaSuccess) return tb_fail(ctx, TB_E_CUDA, "tb_undistort_image_host: %s", cudaGetErrorString(e));
    return rc;
}

struct IngestWs { uint8_t *lut = nullptr; size_t bytes = 0; };
void tb_ingest_ws_free(tb_ctx *ctx)
{
    if (!ctx->iws) return;
    cudaFree(ctx->iws->lut);
    delete ctx->iws;
    ctx->iws = nullptr;
}

extern "C" int tb_clahe_device(tb_ctx *ctx, const void *src, int w, int h, size_t src_pitch, size_t src_frame_stride, void *dst,
                               size_t dst_pitch, size_t dst_frame_stride, int n_frames, double clip_limit, int tiles_x, int tiles_y)
{
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
    if (need > ws.bytes) {
        TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ws.lut); ws.lut = nullptr; ws.bytes = 0;
        TB_CUDA(ctx, cudaMalloc(&ws.lut, need));
        ws.bytes = need;
    }
    {
        ProfScope prof__(ctx, "clahe_tile_lut");
        clahe_tile_lut<<<dim3(tiles_x * tiles_y, n_frames), 256, 0, ctx->stream>>>(g, (const uint8_t *)src, src_pitch, src_frame_stride, ws.lut);
        TB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope prof__(ctx, "clahe_apply");
        const int rows = g.th >= CLAHE_ROWS ? CLAHE_ROWS : 1;     // rows of one CTA touch at most 3 tile rows (2 when rows = 1)
        const size_t smem = (size_t)3 * tiles_x * 256;
        if (smem > 200 * 1024) return tb_fail(ctx, TB_E_UNSUPPORTED, "CLAHE: too many tile columns for the shared-memory LUT window");
        if (smem > 48 * 1024) TB_CUDA(ctx, cudaFuncSetAttribute(clahe_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        clahe_apply<<<dim3(div_up(h, rows), n_frames), 256, smem, ctx->stream>>>(g, (const uint8_t *)src, src_pitch, src_frame_stride, ws.lut,
                                                                                 (uint8_t *)dst, dst_pitch, dst_frame_stride, rows);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

extern "C" int tb_clahe_host(tb_ctx *ctx, const uint8_t *src, int w, int h, size_t src_stride, uint8_t *dst, size_t dst_stride,
                             double clip_limit, int tiles_x, int tiles_y)
{
    if (!ctx || !src || !dst || w <= 0 || h <= 0) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_clahe_host: bad arguments") : TB_E_INVALID;
    uint8_t *ds = nullptr, *dd = nullptr;
    if (cudaMalloc(&ds, (size_t)w * h) != cudaSuccess || cudaMalloc(&dd, (size_t)w * h) != cudaSuccess) {
        cudaFree(ds); cudaFree(dd);
        return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    }
    int rc = TB_OK;
    cudaError_t e = cudaMemcpy2DAsync(ds, w, src, src_stride, w, h, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = tb_clahe_device(ctx, ds, w, h, w, 0, dd, w, 0, 1, clip_limit, tiles_x, tiles_y);
    if (e == cudaSuccess && rc == TB_OK) e = cudaMemcpy2DAsync(dst, dst_stride, dd, w, w, h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == TB_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(ds); cudaFree(dd);
    if (e != cudaSuccess) return tb_fail(ctx, TB_E_CUDA, "tb_clahe_host: %s", cudaGetErrorString(e));
    return rc;
}
