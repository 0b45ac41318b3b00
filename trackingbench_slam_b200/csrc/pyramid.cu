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
#include "kernels.cuh"
#include <cfloat>
#include <cmath>


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
__global__ void __launch_bounds__(256)
pyr_resize_linear_u8(const uint8_t *__restrict__ src, int spitch, size_t sframe,
                     uint8_t *__restrict__ dst, int dpitch, size_t dframe,
                     int dw, int dh, const int4 *__restrict__ xtab, const int4 *__restrict__ ytab)
{
    const int dx0 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int dy = blockIdx.y * 4 + threadIdx.y;
    if (dx0 >= dw || dy >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sframe;
    uint8_t *d = dst + (size_t)blockIdx.z * dframe;
    const int4 yt = __ldg(&ytab[dy]);
    const uint8_t *r0 = s + (size_t)yt.x * spitch;
    const uint8_t *r1 = s + (size_t)yt.y * spitch;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int dx = dx0 + k;
        if (dx < dw) {
            const int4 xt = __ldg(&xtab[dx]);
            int h0 = (int)r0[xt.x] * xt.z + (int)r0[xt.y] * xt.w;
            int h1 = (int)r1[xt.x] * xt.z + (int)r1[xt.y] * xt.w;
            int v = (((yt.z * (h0 >> 4)) >> 16) + ((yt.w * (h1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xFF) << (8 * k);
        }
    }
    if (dx0 + 3 < dw) {
        *reinterpret_cast<uint32_t *>(d + (size_t)dy * dpitch + dx0) = out;
    } else {
        for (int k = 0; k < 4 && dx0 + k < dw; k++) d[(size_t)dy * dpitch + dx0 + k] = (uint8_t)(out >> (8 * k));
    }
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
                     size_t dframe, int dw, int dh, const int4 *xtab, const int4 *ytab, bool area2, int n_frames)
{
    ProfScope prof__(ctx, "pyr_resize");
    if (area2) {
        dim3 grid(div_up(dw, 64), div_up(dh, 4), n_frames), block(64, 4);
        pyr_resize_area2_u8<<<grid, block, 0, ctx->stream>>>(src, spitch, sframe, dst, dpitch, dframe, dw, dh);
    } else {
        dim3 grid(div_up(dw, 256), div_up(dh, 4), n_frames), block(64, 4);
        pyr_resize_linear_u8<<<grid, block, 0, ctx->stream>>>(src, spitch, sframe, dst, dpitch, dframe, dw, dh, xtab, ytab);
    }
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}
