// describe.cu — K5 (7x7 Gaussian), K4 (intensity-centroid orientation) and K6 (steered rBRIEF-256),
// plus the final keypoint assembly.  Replaces, in the reference src/extractors/ORBextractor.cpp:
//   GaussianBlur(work, work, Size(7,7), 2, 2, BORDER_REFLECT_101)   :959-960
//   IC_Angle / computeOrientation                                     :17-44, :407-414
//   computeOrbDescriptor / computeDescriptors                         :46-87, :831-838
//   keypoint finishing (border offset, octave, size, pt *= sf)        :812-822, :969-977
#include "common.cuh"
#include "kernels.cuh"

// ------------------------------------------------------------------------------------------
// K5: cv::GaussianBlur 7x7 sigma=2 on 8U = OpenCV's fixed-point kernel {18,34,48,56,48,34,18}/256:
// horizontal pass in 8.8 (u16, no rounding), vertical pass in 16.16, one rounding (v+32768)>>16,
// BORDER_REFLECT_101.  Tile = 128x32 outputs per CTA; the (128+8)x(32+6) u8 source box is staged
// in shared memory with 4-byte loads, the horizontal pass writes 4xu16 (8-byte) groups, the
// vertical pass reads 7 such groups per 4 outputs and stores 4 bytes.  HBM traffic = 1 read +
// 1 write per pixel (+ halo re-reads served by L2).
// ------------------------------------------------------------------------------------------
#define GB_TW 128
#define GB_TH 32
#define GB_INW (GB_TW + 8)   // bytes per staged row: [x0-4, x0+132)
#define GB_INH (GB_TH + 6)

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256)
gauss7x7_u8(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h, int pitch, size_t frame_stride)
{
    __shared__ __align__(16) uint8_t in[GB_INH][GB_INW];
    __shared__ __align__(16) uint16_t hs[GB_INH][GB_TW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * GB_TW, y0 = blockIdx.y * GB_TH;
    const uint8_t *s = src + (size_t)blockIdx.z * frame_stride;
    uint8_t *d = dst + (size_t)blockIdx.z * frame_stride;

    // stage: word ww of row rr covers x in [x0-4+4ww, x0+4ww)
    for (int i = tid; i < GB_INH * (GB_INW / 4); i += 256) {
        int rr = i / (GB_INW / 4), ww = i - rr * (GB_INW / 4);
        int sy = reflect101(y0 - 3 + rr, h);
        int xs = x0 - 4 + 4 * ww;
        uint32_t v;
        if (xs >= 0 && xs + 3 < w) {
            v = __ldg(reinterpret_cast<const uint32_t *>(s + (size_t)sy * pitch + xs));
        } else {
            v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) v |= (uint32_t)s[(size_t)sy * pitch + reflect101(xs + k, w)] << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(&in[rr][4 * ww]) = v;
    }
    __syncthreads();
    // horizontal: item = (row rr, group g of 4 outputs); output x0+4g+k needs bytes 4g+1+k .. 4g+7+k
    for (int i = tid; i < GB_INH * (GB_TW / 4); i += 256) {
        int rr = i / (GB_TW / 4), g = i - rr * (GB_TW / 4);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(&in[rr][4 * g]);
        uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
        int b[12];
#pragma unroll
        for (int k = 0; k < 4; k++) { b[k] = (w0 >> (8 * k)) & 0xFF; b[4 + k] = (w1 >> (8 * k)) & 0xFF; b[8 + k] = (w2 >> (8 * k)) & 0xFF; }
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            o[k] = 18 * (b[k + 1] + b[k + 7]) + 34 * (b[k + 2] + b[k + 6]) + 48 * (b[k + 3] + b[k + 5]) + 56 * b[k + 4];
        uint2 pk = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
        *reinterpret_cast<uint2 *>(&hs[rr][4 * g]) = pk;
    }
    __syncthreads();
    // vertical
    for (int i = tid; i < GB_TH * (GB_TW / 4); i += 256) {
        int yy = i / (GB_TW / 4), g = i - yy * (GB_TW / 4);
        int x = x0 + 4 * g, y = y0 + yy;
        if (x >= w || y >= h) continue;
        uint32_t acc[4] = {0, 0, 0, 0};
        const int K[7] = {18, 34, 48, 56, 48, 34, 18};
#pragma unroll
        for (int j = 0; j < 7; j++) {
            uint2 v = *reinterpret_cast<const uint2 *>(&hs[yy + j][4 * g]);
            acc[0] += K[j] * (v.x & 0xFFFF); acc[1] += K[j] * (v.x >> 16);
            acc[2] += K[j] * (v.y & 0xFFFF); acc[3] += K[j] * (v.y >> 16);
        }
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) out |= ((acc[k] + 32768u) >> 16) << (8 * k);
        if (x + 3 < w) *reinterpret_cast<uint32_t *>(d + (size_t)y * pitch + x) = out;
        else for (int k = 0; k < 4 && x + k < w; k++) d[(size_t)y * pitch + x + k] = (uint8_t)(out >> (8 * k));
    }
}

int tb_launch_gauss(tb_ctx *ctx, const uint8_t *src, uint8_t *dst, int w, int h, int pitch, size_t frame_stride, int n_frames)
{
    ProfScope prof__(ctx, "gauss7x7_u8");
    dim3 grid(div_up(w, GB_TW), div_up(h, GB_TH), n_frames);
    gauss7x7_u8<<<grid, 256, 0, ctx->stream>>>(src, dst, w, h, pitch, frame_stride);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

// ------------------------------------------------------------------------------------------
// K4 + K6 + assembly: one warp per kept keypoint.
// ------------------------------------------------------------------------------------------
__device__ __constant__ int8_t c_pattern[256 * 4] = {
#include "orb_pattern.inc"
};
// umax of the radius-15 disc (ORBextractor.cpp:389-404): row |v| spans u in [-umax, umax]
__device__ __constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2 (OpenCV core/mathfuncs_core scalar path): float polynomial, no FMA.
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = (float)2.2204460492503131e-16;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}


__global__ void __launch_bounds__(256)
orb_describe(DescArgs a)
{
    __shared__ uint32_t s_pat[256]; // [k][lane]: pair 8*lane+k as (x0,y0,x1,y1) int8x4, conflict-free
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = blockIdx.y, f = blockIdx.z;
    {
        int k = tid >> 5, ln = tid & 31; // tid = k*32 + ln
        const int8_t *q = &c_pattern[4 * (8 * ln + k)];
        s_pat[tid] = (uint32_t)(uint8_t)q[0] | ((uint32_t)(uint8_t)q[1] << 8) | ((uint32_t)(uint8_t)q[2] << 16) | ((uint32_t)(uint8_t)q[3] << 24);
    }
    __syncthreads();
    const int *nk = a.n_kept + (size_t)f * a.n_levels;
    int ofs = 0, total = 0;
    for (int i = 0; i < a.n_levels; i++) { int c = nk[i]; if (i < l) ofs += c; total += c; }
    if (blockIdx.x == 0 && l == 0 && tid == 0) {
        a.n_kp[f] = min(total, a.kp_cap);
        if (total > a.kp_cap) atomicExch(a.overflow, 4);
    }
    const int i = blockIdx.x * 8 + warp;
    if (i >= nk[l]) return;
    const DescLevel L = a.lv[l];
    const uint32_t c = L.kept[(size_t)f * L.kept_cap + i];
    const int cx = cand_x(c) + 16, cy = cand_y(c) + 16; // + minBorder (:816-817)
    const int out = ofs + i;
    if (out >= a.kp_cap) return;

    // ---- IC_Angle on the un-blurred level: lanes 0..30 = u in [-15,15], loop over rows
    const uint8_t *raw = L.raw + (size_t)f * L.frame_stride + (size_t)cy * L.pitch + cx;
    int m10 = 0, m01 = 0;
    const int u = lane - 15;
    if (lane < 31) {
#pragma unroll 1
        for (int v = -15; v <= 15; v++) {
            int um = c_umax[v < 0 ? -v : v];
            if (u >= -um && u <= um) {
                int val = raw[v * L.pitch + u];
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered rBRIEF on the blurred level; lane i produces byte i
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    // glibc cosf/sinf are (nearly always) correctly rounded; evaluate in double and round once
    const float ca = (float)cos((double)ang), sb = (float)sin((double)ang);
    const uint8_t *blur = L.blur + (size_t)f * L.frame_stride + (size_t)cy * L.pitch + cx;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t q = s_pat[k * 32 + lane];
        float x0 = (float)(int8_t)(q & 0xFF), y0 = (float)(int8_t)((q >> 8) & 0xFF);
        float x1 = (float)(int8_t)((q >> 16) & 0xFF), y1 = (float)(int8_t)(q >> 24);
        int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sb), __fmul_rn(y0, ca)));
        int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sb)));
        int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sb), __fmul_rn(y1, ca)));
        int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sb)));
        int t0 = blur[r0 * L.pitch + c0], t1 = blur[r1 * L.pitch + c1];
        val |= (t0 < t1) << k;
    }
    a.desc[((size_t)f * a.kp_cap + out) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        tb_keypoint kp;
        kp.x = (float)cx; kp.y = (float)cy;
        if (l != 0) { kp.x = __fmul_rn(kp.x, L.sf); kp.y = __fmul_rn(kp.y, L.sf); } // :969-974 (multiplies; replicated)
        kp.size = L.size;
        kp.angle = angle;
        kp.response = (float)(cand_m(c) - 1);
        kp.octave = l;
        kp.class_id = -1;
        a.kps[(size_t)f * a.kp_cap + out] = kp;
    }
}

int tb_launch_describe(tb_ctx *ctx, const DescArgs &a, int max_kept_per_level, int n_frames)
{
    ProfScope prof__(ctx, "orb_describe");
    dim3 grid(div_up(max_kept_per_level, 8), a.n_levels, n_frames);
    orb_describe<<<grid, 256, 0, ctx->stream>>>(a);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}
