// describe.cu — K5 (7x7 Gaussian), K4 (intensity-centroid orientation) and K6 (steered rBRIEF-256),
// plus the final keypoint assembly.  Replaces, in the reference src/extractors/ORBextractor.cpp:
//   GaussianBlur(work, work, Size(7,7), 2, 2, BORDER_REFLECT_101)   :959-960
//   IC_Angle / computeOrientation                                     :17-44, :407-414
//   computeOrbDescriptor / computeDescriptors                         :46-87, :831-838
//   keypoint finishing (border offset, octave, size, pt *= sf)        :812-822, :969-977
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>

// ------------------------------------------------------------------------------------------
// K5: cv::GaussianBlur 7x7 sigma=2 on 8U = OpenCV's fixed-point kernel {18,34,48,56,48,34,18}/256:
// horizontal pass in 8.8 (u16, no rounding), vertical pass in 16.16, one rounding (v+32768)>>16,
// BORDER_REFLECT_101.  No shared memory: a thread owns one 4-pixel column group and walks GB_ROWS
// output rows.  Per input row it loads its word and both neighbours (coalesced, L1 serves the overlap),
// forms the 4 horizontal sums with byte permutes + two DP4A each, and keeps the last 7 rows of sums in
// registers for the vertical pass.  HBM traffic = 1 read + 1 write per pixel (+ 6/GB_ROWS halo rows).
// ------------------------------------------------------------------------------------------
#define GB_ROWS 64

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

struct W3 { uint32_t wm, w0, wp; };

// horizontal 7-tap sums of the 4 pixels of the middle word (u16 range, exact)
__device__ __forceinline__ void gauss_hrow(const W3 a, uint32_t h[4])
{
    // shifted KERNELS instead of shifted data: 10 DP4A and no byte permutes (the permutes sat on the busier integer pipe)
    const uint32_t K0m = (18u << 8) | (34u << 16) | (48u << 24), K0c = 56u | (48u << 8) | (34u << 16) | (18u << 24);
    const uint32_t K1m = (18u << 16) | (34u << 24), K1c = 48u | (56u << 8) | (48u << 16) | (34u << 24), K1p = 18u;
    const uint32_t K2m = (18u << 24), K2c = 34u | (48u << 8) | (56u << 16) | (48u << 24), K2p = 34u | (18u << 8);
    const uint32_t K3c = 18u | (34u << 8) | (48u << 16) | (56u << 24), K3p = 48u | (34u << 8) | (18u << 16);
    h[0] = __dp4a(a.wm, K0m, __dp4a(a.w0, K0c, 0u));
    h[1] = __dp4a(a.wm, K1m, __dp4a(a.w0, K1c, __dp4a(a.wp, K1p, 0u)));
    h[2] = __dp4a(a.wm, K2m, __dp4a(a.w0, K2c, __dp4a(a.wp, K2p, 0u)));
    h[3] = __dp4a(a.w0, K3c, __dp4a(a.wp, K3p, 0u));
}

// vertical 7-tap of the 4 pixels of a word column; window rows oldest..newest = win[(K+1)%7] .. win[K]; returns the 4 output
// bytes ((v + 32768) >> 16 is byte 2 of v + 32768 and never exceeds 255)
template <int K>
__device__ __forceinline__ uint32_t gauss_vcol(const uint32_t (&win)[7][4])
{
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        v[j] = 56u * win[(K + 4) % 7][j] + 32768u;
        v[j] += 48u * (win[(K + 3) % 7][j] + win[(K + 5) % 7][j]);
        v[j] += 34u * (win[(K + 2) % 7][j] + win[(K + 6) % 7][j]);
        v[j] += 18u * (win[(K + 1) % 7][j] + win[K][j]);
    }
    return __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
}

// Border word columns (the first one and the last one or two of a row): each of their three input words is 4 reflected
// pixels that lie inside two adjacent aligned words of the row, so it is one PRMT of two aligned loads; word offsets and
// selectors are computed on the host per level (BORDER_REFLECT_101 in x without byte gathers).
struct GaussEdge {
    int wofs[4][3];       // [border column][wm, w0, wp]: index of the first aligned word
    uint32_t sel[4][3];   // PRMT selector over (word, word + 1)
};

// Interior word columns (their 12-byte neighbourhood lies inside the row): a thread owns one 4-pixel word column and a
// strip of ROWS output rows.
//  * fast path (a full strip with its 3+3 halo rows inside the image — almost every thread): plain 4-byte loads through
//    a row pointer that advances by the pitch, no reflection, no per-row index arithmetic; the 7-row chunks are peeled so
//    that no row needs an "is an output ready" test;
//  * general path (first and last strips): per-row REFLECT_101.
// The first word column and the last ones are done by gauss7x7_border below. (Folding them into this launch, with byte
// gathers or with PRMT selectors, was measured slower: 2.98 / 1.88 ms vs 1.40 ms per step — a few slow threads stretch
// every CTA that holds them.)
template <int ROWS>
__global__ void __launch_bounds__(128, 8)
gauss7x7_u8(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h, int pitch, size_t frame_stride,
            int xw_lo, int xw_hi)   // interior word columns are [xw_lo, xw_hi)
{
    const int xw = xw_lo + blockIdx.x * 128 + threadIdx.x;
    if (xw >= xw_hi) return;
    const int y0 = blockIdx.y * ROWS;
    const int ny = min(ROWS, h - y0);
    const uint8_t *s = src + (size_t)blockIdx.z * frame_stride;
    uint8_t *d = dst + (size_t)blockIdx.z * frame_stride;
    const int x = 4 * xw;
    uint32_t win[7][4];
    if ((ROWS + 6) % 7 == 0 && y0 >= 3 && y0 + ROWS + 3 <= h) {
        const uint8_t *sp = s + (size_t)(y0 - 3) * pitch + x;
        uint8_t *dp = d + (size_t)y0 * pitch + x;
        {   // first chunk: rows y0-3 .. y0+3 fill the window, one output
            W3 in[7];
#pragma unroll
            for (int k = 0; k < 7; k++) {
                const uint32_t *p = reinterpret_cast<const uint32_t *>(sp);
                in[k].wm = __ldg(p - 1); in[k].w0 = __ldg(p); in[k].wp = __ldg(p + 1);
                sp += pitch;
            }
#pragma unroll
            for (int k = 0; k < 7; k++) gauss_hrow(in[k], win[k]);
            *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<6>(win);
            dp += pitch;
        }
        for (int c = 1; c < (ROWS + 6) / 7; c++) {
            W3 in[7];
#pragma unroll
            for (int k = 0; k < 7; k++) {          // 21 independent requests in flight before any arithmetic
                const uint32_t *p = reinterpret_cast<const uint32_t *>(sp);
                in[k].wm = __ldg(p - 1); in[k].w0 = __ldg(p); in[k].wp = __ldg(p + 1);
                sp += pitch;
            }
            gauss_hrow(in[0], win[0]); *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<0>(win); dp += pitch;
            gauss_hrow(in[1], win[1]); *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<1>(win); dp += pitch;
            gauss_hrow(in[2], win[2]); *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<2>(win); dp += pitch;
            gauss_hrow(in[3], win[3]); *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<3>(win); dp += pitch;
            gauss_hrow(in[4], win[4]); *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<4>(win); dp += pitch;
            gauss_hrow(in[5], win[5]); *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<5>(win); dp += pitch;
            gauss_hrow(in[6], win[6]); *reinterpret_cast<uint32_t *>(dp) = gauss_vcol<6>(win); dp += pitch;
        }
        return;
    }
    // general path: input rows y0-3 .. y0+ny+2; output row (r-6) is ready once input row r is in the window.
    // Each chunk first issues the loads of its 7 rows, then computes.
    for (int base = 0; base < ny + 6; base += 7) {
        W3 in[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            int yy = y0 - 3 + min(base + k, ny + 5);
            yy = yy < 0 ? -yy : yy;                      // REFLECT_101 (h >= 4 guaranteed by the launcher)
            yy = yy >= h ? 2 * (h - 1) - yy : yy;
            const uint32_t *p = reinterpret_cast<const uint32_t *>(s + (size_t)yy * pitch + x);
            in[k].wm = __ldg(p - 1); in[k].w0 = __ldg(p); in[k].wp = __ldg(p + 1);
        }
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int r = base + k;
            if (r < ny + 6) {
                gauss_hrow(in[k], win[k]);
                if (r >= 6) {
                    uint32_t out = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t v = 18u * (win[(k + 1) % 7][j] + win[k][j]) + 34u * (win[(k + 2) % 7][j] + win[(k + 6) % 7][j]) +
                                           48u * (win[(k + 3) % 7][j] + win[(k + 5) % 7][j]) + 56u * win[(k + 4) % 7][j];
                        out |= ((v + 32768u) >> 16) << (8 * j);
                    }
                    *reinterpret_cast<uint32_t *>(d + (size_t)(y0 + r - 6) * pitch + x) = out;
                }
            }
        }
    }
}

// Border word columns, one warp per (border column, strip of GB_BROWS output rows): lane = input row (3 halo rows on
// either side), so every lane issues its six word loads at once (one memory round trip per warp), does the horizontal
// pass of its row and fetches the six neighbouring rows' sums with shuffles for the vertical pass. The per-thread
// 64-row walk of the main kernel is latency bound on so few columns (it took a quarter of the Gaussian's time).
#define GB_BROWS 26    // 32 lanes = 26 output rows + 3 + 3 halo rows
__global__ void __launch_bounds__(128)
gauss7x7_border(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h, int pitch, size_t frame_stride,
                int xw_lo, int xw_hi, int nb, int nstrips, const GaussEdge edge)
{
    const int lane = threadIdx.x & 31;
    const int job = blockIdx.x * 4 + (threadIdx.x >> 5);        // (strip, border column)
    if (job >= nstrips * nb) return;
    const int bi = job % nb, strip = job / nb;
    const int xw = bi < xw_lo ? bi : xw_hi + (bi - xw_lo);
    const int x = 4 * xw, y0 = strip * GB_BROWS;
    int yy = y0 - 3 + lane;                                      // BORDER_REFLECT_101 in y (h >= 4 guaranteed by the launcher)
    yy = yy < 0 ? -yy : yy;
    yy = yy >= h ? 2 * (h - 1) - yy : yy;
    yy = min(max(yy, 0), h - 1);                                 // lanes past the last needed halo row: any valid row
    const uint32_t *p = reinterpret_cast<const uint32_t *>(src + (size_t)blockIdx.z * frame_stride + (size_t)yy * pitch);
    W3 in;
    in.wm = __byte_perm(__ldg(p + edge.wofs[bi][0]), __ldg(p + edge.wofs[bi][0] + 1), edge.sel[bi][0]);
    in.w0 = __byte_perm(__ldg(p + edge.wofs[bi][1]), __ldg(p + edge.wofs[bi][1] + 1), edge.sel[bi][1]);
    in.wp = __byte_perm(__ldg(p + edge.wofs[bi][2]), __ldg(p + edge.wofs[bi][2] + 1), edge.sel[bi][2]);
    uint32_t hh[4];
    gauss_hrow(in, hh);
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t a3 = __shfl_sync(0xffffffffu, hh[j], (lane + 29) & 31), b3 = __shfl_sync(0xffffffffu, hh[j], (lane + 3) & 31);
        const uint32_t a2 = __shfl_sync(0xffffffffu, hh[j], (lane + 30) & 31), b2 = __shfl_sync(0xffffffffu, hh[j], (lane + 2) & 31);
        const uint32_t a1 = __shfl_sync(0xffffffffu, hh[j], (lane + 31) & 31), b1 = __shfl_sync(0xffffffffu, hh[j], (lane + 1) & 31);
        const uint32_t v = 18u * (a3 + b3) + 34u * (a2 + b2) + 48u * (a1 + b1) + 56u * hh[j];
        out |= ((v + 32768u) >> 16) << (8 * j);
    }
    const int y = y0 + lane - 3;
    if (lane >= 3 && lane < 3 + GB_BROWS && y < h) {
        uint8_t *d = dst + (size_t)blockIdx.z * frame_stride + (size_t)y * pitch + x;
        if (x + 3 < w) *reinterpret_cast<uint32_t *>(d) = out;
        else for (int j = 0; j < 4 && x + j < w; j++) d[j] = (uint8_t)(out >> (8 * j));
    }
}

// Images with fewer than 4 rows (never a pyramid level in practice): generic per-pixel version.
__global__ void gauss7x7_u8_tiny(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h, int pitch, size_t frame_stride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t *s = src + (size_t)blockIdx.z * frame_stride;
    const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    uint32_t acc = 0;
    for (int j = 0; j < 7; j++) {
        const uint8_t *row = s + (size_t)reflect101(y + j - 3, h) * pitch;
        uint32_t hs = 0;
        for (int i = 0; i < 7; i++) hs += K[i] * row[reflect101(x + i - 3, w)];
        acc += K[j] * hs;
    }
    dst[(size_t)blockIdx.z * frame_stride + (size_t)y * pitch + x] = (uint8_t)((acc + 32768u) >> 16);
}

int tb_launch_gauss(tb_ctx *ctx, cudaStream_t stream, const uint8_t *src, uint8_t *dst, int w, int h, int pitch, size_t frame_stride, int n_frames)
{
    if (h < 4 || w < 16) {
        ProfScope prof__(ctx, "gauss7x7_u8");
        dim3 grid(div_up(w, 128), h, n_frames);
        gauss7x7_u8_tiny<<<grid, 128, 0, stream>>>(src, dst, w, h, pitch, frame_stride);
        TB_LAUNCH_CHECK(ctx);
        return TB_OK;
    }
    // interior word columns: 4*xw >= 4 and 4*xw + 8 <= w; the others (at most 4) get their words through GaussEdge
    const int nwords = (w + 3) >> 2, xw_lo = 1, xw_hi = (w - 8) / 4 + 1;
    GaussEdge edge;
    memset(&edge, 0, sizeof edge);
    const int nb = xw_lo + (nwords - xw_hi);
    if (nb > 4) return tb_fail(ctx, TB_E_UNSUPPORTED, "gauss: unexpected border layout");
    for (int bi = 0; bi < nb; bi++) {
        const int xw = bi < xw_lo ? bi : xw_hi + (bi - xw_lo);
        for (int t = 0; t < 3; t++) {
            int px[4], lo = 1 << 30;
            for (int b = 0; b < 4; b++) {
                int q = 4 * (xw + t - 1) + b;                 // BORDER_REFLECT_101 in x
                while (q < 0 || q >= w) q = q < 0 ? -q : 2 * (w - 1) - q;
                px[b] = q; lo = std::min(lo, q);
            }
            const int wa = lo >> 2;
            uint32_t sel = 0;
            for (int b = 0; b < 4; b++) {
                if (px[b] - 4 * wa > 7) return tb_fail(ctx, TB_E_UNSUPPORTED, "gauss: reflected word spans more than two words");
                sel |= (uint32_t)(px[b] - 4 * wa) << (4 * b);
            }
            edge.wofs[bi][t] = wa; edge.sel[bi][t] = sel;
        }
    }
    {
        ProfScope prof_i__(ctx, "gauss7x7_u8");
        dim3 grid(div_up(xw_hi - xw_lo, 128), div_up(h, GB_ROWS), n_frames);
        gauss7x7_u8<GB_ROWS><<<grid, 128, 0, stream>>>(src, dst, w, h, pitch, frame_stride, xw_lo, xw_hi);
        TB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope prof_b__(ctx, "gauss7x7_border");
        const int nstrips = div_up(h, GB_BROWS);
        gauss7x7_border<<<dim3(div_up(nstrips * nb, 4), 1, n_frames), 128, 0, stream>>>(src, dst, w, h, pitch, frame_stride, xw_lo, xw_hi, nb,
                                                                                          nstrips, edge);
        TB_LAUNCH_CHECK(ctx);
    }
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


// pattern transposed once per device into global memory: g_pat_t[k*32 + lane] = pair 8*lane+k as (x0,y0,x1,y1) int8x4
// (reading c_pattern with a per-lane index serialises in the constant cache, so no kernel does that per CTA)
__device__ uint32_t g_pat_t[256];
__global__ void describe_init_pattern()
{
    const int tid = threadIdx.x, k = tid >> 5, ln = tid & 31;
    const int8_t *q = &c_pattern[4 * (8 * ln + k)];
    g_pat_t[tid] = (uint32_t)(uint8_t)q[0] | ((uint32_t)(uint8_t)q[1] << 8) | ((uint32_t)(uint8_t)q[2] << 16) | ((uint32_t)(uint8_t)q[3] << 24);
}
__global__ void orb_describe_fused(DescArgs a);
void tb_describe_init_device()
{
    describe_init_pattern<<<1, 256>>>();
    cudaFuncSetAttribute(orb_describe_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaDeviceSynchronize();
}

#define DESC_KPB 64   // keypoints per CTA (8 warps x 8 rounds)
__global__ void __launch_bounds__(256, 4)
orb_describe(DescArgs a)
{
    __shared__ uint32_t s_pat[256]; // [k][lane]: pair 8*lane+k as (x0,y0,x1,y1) int8x4, conflict-free
    __shared__ int s_ofs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = a.level_base + blockIdx.y, f = blockIdx.z;
    const int *nk = a.n_kept + (size_t)f * a.n_levels;
    const int nkl = nk[l];
    const bool first = blockIdx.x == 0 && l == a.n_levels - 1;   // the frame's total is written once every level's count exists
    if (blockIdx.x * DESC_KPB >= nkl && !first) return;          // nothing for this CTA (uniform)
    s_pat[tid] = g_pat_t[tid];
    if (warp == 0) {
        // output offset of this level = keypoints of the levels below; the frame's total goes to n_kp
        int c = lane < a.n_levels ? nk[lane] : 0;
        int below = lane < l ? c : 0, total = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { below += __shfl_xor_sync(0xffffffffu, below, o); total += __shfl_xor_sync(0xffffffffu, total, o); }
        if (lane == 0) {
            s_ofs = below;
            if (first) {
                a.n_kp[f] = min(total, a.kp_cap);
                if (total > a.kp_cap) atomicExch(a.overflow, 4);
            }
        }
    }
    __syncthreads();
    const int ofs = s_ofs;
    const DescLevel L = a.lv[l];
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    for (int i = blockIdx.x * DESC_KPB + warp; i < min(nkl, (int)(blockIdx.x + 1) * DESC_KPB); i += 8) {
        const uint32_t c = L.kept[(size_t)f * L.kept_cap + i];
        const int cx = cand_x(c) + 16, cy = cand_y(c) + 16; // + minBorder (:816-817)
        const int out = ofs + i;
        if (out >= a.kp_cap) break;

        // ---- IC_Angle on the un-blurred level: lanes 0..30 = u in [-15,15], loop over rows
        const uint8_t *raw = L.raw + (size_t)f * L.frame_stride + (size_t)cy * L.pitch + cx;
        int m10 = 0, m01 = 0;
        const int u = lane - 15;
        if (lane < 31) {
            // all 31 row loads are independent: issue them together (the kernel is latency bound otherwise)
            const int au = u < 0 ? -u : u;
            int vals[31];
#pragma unroll
            for (int v = -15; v <= 15; v++) vals[v + 15] = (au <= c_umax[v < 0 ? -v : v]) ? (int)raw[v * L.pitch + u] : 0;
            int s0 = 0;
#pragma unroll
            for (int v = -15; v <= 15; v++) { s0 += vals[v + 15]; m01 += v * vals[v + 15]; }
            m10 = u * s0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        const float angle = fast_atan2_deg((float)m01, (float)m10);

        // ---- steered rBRIEF on the blurred level; lane i produces byte i
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
}

// ------------------------------------------------------------------------------------------
// Fused K4+K5+K6 (the default): blur ON DEMAND.  computeOrbDescriptor only ever reads the blurred level inside a 37x37
// window around each keypoint (the pattern's largest rotated coordinate is 18), i.e. <= 2000 x 1369 of the ~2 M pixels of a
// 1080p level, and the dense Gaussian over all levels was 1.2 ms of the 8.2 ms step.  One warp per keypoint:
//   1. stage the 43-row raw neighbourhood (rows cy-21..cy+21) as 12 aligned 32-bit words per row (the span that covers
//      columns cx-21..cx+21) in shared memory; BORDER_REFLECT_101 rows are remapped, keypoints whose window leaves the level
//      on the left/right (the outermost cell columns only) take a byte-wise path;
//   2. IC_Angle from the staged bytes (same integer moments, same fastAtan2);
//   3. horizontal 7-tap sums H (8.8 fixed point, u16, exact) of 40 columns x 43 rows with DP4A on shifted kernels, stored
//      column-major so that a vertical window is contiguous;
//   4. for each of the 512 pattern samples: 7 vertical taps as 4 DP2A on the u16 pairs, (v + 32768) >> 16 — the very
//      integer OpenCV's fixed-point GaussianBlur produces — and the bit comparisons.
// The dense kernels above stay for tb_frontend_download_level(blurred = 1) (the stage-by-stage parity tests).
// ------------------------------------------------------------------------------------------
#define FD_WARPS 8
#define FD_PROW 13                      // words per staged row (12 used; 13 keeps lanes = consecutive rows conflict-free)
#define FD_PWORDS (43 * FD_PROW)
#define FD_HS 44                        // u16 per H column (43 rows + 1 pad); even: a column starts on a word
#define FD_HWORDS (40 * FD_HS / 2)
#define FD_WARP_WORDS (FD_PWORDS + FD_HWORDS + 1)
extern __shared__ __align__(16) uint32_t fd_smem[];

__global__ void __launch_bounds__(FD_WARPS * 32, 4)
orb_describe_fused(DescArgs a)
{
    __shared__ uint32_t s_pat[256]; // [k][lane]: pair 8*lane+k as (x0,y0,x1,y1) int8x4, conflict-free
    __shared__ int s_ofs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = a.level_base + blockIdx.y, f = blockIdx.z;
    const int *nk = a.n_kept + (size_t)f * a.n_levels;
    const int nkl = nk[l];
    const bool first = blockIdx.x == 0 && l == a.n_levels - 1;   // the frame's total is written once every level's count exists
    if (blockIdx.x * DESC_KPB >= nkl && !first) return;          // nothing for this CTA (uniform)
    s_pat[tid] = g_pat_t[tid];
    if (warp == 0) {
        int c = lane < a.n_levels ? nk[lane] : 0;
        int below = lane < l ? c : 0, total = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { below += __shfl_xor_sync(0xffffffffu, below, o); total += __shfl_xor_sync(0xffffffffu, total, o); }
        if (lane == 0) {
            s_ofs = below;
            if (first) {
                a.n_kp[f] = min(total, a.kp_cap);
                if (total > a.kp_cap) atomicExch(a.overflow, 4);
            }
        }
    }
    __syncthreads();
    const int ofs = s_ofs;
    const DescLevel L = a.lv[l];
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    uint32_t *Pw = fd_smem + (size_t)warp * FD_WARP_WORDS;      // staged rows
    uint32_t *Hw = Pw + FD_PWORDS;                              // H, column-major u16 pairs
    const uint8_t *Pb = reinterpret_cast<const uint8_t *>(Pw);
    uint16_t *Hh = reinterpret_cast<uint16_t *>(Hw);
    const uint8_t *img = L.raw + (size_t)f * L.frame_stride;
    const int W = L.w, Hgt = L.h;
    const int st_r = lane >= 12 ? 1 : 0, st_k = lane - 12 * st_r;   // staging role of this lane (lanes 24..31 idle there)
    uint32_t disc = 0;                                          // bit v+15: row v of the radius-15 disc holds column u = lane-15
    {
        const int au = lane < 15 ? 15 - lane : lane - 15;
#pragma unroll
        for (int v = -15; v <= 15; v++) if (au <= c_umax[v < 0 ? -v : v]) disc |= 1u << (v + 15);
    }

    // Staging of keypoint i+8 (this warp's next one) is issued with cp.async right after the H pass of keypoint i — the staged
    // rows are dead from then on — so the global-memory latency hides behind the sampling phase.
    const int i_end = min(nkl, (int)(blockIdx.x + 1) * DESC_KPB);
    auto inside = [&](int cx, int cy) { return cx - 21 >= 0 && cx + 21 < W && cy >= 21 && cy + 21 < Hgt; };
    auto prefetch = [&](uint32_t c) {                           // 43 rows x 12 aligned words -> P, asynchronously
        const int cx = cand_x(c) + 16, cy = cand_y(c) + 16;
        if (!inside(cx, cy)) return;
        const int x0 = cx - 21, sh = x0 & 3;
        const uint8_t *base = img + (size_t)(cy - 21) * L.pitch + (x0 - sh);   // 4-byte aligned (pitch, frame stride: multiples of 64)
        // lanes 0..11 copy the 12 words of an even row, lanes 12..23 those of the odd row below it (22 steps; the address of a
        // step is one multiply-add away from the lane's first one — a flat word index cost 15 instructions per copy)
        if (lane < 24) {
            const uint8_t *src = base + (size_t)st_r * L.pitch + 4 * st_k;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Pw) + 4u * (uint32_t)(st_r * FD_PROW + st_k);
#pragma unroll
            for (int j = 0; j < 22; j++)
                if (2 * j + st_r < 43)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * (uint32_t)(2 * FD_PROW * j)),
                                 "l"(src + (size_t)(2 * j) * L.pitch) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int i = blockIdx.x * DESC_KPB + warp;
    uint32_t c_next = 0;
    if (i < i_end) { c_next = L.kept[(size_t)f * L.kept_cap + i]; prefetch(c_next); }
    for (; i < i_end; i += FD_WARPS) {
        const uint32_t c = c_next;
        const int cx = cand_x(c) + 16, cy = cand_y(c) + 16; // + minBorder (:816-817)
        const int out = ofs + i;
        if (out >= a.kp_cap) break;
        if (i + FD_WARPS < i_end) c_next = L.kept[(size_t)f * L.kept_cap + i + FD_WARPS];

        // ---- 1. rows cy-21 .. cy+21 are staged (byte column of patch column 0 inside a staged row = sh)
        const int x0 = cx - 21;
        int sh;
        if (inside(cx, cy)) {
            sh = x0 & 3;
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else {                                                // window leaves the level (outermost cells only): byte-wise with reflection
            sh = 0;
            uint8_t *Pm = reinterpret_cast<uint8_t *>(Pw);
#pragma unroll 1
            for (int it = lane; it < 43 * 48; it += 32) {
                const int r = it / 48, k = it - r * 48;
                const int y = reflect101(cy - 21 + r, Hgt), x = reflect101(x0 + k, W);
                Pm[r * (FD_PROW * 4) + k] = img[(size_t)y * L.pitch + x];
            }
        }
        __syncwarp();

        // ---- 2. IC_Angle on the staged (un-blurred) bytes: lanes 0..30 = u in [-15,15], rows v in [-15,15]
        int m10 = 0, m01 = 0;
        {
            const int u = lane - 15;
            if (lane < 31) {
                const uint8_t *pc = Pb + 21 * (FD_PROW * 4) + sh + 21 + u;
                int s0 = 0;
#pragma unroll
                for (int v = -15; v <= 15; v++) {
                    const int val = ((disc >> (v + 15)) & 1u) ? (int)pc[v * (FD_PROW * 4)] : 0;
                    s0 += val; m01 += v * val;
                }
                m10 = u * s0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m10 += __shfl_xor_sync(0xffffffffu, m10, o);
                m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            }
        }
        const float angle = fast_atan2_deg((float)m01, (float)m10);

        // ---- 3. horizontal sums: item = (4-column group g, row r); the four sums of patch columns 4g..4g+3 (+0..6)
        {
            const uint32_t K0a = 18u | (34u << 8) | (48u << 16) | (56u << 24), K0b = 48u | (34u << 8) | (18u << 16);
            const uint32_t K1a = (18u << 8) | (34u << 16) | (48u << 24), K1b = 56u | (48u << 8) | (34u << 16) | (18u << 24);
            const uint32_t K2a = (18u << 16) | (34u << 24), K2b = 48u | (56u << 8) | (48u << 16) | (34u << 24), K2c = 18u;
            const uint32_t K3a = (18u << 24), K3b = 34u | (48u << 8) | (56u << 16) | (48u << 24), K3c = 34u | (18u << 8);
            const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)sh;   // PRMT selector: bytes sh .. sh+3 of a word pair
            // a lane owns a row (two passes: rows 0..31, then 32..42) and slides over its 12 words: one load and one byte
            // permute per 4-column group, no address arithmetic (the flat (group, row) index spent a third of this phase on it)
#pragma unroll
            for (int pass = 0; pass < 2; pass++) {
                const int r = lane + 32 * pass;
                if (r < 43) {
                    const uint32_t *row = Pw + r * FD_PROW;
                    uint16_t *hc = Hh + r;
                    uint32_t qa = row[0], qb = row[1], qc = row[2];
                    uint32_t w0 = __byte_perm(qa, qb, sel), w1 = __byte_perm(qb, qc, sel);
#pragma unroll
                    for (int g = 0; g < 10; g++) {
                        const uint32_t qd = g < 9 ? row[g + 3] : 0u;       // group 9 ends at word 11
                        const uint32_t w2 = __byte_perm(qc, qd, sel);
                        hc[(4 * g) * FD_HS] = (uint16_t)__dp4a(w0, K0a, __dp4a(w1, K0b, 0u));
                        hc[(4 * g + 1) * FD_HS] = (uint16_t)__dp4a(w0, K1a, __dp4a(w1, K1b, 0u));
                        hc[(4 * g + 2) * FD_HS] = (uint16_t)__dp4a(w0, K2a, __dp4a(w1, K2b, __dp4a(w2, K2c, 0u)));
                        hc[(4 * g + 3) * FD_HS] = (uint16_t)__dp4a(w0, K3a, __dp4a(w1, K3b, __dp4a(w2, K3c, 0u)));
                        w0 = w1; w1 = w2; qc = qd;
                    }
                }
            }
        }
        __syncwarp();                                           // P is dead, H is complete
        if (i + FD_WARPS < i_end && ofs + i + FD_WARPS < a.kp_cap) prefetch(c_next);

        // ---- 4. steered rBRIEF on the on-demand blur; lane i produces byte i
        const float ang = __fmul_rn(angle, factorPI);
        // glibc cosf/sinf are (nearly always) correctly rounded; evaluate in double and round once
        double dsn, dcs;
        sincos((double)ang, &dsn, &dcs);                        // one argument reduction for both
        const float ca = (float)dcs, sb = (float)dsn;
        auto blurred = [&](int rr, int cc) -> int {             // rr, cc in [-18, 18] relative to the keypoint
            const int j = cc + 18, r0 = rr + 18;                // H column, first of the 7 H rows
            const bool odd = r0 & 1;
            const uint32_t *hw = Hw + j * (FD_HS / 2) + (r0 >> 1);
            const uint32_t A0 = hw[0], A1 = hw[1], A2 = hw[2], A3 = hw[3];
            const uint32_t k0 = odd ? (18u << 8) : (18u | (34u << 8)), k1 = odd ? (34u | (48u << 8)) : (48u | (56u << 8));
            const uint32_t k2 = odd ? (56u | (48u << 8)) : (48u | (34u << 8)), k3 = odd ? (34u | (18u << 8)) : 18u;
            const uint32_t v = __dp2a_lo(A0, k0, __dp2a_lo(A1, k1, __dp2a_lo(A2, k2, __dp2a_lo(A3, k3, 32768u))));
            return (int)(v >> 16);
        };
        int val = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t q = s_pat[k * 32 + lane];
            float px0 = (float)(int8_t)(q & 0xFF), py0 = (float)(int8_t)((q >> 8) & 0xFF);
            float px1 = (float)(int8_t)((q >> 16) & 0xFF), py1 = (float)(int8_t)(q >> 24);
            int r0 = __float2int_rn(__fadd_rn(__fmul_rn(px0, sb), __fmul_rn(py0, ca)));
            int c0 = __float2int_rn(__fsub_rn(__fmul_rn(px0, ca), __fmul_rn(py0, sb)));
            int r1 = __float2int_rn(__fadd_rn(__fmul_rn(px1, sb), __fmul_rn(py1, ca)));
            int c1 = __float2int_rn(__fsub_rn(__fmul_rn(px1, ca), __fmul_rn(py1, sb)));
            const int t0 = blurred(r0, c0), t1 = blurred(r1, c1);
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
    asm volatile("cp.async.wait_all;" ::: "memory");          // nothing in flight when the CTA retires (capacity break above)
}

bool tb_describe_fused()
{
    static const bool v = !(getenv("TB_DESC_FUSED") && atoi(getenv("TB_DESC_FUSED")) == 0);
    return v;
}

int tb_launch_describe(tb_ctx *ctx, const DescArgs &a0, int max_kept_per_level, int n_frames, int level_base, int n_levels_launch, cudaStream_t stream)
{
    DescArgs a = a0;
    a.level_base = level_base;
    dim3 grid(div_up(max_kept_per_level, DESC_KPB), n_levels_launch, n_frames);
    if (tb_describe_fused()) {
        ProfScope prof__(ctx, "orb_describe_fused");
        orb_describe_fused<<<grid, FD_WARPS * 32, sizeof(uint32_t) * FD_WARP_WORDS * FD_WARPS, stream>>>(a);
        TB_LAUNCH_CHECK(ctx);
        return TB_OK;
    }
    ProfScope prof__(ctx, "orb_describe");
    orb_describe<<<grid, 256, 0, stream>>>(a);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}
