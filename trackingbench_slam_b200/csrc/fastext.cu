// fastext.cu — FASTExtractor (SVO-style grid FAST): FAST-10 + 3x3 non-max + Shi-Tomasi + one-per-cell.
// Replaces FASTExtractor::operator() / AddPoints / shiTomasiScore
// (reference src/extractors/FASTextractor.cpp:8-80, :129-150, :87-127).
//
// The FAST-10 detector / score / non-max suppression live in uzh-rpg/fast (third_part/fast_lib), which is
// NOT in the reference tree: restated from the published algorithm exactly as oracle/tb_oracle.c does
// (10 contiguous ring pixels beyond +-barrier, score = m10-1, a corner is dropped when an 8-neighbour
// corner scores >= it).  PARITY UNPINNED for the detector; the grid / Shi-Tomasi / selection logic is
// in-tree and follows the cited lines (incl. gridPoints.at(k) throwing for k >= nFeatures).
#include "common.cuh"
#include "kernels.cuh"
#include <cmath>

// m10 = max over the 16 arcs of length 10 of min(d) / min(-d); corner at b <=> m10 > b
__device__ int fast10_m(const uint8_t *p, int stride)
{
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[26];
    const int v = p[0];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - p[dy[k] * stride + dx[k]];
#pragma unroll
    for (int k = 16; k < 26; k++) d[k] = d[k - 16];
    int A = -256, B = -256;
#pragma unroll
    for (int s = 0; s < 16; s++) {
        int mn = 256, mx = -256;
#pragma unroll
        for (int k = 0; k < 10; k++) { mn = min(mn, d[s + k]); mx = max(mx, d[s + k]); }
        A = max(A, mn);
        B = max(B, -mx);
    }
    return max(A, B);
}

// score map: m10 (1..255) for corners at `barrier`, 0 elsewhere and in the 3-px border
__global__ void __launch_bounds__(256)
fast10_score_map(const uint8_t *__restrict__ img, int w, int h, int pitch, int barrier, uint8_t *__restrict__ smap)
{
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    int m = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        const uint8_t *p = img + (size_t)y * pitch + x;
        const int v = p[0], lo = v - barrier, hi = v + barrier;
        const int a0 = p[3 * pitch], a8 = p[-3 * pitch];
        if (a0 < lo || a0 > hi || a8 < lo || a8 > hi) {      // every 10-arc holds one pixel of each antipodal pair
            const int a4 = p[3], a12 = p[-3];
            if (a4 < lo || a4 > hi || a12 < lo || a12 > hi) {
                m = fast10_m(p, pitch);
                if (m <= barrier) m = 0;
            }
        }
    }
    smap[(size_t)y * pitch + x] = (uint8_t)m;
}

// FASTExtractor::shiTomasiScore (FASTextractor.cpp:87-127): sequential float accumulation, no FMA
__device__ float shi_tomasi(const uint8_t *img, int w, int h, int pitch, int u, int v)
{
    float dXX = 0.f, dYY = 0.f, dXY = 0.f;
    const int x_min = u - 4, x_max = u + 4, y_min = v - 4, y_max = v + 4;
    if (x_min < 1 || x_max >= w - 1 || y_min < 1 || y_max >= h - 1) return 0.f;
    for (int y = y_min; y < y_max; ++y) {
        const uint8_t *row = img + (size_t)y * pitch + x_min;
        for (int x = 0; x < 8; ++x) {
            const float dx = __fsub_rn((float)row[x + 1], (float)row[x - 1]);
            const float dy = __fsub_rn((float)row[x + pitch], (float)row[x - pitch]);
            dXX = __fadd_rn(dXX, __fmul_rn(dx, dx));
            dYY = __fadd_rn(dYY, __fmul_rn(dy, dy));
            dXY = __fadd_rn(dXY, __fmul_rn(dx, dy));
        }
    }
    const float box2 = __fmul_rn(2.f, 64.f);
    dXX = __fdiv_rn(dXX, box2); dYY = __fdiv_rn(dYY, box2); dXY = __fdiv_rn(dXY, box2);
    const float tr = __fadd_rn(dXX, dYY);
    const float det4 = __fmul_rn(4.f, __fsub_rn(__fmul_rn(dXX, dYY), __fmul_rn(dXY, dXY)));
    return __fmul_rn(0.5f, __fsub_rn(tr, __fsqrt_rn(__fsub_rn(__fmul_rn(tr, tr), det4))));
}

// non-max (a neighbour corner with score >= own suppresses), cell lookup, Shi-Tomasi, best per cell.
// The reference visits levels in order and corners row-major and replaces on strict '>' : the winner of a
// cell is the max score, earliest (level, y, x) on ties  ==  atomicMax of (score bits, ~order).
__global__ void __launch_bounds__(256)
fast10_nms_grid(const uint8_t *__restrict__ img, const uint8_t *__restrict__ smap, int w, int h, int pitch, int level,
                float inv_sf, int cell_size, int grid_n_cols, int n_cells, int n_features, const uint8_t *__restrict__ occupied,
                unsigned long long *__restrict__ cell_key, int *__restrict__ err)
{
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x < 3 || x >= w - 3 || y < 3 || y >= h - 3) return;
    const int m = smap[(size_t)y * pitch + x];
    if (!m) return;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            const int m2 = smap[(size_t)(y + dy) * pitch + x + dx];
            if (m2 && m2 >= m) return;
        }
    const int k = (int)__fdiv_rn(__fmul_rn((float)y, inv_sf), (float)cell_size) * grid_n_cols +
                  (int)__fdiv_rn(__fmul_rn((float)x, inv_sf), (float)cell_size);
    if (occupied && k >= 0 && k < n_cells && occupied[k]) return;
    const float score = shi_tomasi(img, w, h, pitch, x, y);
    if (k < 0 || k >= n_features) { atomicExch(err, 1); return; }   // gridPoints.at(k) throws in the reference
    if (score > 0.f) {
        const unsigned long long order = ((unsigned long long)level << 24) | ((unsigned long long)y << 12) | (unsigned long long)x;
        atomicMax(&cell_key[k], ((unsigned long long)__float_as_uint(score) << 32) | (0xFFFFFFFFull - order));
    }
}

struct InvSf { float v[TB_MAX_LEVELS]; };
__global__ void __launch_bounds__(256)
fast_grid_emit(const unsigned long long *__restrict__ cell_key, int n_features, float threshold, InvSf inv, tb_keypoint *out, int *out_n)
{
    __shared__ int s_run, s_warp[8];
    const int tid = threadIdx.x;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < n_features; base += 256) {
        const int k = base + tid;
        bool emit = false;
        tb_keypoint kp;
        if (k < n_features) {
            const unsigned long long key = cell_key[k];
            const float score = __uint_as_float((unsigned)(key >> 32));
            if (key && score > threshold) {
                const unsigned order = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
                const int L = order >> 24, y = (order >> 12) & 0xFFF, x = order & 0xFFF;
                kp.x = __fmul_rn((float)x, inv.v[L]); kp.y = __fmul_rn((float)y, inv.v[L]);
                kp.size = 1.f; kp.angle = 0.f; kp.response = score; kp.octave = L; kp.class_id = -1;
                emit = true;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, emit);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        int off = s_run;
        for (int wq = 0; wq < (tid >> 5); wq++) off += s_warp[wq];
        if (emit) out[off + __popc(bal & ((1u << (tid & 31)) - 1))] = kp;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int wq = 0; wq < 8; wq++) t += s_warp[wq]; s_run += t; }
        __syncthreads();
    }
    if (tid == 0) *out_n = s_run;
}

extern "C" int tb_fast_extract_host(tb_ctx *ctx, const uint8_t *const *levels, const int32_t *w, const int32_t *h, const size_t *strides,
                                    int n_levels, const float *inv_sf, int n_features, float threshold,
                                    const tb_keypoint *exit_kps, int n_exit, tb_keypoint *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !levels || !w || !h || !strides || !inv_sf || !n || n_levels <= 0 || n_levels > TB_MAX_LEVELS || n_features <= 0 || n_exit < 0)
        return TB_E_INVALID;
    *n = 0;
    if (!levels[0] || w[0] <= 0 || h[0] <= 0) return TB_OK;   // silent return on empty input (FASTextractor.cpp:16-17)
    for (int l = 0; l < n_levels; l++) if (w[l] > 4095 || h[l] > 4095) return tb_fail(ctx, TB_E_UNSUPPORTED, "level dimensions above 4095 are not supported");
    // grid geometry (FASTextractor.cpp:18-20)
    volatile float area_per = (float)w[0] * (float)h[0] / (float)n_features;
    const int cell_size = (int)sqrtf(area_per);
    if (cell_size <= 0) return tb_fail(ctx, TB_E_INVALID, "nFeatures too large for the image (cell_size 0)");
    const int grid_n_cols = (int)((float)w[0] / (float)cell_size), grid_n_rows = (int)((float)h[0] / (float)cell_size);
    const int n_cells = grid_n_cols * grid_n_rows;
    std::vector<uint8_t> occ;
    if (exit_kps) {   // AddPoints (:129-150): grid_occupancy_.at(...) = true
        occ.assign(n_cells > 0 ? n_cells : 1, 0);
        for (int i = 0; i < n_exit; i++) {
            const int k = (int)(exit_kps[i].y / (float)cell_size) * grid_n_cols + (int)(exit_kps[i].x / (float)cell_size);
            if (k < 0 || k >= n_cells) return tb_fail(ctx, TB_E_INVALID, "AddPoints: exit point outside the occupancy grid (the reference throws std::out_of_range)");
            occ[k] = 1;
        }
    }
    ArenaScope bufs(ctx);   // per-ctx scratch, rewound when the call returns (no cudaMalloc/cudaFree in steady state)
    auto dalloc = [&](size_t bytes) -> void * { return bufs.alloc<uint8_t>(bytes + 256); };
    unsigned long long *cell_key = (unsigned long long *)dalloc(sizeof(unsigned long long) * n_features);
    int *flags = (int *)dalloc(sizeof(int) * 2);
    tb_keypoint *d_out = (tb_keypoint *)dalloc(sizeof(tb_keypoint) * n_features);
    uint8_t *d_occ = exit_kps ? (uint8_t *)dalloc(occ.size()) : nullptr;
    if (!cell_key || !flags || !d_out || (exit_kps && !d_occ)) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    TB_CUDA(ctx, cudaMemsetAsync(cell_key, 0, sizeof(unsigned long long) * n_features, ctx->stream));
    TB_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int) * 2, ctx->stream));
    if (d_occ) TB_CUDA(ctx, cudaMemcpyAsync(d_occ, occ.data(), occ.size(), cudaMemcpyHostToDevice, ctx->stream));
    InvSf inv; memset(&inv, 0, sizeof inv);
    for (int l = 0; l < n_levels; l++) {
        inv.v[l] = inv_sf[l];
        const int pitch = (int)align_up((size_t)w[l], 64);
        uint8_t *d_img = (uint8_t *)dalloc((size_t)pitch * h[l]), *d_map = (uint8_t *)dalloc((size_t)pitch * h[l]);
        if (!d_img || !d_map) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
        TB_CUDA(ctx, cudaMemcpy2DAsync(d_img, pitch, levels[l], strides[l], w[l], h[l], cudaMemcpyHostToDevice, ctx->stream));
        dim3 grid(div_up(w[l], 64), div_up(h[l], 4));
        fast10_score_map<<<grid, 256, 0, ctx->stream>>>(d_img, w[l], h[l], pitch, 20, d_map);   // barrier 20 is hard-coded (:38)
        TB_LAUNCH_CHECK(ctx);
        fast10_nms_grid<<<grid, 256, 0, ctx->stream>>>(d_img, d_map, w[l], h[l], pitch, l, inv_sf[l], cell_size, grid_n_cols, n_cells,
                                                       n_features, d_occ, cell_key, flags);
        TB_LAUNCH_CHECK(ctx);
    }
    fast_grid_emit<<<1, 256, 0, ctx->stream>>>(cell_key, n_features, threshold, inv, d_out, flags + 1);
    TB_LAUNCH_CHECK(ctx);
    int hf[2] = {0, 0};
    TB_CUDA(ctx, cudaMemcpyAsync(hf, flags, sizeof hf, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hf[0]) return tb_fail(ctx, TB_E_INVALID, "a corner maps to grid cell >= nFeatures: the reference throws std::out_of_range from gridPoints.at(k)");
    *n = hf[1];
    const int m = hf[1] < capacity ? hf[1] : capacity;
    if (m > 0) TB_CUDA(ctx, cudaMemcpy(out, d_out, sizeof(tb_keypoint) * m, cudaMemcpyDeviceToHost));
    return hf[1] > capacity ? tb_fail(ctx, TB_E_CAPACITY, "keypoint output capacity too small") : TB_OK;
}
