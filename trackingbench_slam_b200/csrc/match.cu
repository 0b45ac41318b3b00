// match.cu — K7: brute-force Hamming matching with LOP3 + POPC.
// Replaces, in the reference src/matchers/matcher.cpp:
//   Matcher::DescriptorDistance                                   :793-808
//   cv::BFMatcher(NORM_HAMMING, crossCheck=true).match             :20, :207
//   Matcher::searchByBF (all-levels branch) filter                 :209-218
// and provides the best / second-best / argbest kernel shared by the windowed matchers.
//
// Layout: descriptors are rows of 32 bytes (8 x u32).  A CTA owns 128 query rows (one per thread,
// 8 registers) and streams the train set through shared memory in tiles of 128 rows staged with
// 16-byte loads; every thread reads the same train row at the same time (shared-memory
// broadcast), so the inner loop is 8 LOP3(xor) + 8 POPC + adds per pair and nothing else.
// Ties are broken by lowest train index (strict < while scanning in index order).
#include "common.cuh"
#include "kernels.cuh"

#define KNN_THREADS 128
#define KNN_TILE 128

__device__ __forceinline__ int ham256(const uint32_t q[8], const uint4 a, const uint4 b)
{
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
           __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

// pairs: device array; blockIdx.y = pair. Outputs indexed [pair*out_stride + i].
// best_idx = -1 / best = second = INT_MAX-like (0x7fffffff) when the train set is empty.
template <bool SECOND>
__global__ void __launch_bounds__(KNN_THREADS)
hamming_knn(const PairDesc *__restrict__ pairs, int out_stride, int *__restrict__ best_idx,
            int *__restrict__ best_dist, int *__restrict__ second_dist)
{
    __shared__ uint4 tile[KNN_TILE * 2];
    const PairDesc pd = pairs[blockIdx.y];
    const int q0 = blockIdx.x * KNN_THREADS;
    if (q0 >= pd.nq) return;
    const int i = q0 + threadIdx.x;
    const bool active = i < pd.nq;
    uint32_t q[8];
    {
        const uint4 *qp = reinterpret_cast<const uint4 *>(pd.q) + 2 * (size_t)(active ? i : q0);
        uint4 a = __ldg(qp), b = __ldg(qp + 1);
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    int best = 0x7fffffff, second = 0x7fffffff, bidx = -1;
    const uint4 *tp = reinterpret_cast<const uint4 *>(pd.t);
    for (int t0 = 0; t0 < pd.nt; t0 += KNN_TILE) {
        const int nt = min(KNN_TILE, pd.nt - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < 2 * nt; k += KNN_THREADS) tile[k] = __ldg(tp + 2 * (size_t)t0 + k);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nt; j++) {
            int d = ham256(q, tile[2 * j], tile[2 * j + 1]);
            if (SECOND) {
                if (d < best) { second = best; best = d; bidx = t0 + j; }
                else if (d < second) second = d;
            } else {
                if (d < best) { best = d; bidx = t0 + j; }
            }
        }
    }
    if (active) {
        size_t o = (size_t)blockIdx.y * out_stride + i;
        best_idx[o] = bidx;
        best_dist[o] = best;
        if (SECOND) second_dist[o] = second;
    }
}

// Pair table for a frontend batch: pairs[p] = (q_frames[p] -> t_frames[p]), pairs[P+p] = reverse.
__global__ void build_pairs(PairDesc *pairs, int P, const int *q_frames, const int *t_frames,
                            const uint8_t *desc, size_t frame_stride, const int *counts)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int qf = q_frames[p], tf = t_frames[p];
    PairDesc a = {desc + qf * frame_stride, desc + tf * frame_stride, counts[qf], counts[tf]};
    PairDesc b = {a.t, a.q, a.nt, a.nq};
    pairs[p] = a;
    pairs[P + p] = b;
}

// In-place exclusive block scan helper for 256 threads over a[0..n)
__device__ int scan256(int *a, int n)
{
    __shared__ int s_warp[8];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int ipt = (n + 255) / 256;
    const int b0 = min(tid * ipt, n), b1 = min(b0 + ipt, n);
    int sum = 0;
    for (int i = b0; i < b1; i++) sum += a[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += t; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    if (tid == 0) { int r = 0; for (int w = 0; w < 8; w++) { int v = s_warp[w]; s_warp[w] = r; r += v; } s_total = r; }
    __syncthreads();
    int run = s_warp[tid >> 5] + inc - sum;
    for (int i = b0; i < b1; i++) { int v = a[i]; a[i] = run; run += v; }
    __syncthreads();
    return s_total;
}

// Cross-check (nn21[nn12[i]] == i), optional searchByBF filter, ordered compaction (ascending i).
// One CTA per pair. scratch: int[P][out_stride] flags.
__global__ void __launch_bounds__(256)
bf_cross_filter(const PairDesc *__restrict__ pairs, int P, int out_stride, const int *__restrict__ best_idx,
                const int *__restrict__ best_dist, int *__restrict__ scratch, int apply_filter, float ratio, float minTh,
                tb_match *__restrict__ out, int capacity, int *__restrict__ out_count)
{
    const int p = blockIdx.x, tid = threadIdx.x;
    const int n1 = pairs[p].nq, n2 = pairs[p].nt;
    const int *nn12 = best_idx + (size_t)p * out_stride, *d12 = best_dist + (size_t)p * out_stride;
    const int *nn21 = best_idx + (size_t)(P + p) * out_stride;
    int *flag = scratch + (size_t)p * out_stride;
    __shared__ int s_min;
    if (tid == 0) s_min = 0x7fffffff;
    __syncthreads();
    int lmin = 0x7fffffff;
    for (int i = tid; i < n1; i += 256) {
        int j = nn12[i];
        int ok = (j >= 0 && j < n2 && nn21[j] == i) ? 1 : 0;
        flag[i] = ok;
        if (ok) lmin = min(lmin, d12[i]);
    }
    atomicMin(&s_min, lmin);
    __syncthreads();
    if (apply_filter) {
        const float thr = fminf(__fmul_rn(ratio, (float)s_min), minTh);
        for (int i = tid; i < n1; i += 256)
            if (flag[i] && !((float)d12[i] < thr)) flag[i] = 0;
        __syncthreads();
    }
    // ordered compaction; flags live in global scratch, scan through a shared staging chunk
    __shared__ int s_chunk[2048];
    __shared__ int s_run;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < n1; base += 2048) {
        int m = min(2048, n1 - base);
        for (int i = tid; i < m; i += 256) s_chunk[i] = flag[base + i];
        __syncthreads();
        int tot = scan256(s_chunk, m);
        int run = s_run;
        for (int i = tid; i < m; i += 256) {
            if (flag[base + i]) {
                int pos = run + s_chunk[i];
                if (pos < capacity) {
                    tb_match mt = {base + i, nn12[base + i], 0, (float)d12[base + i]};
                    out[(size_t)p * capacity + pos] = mt;
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_run = run + tot;
        __syncthreads();
    }
    if (tid == 0) out_count[p] = s_run;
}

__global__ void hamming_pairs(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int n, int *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a0 = a[2 * i], a1 = a[2 * i + 1];
    uint32_t q[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    out[i] = ham256(q, b[2 * i], b[2 * i + 1]);
}

// ------------------------------------------------------------------------------------------ host side
struct MatchWs { // lazily grown per-ctx workspace
    PairDesc *pairs = nullptr; int pairs_cap = 0;
    int *best_idx = nullptr, *best_dist = nullptr, *second = nullptr, *scratch = nullptr; size_t ws_cap = 0;
    int *frames = nullptr; int frames_cap = 0;
    std::vector<int32_t> frames_shadow;   // host copy of what the device frame lists hold (skip identical re-uploads)
    tb_match *out = nullptr; int *out_count = nullptr; size_t out_cap = 0;
};
static MatchWs &ws_of(tb_ctx *ctx)
{
    if (!ctx->mws) ctx->mws = new MatchWs();
    return *ctx->mws;
}
void tb_match_ws_free(tb_ctx *ctx)
{
    MatchWs *w = ctx->mws;
    if (!w) return;
    cudaFree(w->pairs); cudaFree(w->best_idx); cudaFree(w->best_dist); cudaFree(w->second); cudaFree(w->scratch);
    cudaFree(w->frames); cudaFree(w->out); cudaFree(w->out_count);
    delete w;
    ctx->mws = nullptr;
}

static int ws_reserve(tb_ctx *ctx, int n_pairs2, size_t rows)
{
    MatchWs &w = ws_of(ctx);
    int rc;
    if (n_pairs2 > w.pairs_cap) {
        w.pairs_cap = 0;
        if ((rc = tb_regrow(ctx, w.pairs, (size_t)n_pairs2))) return rc;
        w.pairs_cap = n_pairs2;
    }
    if (rows > w.ws_cap) {
        w.ws_cap = 0;
        if ((rc = tb_regrow(ctx, w.best_idx, rows)) || (rc = tb_regrow(ctx, w.best_dist, rows)) || (rc = tb_regrow(ctx, w.second, rows)) ||
            (rc = tb_regrow(ctx, w.scratch, rows))) return rc;
        w.ws_cap = rows;
    }
    return TB_OK;
}

extern "C" int tb_hamming_pairs_device(tb_ctx *ctx, const void *a, const void *b, int n, void *dist)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || n < 0) return TB_E_INVALID;
    if (n == 0) return TB_OK;
    hamming_pairs<<<div_up(n, 256), 256, 0, ctx->stream>>>((const uint4 *)a, (const uint4 *)b, n, (int *)dist);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

extern "C" int tb_knn2_device(tb_ctx *ctx, const void *q, int nq, const void *t, int nt, void *idx, void *best, void *second)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || nq < 0 || nt < 0) return TB_E_INVALID;
    if (nq == 0) return TB_OK;
    int rc = ws_reserve(ctx, 2, 1);
    if (rc) return rc;
    MatchWs &w = ws_of(ctx);
    PairDesc pd = {(const uint8_t *)q, (const uint8_t *)t, nq, nt};
    if (ctx->match_impl == 1 && nt > 0) {
        // the train set in slices where that fills the SMs better (config 4 sharded over 8 GPUs: 196 CTAs for 148 SMs)
        static const int forced = getenv("TB_KNN_SPLITS") ? atoi(getenv("TB_KNN_SPLITS")) : 0;
        const int S = forced > 0 ? std::min(forced, 8) : tb_knn2_tc_splits(nq, nt, ctx->num_sms);
        if (S <= 1) {
            if ((rc = tb_stage_upload(ctx, w.pairs, &pd, sizeof pd))) return rc;
            return tb_knn2_tc(ctx, w.pairs, nq, 1, (int *)idx, (int *)best, (int *)second);
        }
        if ((rc = ws_reserve(ctx, 8, (size_t)S * nq))) return rc;
        const int rows = (div_up(nt, S) + 127) & ~127;            // slices start at multiples of the tile height
        PairDesc ps[8];
        int used = 0;
        for (int s = 0; s < S && s * rows < nt; s++, used++)
            ps[s] = PairDesc{(const uint8_t *)q, (const uint8_t *)t + (size_t)s * rows * 32, nq, std::min(rows, nt - s * rows)};
        if ((rc = tb_stage_upload(ctx, w.pairs, ps, sizeof(PairDesc) * used))) return rc;
        if ((rc = tb_knn2_tc(ctx, w.pairs, nq, used, w.best_idx, w.best_dist, w.second))) return rc;
        return tb_knn2_tc_merge(ctx, w.best_idx, w.best_dist, w.second, nq, used, rows, (int *)idx, (int *)best, (int *)second);
    }
    if ((rc = tb_stage_upload(ctx, w.pairs, &pd, sizeof pd))) return rc;
    dim3 grid(div_up(nq, KNN_THREADS), 1);
    ProfScope prof__(ctx, "hamming_knn");
    hamming_knn<true><<<grid, KNN_THREADS, 0, ctx->stream>>>(w.pairs, 0, (int *)idx, (int *)best, (int *)second);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

static int bf_single(tb_ctx *ctx, const void *d1, int n1, const void *d2, int n2, int filter, float ratio, float minTh,
                     tb_match *out_host, int capacity, int *n)
{
    if (!ctx || n1 < 0 || n2 < 0 || capacity < 0 || !n) return TB_E_INVALID;
    *n = 0;
    if (n1 == 0 || n2 == 0) return TB_OK;
    size_t stride = (size_t)(n1 > n2 ? n1 : n2);
    int rc = ws_reserve(ctx, 2, 2 * stride);
    if (rc) return rc;
    MatchWs &w = ws_of(ctx);
    if ((size_t)n1 > w.out_cap) {
        w.out_cap = 0;
        if ((rc = tb_regrow(ctx, w.out, (size_t)n1)) || (rc = tb_regrow(ctx, w.out_count, 4))) return rc;
        w.out_cap = n1;
    }
    PairDesc pd[2] = {{(const uint8_t *)d1, (const uint8_t *)d2, n1, n2}, {(const uint8_t *)d2, (const uint8_t *)d1, n2, n1}};
    if ((rc = tb_stage_upload(ctx, w.pairs, pd, sizeof pd))) return rc;
    if (ctx->match_impl == 1 && stride >= TB_TC_MIN_ROWS) {
        // tcgen05 int8 matcher (the default) once the problem fills at least one 256-row CTA; same best_idx/best_dist layout
        if ((rc = tb_match_knn_tc(ctx, 1, w.pairs, (int)stride, w.best_idx, w.best_dist))) return rc;
    } else {
        dim3 grid(div_up((int)stride, KNN_THREADS), 2);
        ProfScope prof__(ctx, "hamming_knn");
        hamming_knn<false><<<grid, KNN_THREADS, 0, ctx->stream>>>(w.pairs, (int)stride, w.best_idx, w.best_dist, nullptr);
        TB_LAUNCH_CHECK(ctx);
    }
    bf_cross_filter<<<1, 256, 0, ctx->stream>>>(w.pairs, 1, (int)stride, w.best_idx, w.best_dist, w.scratch, filter, ratio, minTh,
                                                w.out, n1, w.out_count);
    TB_LAUNCH_CHECK(ctx);
    int cnt = 0;
    TB_CUDA(ctx, cudaMemcpyAsync(&cnt, w.out_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    int ncopy = cnt < capacity ? cnt : capacity;
    if (ncopy > 0) TB_CUDA(ctx, cudaMemcpy(out_host, w.out, sizeof(tb_match) * ncopy, cudaMemcpyDeviceToHost));
    return cnt > capacity ? tb_fail(ctx, TB_E_CAPACITY, "match output capacity too small") : TB_OK;
}

extern "C" int tb_match_bf_cross_device(tb_ctx *ctx, const void *d1, int n1, const void *d2, int n2, tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    return bf_single(ctx, d1, n1, d2, n2, 0, 0.f, 0.f, out, capacity, n);
}

extern "C" int tb_search_by_bf_device(tb_ctx *ctx, const void *d1, int n1, const void *d2, int n2, float ratio, float minTh,
                                      tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    return bf_single(ctx, d1, n1, d2, n2, 1, ratio, minTh, out, capacity, n);
}

static int bf_host(tb_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2, int filter, float ratio, float minTh,
                   tb_match *out, int capacity, int *n)
{
    if (!ctx || n1 < 0 || n2 < 0 || !n) return TB_E_INVALID;
    *n = 0;
    if (n1 == 0 || n2 == 0) return TB_OK;
    ArenaScope buf(ctx);
    uint8_t *dd1 = buf.upload(d1, (size_t)32 * n1), *dd2 = buf.upload(d2, (size_t)32 * n2);
    if (!dd1 || !dd2) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    return bf_single(ctx, dd1, n1, dd2, n2, filter, ratio, minTh, out, capacity, n);
}
extern "C" int tb_match_bf_cross_host(tb_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2, tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    return bf_host(ctx, d1, n1, d2, n2, 0, 0.f, 0.f, out, capacity, n);
}
extern "C" int tb_search_by_bf_host(tb_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2, float ratio, float minTh,
                                    tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    return bf_host(ctx, d1, n1, d2, n2, 1, ratio, minTh, out, capacity, n);
}
extern "C" int tb_hamming_pairs_host(tb_ctx *ctx, const uint8_t *a, const uint8_t *b, int n, int32_t *dist)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || n < 0 || !dist) return TB_E_INVALID;
    if (n == 0) return TB_OK;
    ArenaScope buf(ctx);
    uint8_t *da = buf.upload(a, (size_t)32 * n), *db = buf.upload(b, (size_t)32 * n);
    int *dd = buf.alloc<int>((size_t)n);
    if (!da || !db || !dd) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    int rc = tb_hamming_pairs_device(ctx, da, db, n, dd);
    if (rc) return rc;
    TB_CUDA(ctx, cudaMemcpyAsync(dist, dd, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

// Batched searchByBF over frames of one extract result (called by frontend.cu).
int tb_match_bf_batched(tb_ctx *ctx, int n_pairs, const int32_t *q_frames_host, const int32_t *t_frames_host, int n_frames_total,
                        const uint8_t *desc, size_t frame_stride, const int *counts, int kp_cap,
                        float ratio, float minTh, tb_match *matches, int capacity, int *out_counts)
{
    if (n_pairs <= 0) return TB_OK;
    int rc = ws_reserve(ctx, 2 * n_pairs, (size_t)2 * n_pairs * kp_cap);
    if (rc) return rc;
    MatchWs &w = ws_of(ctx);
    if (2 * n_pairs > w.frames_cap) {
        w.frames_cap = 0;
        w.frames_shadow.clear();
        if ((rc = tb_regrow(ctx, w.frames, (size_t)2 * n_pairs))) return rc;
        w.frames_cap = 2 * n_pairs;
    }
    // The frame index lists are tiny, but a host-to-device copy queues on the same copy engine as the NEXT batch's frame
    // upload (already running in a pipelined caller) and would hold this batch's matcher back until that upload ends:
    // re-upload only when the lists changed (a streaming caller passes the same k <-> k+1 lists every step).
    const bool same = w.frames_shadow.size() == (size_t)2 * n_pairs &&
                      memcmp(w.frames_shadow.data(), q_frames_host, sizeof(int32_t) * n_pairs) == 0 &&
                      memcmp(w.frames_shadow.data() + n_pairs, t_frames_host, sizeof(int32_t) * n_pairs) == 0;
    if (!same) {
        w.frames_shadow.assign(q_frames_host, q_frames_host + n_pairs);
        w.frames_shadow.insert(w.frames_shadow.end(), t_frames_host, t_frames_host + n_pairs);
        if ((rc = tb_stage_upload(ctx, w.frames, w.frames_shadow.data(), sizeof(int) * 2 * n_pairs))) return rc;   // through the pinned ring: no wait
    }
    build_pairs<<<div_up(n_pairs, 128), 128, 0, ctx->stream>>>(w.pairs, n_pairs, w.frames, w.frames + n_pairs, desc, frame_stride, counts);
    TB_LAUNCH_CHECK(ctx);
    if (ctx->match_impl == 1) {
        rc = tb_match_knn_tc(ctx, n_pairs, w.pairs, kp_cap, w.best_idx, w.best_dist);
        if (rc) return rc;
    } else {
        dim3 grid(div_up(kp_cap, KNN_THREADS), 2 * n_pairs);
        ProfScope prof__(ctx, "hamming_knn");
        hamming_knn<false><<<grid, KNN_THREADS, 0, ctx->stream>>>(w.pairs, kp_cap, w.best_idx, w.best_dist, nullptr);
        TB_LAUNCH_CHECK(ctx);
    }
    ProfScope prof2__(ctx, "bf_cross_filter");
    bf_cross_filter<<<n_pairs, 256, 0, ctx->stream>>>(w.pairs, n_pairs, kp_cap, w.best_idx, w.best_dist, w.scratch, 1, ratio, minTh,
                                                      matches, capacity, out_counts);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

extern "C" void tb_three_maxima(const int32_t *hs, int L, int *ind1, int *ind2, int *ind3)
{
    // Matcher::ComputeThreeMaxima (matcher.cpp:810-851)
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = hs[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { *ind3 = -1; }
}
