// match_tc.cu — K8: brute-force Hamming nearest neighbour as an exact +-1 int8 tensor-core contraction.
// Same outputs as K7 (match.cu: hamming_knn<false>), i.e. the arg-min / min of
// Matcher::DescriptorDistance (reference src/matchers/matcher.cpp:793-808) over the train set as
// cv::BFMatcher::match computes it (matcher.cpp:20,207), lowest train index on ties.
//
// Formulation: descriptor bit b -> (1 - 2b) in int8, K = 256.  dot(q,t) = 256 - 2*hamming(q,t), so
// hamming = (256 - dot) / 2 exactly (all products are +-1, |sum| <= 256, int32 accumulation).
// This is the only place the path is a dense contraction (north_star (c)).
//
//  * expand_desc_i8 writes every frame's descriptors once as int8 in the UMMA canonical K-major
//    no-swizzle layout [row/8][k/16][row%8][16 B], so a 128-row (32 KB) or 256-row (64 KB) operand tile is
//    one contiguous block of global memory and is staged with cp.async.bulk (TMA 1-D bulk copy).
//  * hamming_i8mma: one CTA = 128 query rows.  Warp 0 = bulk-copy producer (2-stage ring of train
//    tiles), warp 1 = tcgen05.mma issuer (kind::i8, M=128, N=256, 8 K-steps of 32), accumulators in TMEM
//    (2 x 256 columns, double buffered), warps 2-5 = epilogue: tcgen05.ld 32 columns at a time, running
//    (max dot, arg) per row in registers.  mbarriers connect the three roles; tcgen05.commit frees the
//    shared-memory stage and publishes the accumulator.
#include "common.cuh"
#include "kernels.cuh"

#define TC_M 128          // rows of one MMA (TMEM lanes)
#define TC_MB_MAX 2       // M blocks per CTA (template parameter MB): with 2, 256 query rows share every train tile, which halves
                          // the L2 -> shared-memory traffic that bounds the batched frame-pair matcher; the 200k x 200k kNN-2
                          // keeps 1 (its top-2 epilogue is the bound and 128-row CTAs fill the last wave better)
#define TC_PAD 256        // expanded frames are padded to a multiple of this many rows
#define TC_KB 256         // bytes (= int8 elements) per expanded descriptor
#define TC_THREADS 320      // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quarter)

struct TcPair {
    const int8_t *q, *t;   // expanded descriptor blocks of the query / train frame
    int nq, nt;
};

// ---- expansion: 32-byte descriptors -> 256 int8 (+1 for a 0 bit, -1 for a 1 bit), rows >= count are 0
__global__ void __launch_bounds__(256)
expand_desc_i8(const uint8_t *__restrict__ desc, size_t in_stride, const int *__restrict__ counts, int n_fixed,
               int8_t *__restrict__ out, size_t out_stride, int rows_pad)
{
    const int f = blockIdx.y;
    const int item = blockIdx.x * 256 + threadIdx.x;       // (row, 16-bit chunk)
    const int row = item >> 4, c = item & 15;
    if (row >= rows_pad) return;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < (counts ? counts[f] : n_fixed)) {
        const uint8_t *d = desc + (size_t)f * in_stride + (size_t)32 * row + 2 * c;
        const uint32_t bits = d[0] | ((uint32_t)d[1] << 8);
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t nib = (bits >> (4 * k)) & 0xF;
            const uint32_t b = (nib * 0x00204081u) & 0x01010101u;   // bit i -> byte i
            w[k] = 0x01010101u + 0xFEu * b;                           // per byte: 0 -> 0x01 (+1), 1 -> 0xFF (-1); no carries
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    int8_t *o = out + (size_t)f * out_stride + (size_t)(row >> 3) * (16 * 128) + (size_t)c * 128 + (size_t)(row & 7) * 16;
    *reinterpret_cast<uint4 *>(o) = v;
}

__global__ void build_tc_pairs(TcPair *pairs, int P, const int *q_frames, const int *t_frames,
                               const int8_t *exp, size_t frame_stride, const int *counts)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int qf = q_frames[p], tf = t_frames[p];
    TcPair a = {exp + qf * frame_stride, exp + tf * frame_stride, counts[qf], counts[tf]};
    TcPair b = {a.t, a.q, a.nt, a.nq};
    pairs[p] = a;
    pairs[P + p] = b;
}

// ---- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: 8x16B core matrices, LBO = distance between the
// two K-adjacent core matrices (128 B), SBO = distance between 8-row groups (2048 B), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct __align__(16) TcShared {
    uint64_t fullA, full[2], empty[2], tfull[2], tempty[2];
    uint32_t tmem_base;
    int part_best[TC_M], part_second[TC_M], part_idx[TC_M];   // MB = 1: partial results of the upper-half-column warps
};

#define TC_SMEM_BYTES (160 * 1024 + sizeof(TcShared) + 64)   // MB = 1: A 32 KB + B 2 x 64 KB; MB = 2: A 64 KB + B 2 x 32 KB
extern __shared__ __align__(1024) uint8_t tc_smem[];

template <bool SECOND, int MB>
__global__ void __launch_bounds__(TC_THREADS, 1)
hamming_i8mma(const TcPair *__restrict__ pairs, int out_stride, int *__restrict__ best_idx, int *__restrict__ best_dist,
              int *__restrict__ second_dist)
{
    const TcPair pd = pairs[blockIdx.y];
    constexpr int TC_ROWS = TC_M * MB;
    constexpr int TN = MB == 1 ? 256 : 128;            // train rows per tile: TMEM holds 2 stages x MB blocks x TN = 512 columns
    constexpr int TB_BYTES = TN * TC_KB;
    const int q0 = blockIdx.x * TC_ROWS;
    if (q0 >= pd.nq) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *sA = tc_smem;                                  // 2 x 32 KB (two M blocks)
    uint8_t *sB = tc_smem + TC_ROWS * TC_KB;                // two stages of TN rows
    TcShared *S = reinterpret_cast<TcShared *>(tc_smem + TC_ROWS * TC_KB + 2 * TB_BYTES);
    const int NT = (pd.nt + TN - 1) / TN;

    if (threadIdx.x == 0) {
        mbar_init(&S->fullA, 1);
        for (int s = 0; s < 2; s++) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1); mbar_init(&S->tfull[s], 1); mbar_init(&S->tempty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // TMEM: 512 columns = 2 stages x 2 M blocks of 128x128 int32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S->tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S->tmem_base;
    int best = -1000, second = -1000, bidx = -1;   // running (max dot, second, arg) of an epilogue thread

    if (warp == 0) {
        if (lane == 0) {
            // ---- producer
            const int8_t *qa = pd.q + (size_t)q0 * TC_KB;   // rows q0.. are contiguous in the canonical layout
            mbar_expect_tx(&S->fullA, TC_ROWS * TC_KB);
            for (int c = 0; c < TC_ROWS * TC_KB; c += 16384) bulk_g2s(sA + c, qa + c, 16384, &S->fullA);
            for (int t = 0; t < NT; t++) {
                const int s = t & 1, ph = (t >> 1) & 1;
                mbar_wait(&S->empty[s], ph ^ 1);
                mbar_expect_tx(&S->full[s], TB_BYTES);
                const int8_t *tb = pd.t + (size_t)t * TB_BYTES;
                for (int c = 0; c < TB_BYTES; c += 16384) bulk_g2s(sB + s * TB_BYTES + c, tb + c, 16384, &S->full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer: idesc = D s32 | A s8 | B s8 | K-major both | N=256 | M=128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            mbar_wait(&S->fullA, 0);
            const uint64_t adesc0 = make_desc(smem_u32(sA)), adesc1 = make_desc(smem_u32(sA + TC_M * TC_KB));
            for (int t = 0; t < NT; t++) {
                const int s = t & 1, ph = (t >> 1) & 1;
                mbar_wait(&S->full[s], ph);
                mbar_wait(&S->tempty[s], ph ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = make_desc(smem_u32(sB + s * TB_BYTES));
#pragma unroll
                for (int k = 0; k < TC_KB / 32; k++)   // K step = 32 bytes = two 16-byte core-matrix columns = 256 B in this layout
                    mma_i8(tmem + (s * MB + 0) * TN, adesc0 + (uint64_t)((k * 256) >> 4), bdesc + (uint64_t)((k * 256) >> 4), idesc, k > 0 ? 1u : 0u);
                if (MB > 1)
#pragma unroll
                for (int k = 0; k < TC_KB / 32; k++)
                    mma_i8(tmem + (s * MB + 1) * TN, adesc1 + (uint64_t)((k * 256) >> 4), bdesc + (uint64_t)((k * 256) >> 4), idesc, k > 0 ? 1u : 0u);
                tc_commit(&S->empty[s]);    // shared-memory stage reusable once these MMAs retire
                tc_commit(&S->tfull[s]);    // accumulator ready for the epilogue
            }
        }
    } else {
        // ---- epilogue warps 2..9: TMEM lanes 32*(warp%4) .. +31; warps 2..5 own the query rows of M block 0,
        //      warps 6..9 those of M block 1 — one thread per query row, all columns
        // (MB = 1: both warps of a lane quarter share the rows and split every tile's columns; merged at the end)
        const int lane_base = 32 * (warp & 3);
        const int grp = (warp - 2) >> 2;
        const int mb = MB > 1 ? grp : 0;
        const int cbeg = MB > 1 ? 0 : grp * (TN / 2), cend = MB > 1 ? TN : cbeg + TN / 2;
        for (int t = 0; t < NT; t++) {
            const int s = t & 1, ph = (t >> 1) & 1;
            mbar_wait(&S->tfull[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int col0 = t * TN;
#pragma unroll 1
            for (int c = cbeg; c < cend; c += 32) {
                if (col0 + c >= pd.nt) break;   // warp-uniform
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + (uint32_t)((s * MB + mb) * TN + c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int lim = pd.nt - (col0 + c);   // valid columns in this chunk
                if (!SECOND && lim >= 32) {
                    // best-only, full chunk: the epilogue is bound by the integer ALU pipe (compare + two selects per
                    // element), so fold the column into the key with one IMAD each (dot * 32 + 31 - j: larger key = larger
                    // dot, then lower column) and reduce the 32 keys with a max3 tree: ~0.6 ALU instructions per element
                    int k[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) k[j] = (int)v[j] * 32 + (31 - j);
#pragma unroll
                    for (int j = 0; j < 30; j += 3) k[j / 3] = __vimax3_s32(k[j], k[j + 1], k[j + 2]);        // 30 -> 10
                    const int a0 = __vimax3_s32(k[0], k[1], k[2]), a1 = __vimax3_s32(k[3], k[4], k[5]), a2 = __vimax3_s32(k[6], k[7], k[8]);
                    const int a3 = __vimax3_s32(k[9], k[30], k[31]);
                    const int ck = max(__vimax3_s32(a0, a1, a2), a3);
                    const int cd = ck >> 5;                              // arithmetic shift: floor, exact for negative dots too
                    if (cd > best) { best = cd; bidx = col0 + c + 31 - (ck & 31); }
                    continue;
                }
                if (SECOND && lim >= 32) {
                    // best + second, full chunk: same keys, reduced by a tree of (largest, second largest) pairs —
                    // 2 ops per leaf pair, 3 per merge (max, min, max3): ~2.4 ALU instructions per element instead of ~6
                    int hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int x = (int)v[2 * j] * 32 + (31 - 2 * j), y = (int)v[2 * j + 1] * 32 + (30 - 2 * j);
                        hi[j] = max(x, y); lo[j] = min(x, y);
                    }
#pragma unroll
                    for (int n = 8; n >= 1; n >>= 1) {
#pragma unroll
                        for (int j = 0; j < n; j++) {
                            const int h2 = max(hi[j], hi[j + n]);
                            lo[j] = __vimax3_s32(min(hi[j], hi[j + n]), lo[j], lo[j + n]);
                            hi[j] = h2;
                        }
                    }
                    const int cd = hi[0] >> 5, cs = lo[0] >> 5;
                    if (cd > best) { second = max(best, cs); best = cd; bidx = col0 + c + 31 - (hi[0] & 31); }
                    else second = max(second, cd);
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int d = (int)v[j];
                    if (SECOND) {
                        if (j < lim) {
                            if (d > best) { second = best; best = d; bidx = col0 + c + j; }
                            else if (d > second) second = d;
                        }
                    } else {
                        if (j < lim && d > best) { best = d; bidx = col0 + c + j; }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->tempty[s]);
        }
        if (MB == 1 && grp == 1) { S->part_best[lane_base + lane] = best; S->part_second[lane_base + lane] = second; S->part_idx[lane_base + lane] = bidx; }
        if (MB > 1) {
            const int row = q0 + mb * TC_M + lane_base + lane;
            if (row < pd.nq) {
                const size_t o = (size_t)blockIdx.y * out_stride + row;
                best_idx[o] = bidx;
                best_dist[o] = bidx >= 0 ? (256 - best) >> 1 : 0x7fffffff;
                if (SECOND) second_dist[o] = second > -1000 ? (256 - second) >> 1 : 0x7fffffff;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    if (MB == 1 && warp >= 2 && warp < 6) {
        // merge the two column halves (lowest index wins ties) and write the row's result
        const int r = 32 * (warp & 3) + lane, row = q0 + r;
        const int b1 = S->part_best[r], s1 = S->part_second[r], i1 = S->part_idx[r];
        if (i1 >= 0 && (bidx < 0 || b1 > best || (b1 == best && i1 < bidx))) {
            second = max(s1, best); best = b1; bidx = i1;
        } else if (i1 >= 0) {
            second = max(second, b1);
        }
        if (row < pd.nq) {
            const size_t o = (size_t)blockIdx.y * out_stride + row;
            best_idx[o] = bidx;
            best_dist[o] = bidx >= 0 ? (256 - best) >> 1 : 0x7fffffff;
            if (SECOND) second_dist[o] = second > -1000 ? (256 - second) >> 1 : 0x7fffffff;
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
struct TcWs {
    int8_t *exp = nullptr; size_t exp_bytes = 0;
    TcPair *pairs = nullptr; int pairs_cap = 0;
    int *frames = nullptr; int frames_cap = 0;
};
static TcWs *tcws(tb_ctx *ctx)
{
    if (!ctx->tcws) ctx->tcws = new TcWs();
    return ctx->tcws;
}
void tb_match_tc_free(tb_ctx *ctx)
{
    TcWs *w = ctx->tcws;
    if (!w) return;
    cudaFree(w->exp); cudaFree(w->pairs); cudaFree(w->frames);
    delete w;
    ctx->tcws = nullptr;
}

// Fills best_idx/best_dist[2P][kp_cap] exactly like hamming_knn<false> does for the pair table of match.cu.
int tb_match_knn_tc(tb_ctx *ctx, int n_pairs, const int *d_q_frames, const int *d_t_frames, int n_frames_used,
                    const uint8_t *desc, size_t frame_stride, const int *counts, int kp_cap, int *best_idx, int *best_dist)
{
    TcWs *w = tcws(ctx);
    const int rows_pad = (kp_cap + TC_PAD - 1) / TC_PAD * TC_PAD;
    const size_t exp_stride = (size_t)rows_pad * TC_KB;
    const size_t need = exp_stride * (size_t)n_frames_used;
    if (need > w->exp_bytes) {
        if (w->exp) cudaFree(w->exp);
        TB_CUDA(ctx, cudaMalloc(&w->exp, need));
        w->exp_bytes = need;
    }
    if (2 * n_pairs > w->pairs_cap) {
        if (w->pairs) cudaFree(w->pairs);
        TB_CUDA(ctx, cudaMalloc(&w->pairs, sizeof(TcPair) * 2 * n_pairs));
        w->pairs_cap = 2 * n_pairs;
    }
    {
        ProfScope prof__(ctx, "expand_desc_i8");
        dim3 grid(div_up(rows_pad * 16, 256), n_frames_used);
        expand_desc_i8<<<grid, 256, 0, ctx->stream>>>(desc, frame_stride, counts, 0, w->exp, exp_stride, rows_pad);
        TB_LAUNCH_CHECK(ctx);
    }
    build_tc_pairs<<<div_up(n_pairs, 128), 128, 0, ctx->stream>>>(w->pairs, n_pairs, d_q_frames, d_t_frames, w->exp, exp_stride, counts);
    TB_LAUNCH_CHECK(ctx);
    const size_t smem = TC_SMEM_BYTES;
    {
        ProfScope prof__(ctx, "hamming_i8mma");
        dim3 grid(div_up(kp_cap, TC_M * 2), 2 * n_pairs);
        hamming_i8mma<false, 2><<<grid, TC_THREADS, smem, ctx->stream>>>(w->pairs, kp_cap, best_idx, best_dist, nullptr);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

// best / second / argbest of nq query rows against nt train rows (tb_knn2_device with the tensor-core matcher)
int tb_knn2_tc(tb_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt, int *idx, int *best, int *second)
{
    TcWs *w = tcws(ctx);
    const size_t q_rows = (size_t)(nq + TC_PAD - 1) / TC_PAD * TC_PAD, t_rows = (size_t)(nt + TC_PAD - 1) / TC_PAD * TC_PAD;
    const size_t need = (q_rows + t_rows) * TC_KB;
    if (need > w->exp_bytes) {
        if (w->exp) cudaFree(w->exp);
        w->exp = nullptr; w->exp_bytes = 0;
        TB_CUDA(ctx, cudaMalloc(&w->exp, need));
        w->exp_bytes = need;
    }
    if (w->pairs_cap < 2) {
        if (w->pairs) cudaFree(w->pairs);
        TB_CUDA(ctx, cudaMalloc(&w->pairs, sizeof(TcPair) * 2));
        w->pairs_cap = 2;
    }
    int8_t *eq = w->exp, *et = w->exp + q_rows * TC_KB;
    {
        ProfScope prof__(ctx, "expand_desc_i8");
        expand_desc_i8<<<dim3(div_up((int)q_rows * 16, 256), 1), 256, 0, ctx->stream>>>(q, 0, nullptr, nq, eq, 0, (int)q_rows);
        TB_LAUNCH_CHECK(ctx);
        expand_desc_i8<<<dim3(div_up((int)t_rows * 16, 256), 1), 256, 0, ctx->stream>>>(t, 0, nullptr, nt, et, 0, (int)t_rows);
        TB_LAUNCH_CHECK(ctx);
    }
    TcPair pd = {eq, et, nq, nt};
    TB_CUDA(ctx, cudaMemcpyAsync(w->pairs, &pd, sizeof pd, cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = TC_SMEM_BYTES;
    {
        ProfScope prof__(ctx, "hamming_i8mma");
        hamming_i8mma<true, 1><<<dim3(div_up(nq, TC_M), 1), TC_THREADS, smem, ctx->stream>>>(w->pairs, 0, idx, best, second);
        TB_LAUNCH_CHECK(ctx);
    }
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the pair descriptor came from pageable host memory
    return TB_OK;
}

void tb_match_tc_init_device()
{
    const int smem = (int)(TC_SMEM_BYTES);
    cudaFuncSetAttribute(hamming_i8mma<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(hamming_i8mma<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}
