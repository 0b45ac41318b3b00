// match_tc.cu — K8: brute-force Hamming nearest neighbour as an exact +-1 int8 tensor-core contraction.
// Same outputs as K7 (match.cu: hamming_knn<false>), i.e. the arg-min / min of
// Matcher::DescriptorDistance (reference src/matchers/matcher.cpp:793-808) over the train set as
// cv::BFMatcher::match computes it (matcher.cpp:20,207), lowest train index on ties.
//
// Formulation: descriptor bit b -> (1 - 2b) in int8, K = 256.  dot(q,t) = 256 - 2*hamming(q,t), so
// hamming = (256 - dot) / 2 exactly (all products are +-1, |sum| <= 256, int32 accumulation).
// This is the only place the path is a dense contraction (north_star (c)).
//
//  * The +-1 int8 operands never exist in global memory: expander warps read the packed 32-byte descriptors (8 KB per
//    256 rows) and write the expanded rows straight into shared memory in the UMMA canonical K-major no-swizzle layout
//    [row/8][k/16][row%8][16 B] (generic-proxy stores + fence.proxy.async + mbarrier arrive).  The first version staged
//    pre-expanded descriptors (256 B per row) from global memory with cp.async.bulk and was bound by that 8x L2 -> shared
//    traffic (every CTA streams its whole train frame).
//  * hamming_i8mma: one CTA = 128 (kNN-2) or 256 (frame pairs) query rows.  Warp 0 owns the TMEM allocation, warp 1 =
//    tcgen05.mma issuer (kind::i8, M=128, N=TN, 9 K-steps of 32), accumulators in TMEM (2 stages x MB blocks x TN columns),
//    warps 2-9 = epilogue: tcgen05.ld 32 columns at a time, running (max dot, arg) per row in registers, warps 10-13 =
//    expanders (the ring of train tiles).  mbarriers connect the roles; tcgen05.commit frees the shared-memory stage and
//    publishes the accumulator.
//  * The query block (the A operand) lives in TENSOR MEMORY for the whole CTA: the epilogue warps expand their own query
//    row once (lane = row, 72 columns = 288 int8) and write it with tcgen05.st; the MMAs then read A from TMEM
//    (tcgen05.mma [d], [a], b-desc) and only the train tile from shared memory.  With both operands in shared memory an
//    M128 x N128 x K32 step reads 8 KB per 64 tensor cycles = the whole 128 B/clk of the shared-memory pipe, which the
//    expanders' stores and table look-ups share (ncu: pipe 90 % busy, tensor pipe 58 %); with A in TMEM it reads half.
#include "common.cuh"
#include "kernels.cuh"

#define TC_M 128          // rows of one MMA (TMEM lanes)
#define TC_MB_MAX 2       // M blocks per CTA (template parameter MB): with 2, 256 query rows share every train tile, which halves
                          // the L2 -> shared-memory traffic that bounds the batched frame-pair matcher; the 200k x 200k kNN-2
                          // keeps 1 (its top-2 epilogue is the bound and 128-row CTAs fill the last wave better)
#define TC_KB 288         // bytes (= int8 elements) per expanded row: 256 descriptor bits + one more K step of 32 that carries the
                          // tie-break key (see below)
#define TC_SBO (TC_KB / 16 * 128)   // canonical layout: distance between 8-row groups
#define TC_THREADS 640    // warps 0..3 MMA issuers (warp 0 also owns the TMEM allocation), warps 4..11 epilogue (two per TMEM lane quarter), warps 12..19 expanders (two groups of 4: even / odd tiles)
#define TC_EXP_THREADS 128
#define TC_EXP_WARP0 12
#define TC_EPI_WARP0 4
#define TC_LUT_BYTES (256 * 16 * 8)   // byte -> 8 x int8 table, 16 replicas (one per lane of a half warp: conflict-free LDS.64)
#define TC_ACOLS (TC_KB / 4)          // TMEM columns of one expanded query block (A operand): 4 int8 per 32-bit cell
#define TC_NSB_MAX 6

// The key of the arg-max is produced by the tensor core itself.  Query rows are expanded to +-32 and carry a 1 in element
// 256; train row j is expanded to +-1 and carries 31 - j%32 in element 256 (elements 257..287 are zero).  The accumulator is
// then 32*dot + 31 - j%32: within a 32-column chunk a larger value means a larger dot product and, on ties, the lower column —
// the epilogue only takes maxima (the first version spent one IMAD per element building these keys on the integer pipes,
// which ncu showed to be the bound: ALU pipe 68 % busy, tensor pipe 45 %).
//
// Layout of an expanded row r inside its tile: chunk c (16 bytes = 16 K elements) at (r/8)*TC_SBO + c*128 + (r%8)*16 — the
// UMMA canonical K-major no-swizzle layout; the eight lanes of a quarter warp (consecutive rows) store 128 contiguous bytes.

// train rows: two table look-ups per chunk (the first arithmetic expansion cost ~21 integer instructions per chunk).
// TC_EXPAND_ARITH: the byte -> 8 x (+-1) expansion without the table, 8 instructions per byte: bit i of a nibble is moved to the
// most significant bit of byte i by ONE multiply (the partial products land on distinct bits, no carries), PRMT's sign-replicate
// mode turns every byte into 0x00 / 0xFF, OR 0x01 makes that +1 / -1.  Takes the 16 KB of table reads per 64-row tile off the
// shared-memory pipe and puts ~8 instructions per byte on the integer pipes — which the expander warps do not have to spare.
#ifndef TC_EXPAND_ARITH
#define TC_EXPAND_ARITH 0   // measured slower than the table (third session: 0.412 vs 0.401 ms batched, 7.45 vs 6.91 ms kNN-2)
#endif
__device__ __forceinline__ uint2 tc_expand_byte(uint32_t xw, int k)   // byte k of xw -> eight +-1 bytes
{
    const uint32_t b = k == 0 ? xw : __byte_perm(xw, 0u, 0x4440 + k);        // byte k in bits 0..7 (k = 0: the masks below suffice)
    const uint32_t lo = (b & 0x0Fu) * 0x10204080u;                           // nibble bit i -> bit 8 i + 7
    const uint32_t hi = (b & 0xF0u) * 0x01020408u;                           // bit 4 + i   -> bit 8 i + 7
    return make_uint2(__byte_perm(lo, 0u, 0xBA98) | 0x01010101u, __byte_perm(hi, 0u, 0xBA98) | 0x01010101u);
}
#if TC_EXPAND_ARITH
#define TC_EXPAND4(x, e0, e1, e2, e3) const uint2 e0 = tc_expand_byte(x, 0), e1 = tc_expand_byte(x, 1), e2 = tc_expand_byte(x, 2), e3 = tc_expand_byte(x, 3)
#else
#define TC_EXPAND4(x, e0, e1, e2, e3) const uint2 e0 = lut[(x & 0xFFu) * 16], e1 = lut[((x >> 8) & 0xFFu) * 16], e2 = lut[((x >> 16) & 0xFFu) * 16], e3 = lut[(x >> 24) * 16]
#endif
__device__ __forceinline__ void expand_row_lut(uint8_t *tile, int r, const uint4 lo, const uint4 hi, bool valid, const uint2 *lut)
{
    uint8_t *o = tile + (size_t)(r >> 3) * TC_SBO + (size_t)(r & 7) * 16;
    if (!valid) {   // rows past the train set: never read back (the epilogue masks their columns); keep the tile defined
#pragma unroll
        for (int c = 0; c < 16; c++) *reinterpret_cast<uint4 *>(o + c * 128) = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint32_t w32[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t x = w32[w];
        TC_EXPAND4(x, e0, e1, e2, e3);
        *reinterpret_cast<uint4 *>(o + (2 * w) * 128) = make_uint4(e0.x, e0.y, e1.x, e1.y);
        *reinterpret_cast<uint4 *>(o + (2 * w + 1) * 128) = make_uint4(e2.x, e2.y, e3.x, e3.y);
    }
}
// a quarter of a train row (tiles of 32 local rows: four expander threads per row): descriptor bytes 8 * part .. + 7 = chunks
// 4 * part .. + 3
__device__ __forceinline__ void expand_quarter_lut(uint8_t *tile, int r, int part, const uint2 d, bool valid, const uint2 *lut)
{
    uint8_t *o = tile + (size_t)(r >> 3) * TC_SBO + (size_t)(r & 7) * 16 + (size_t)part * 4 * 128;
    if (!valid) {
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(o + c * 128) = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint32_t w32[2] = {d.x, d.y};
#pragma unroll
    for (int w = 0; w < 2; w++) {
        const uint32_t x = w32[w];
        TC_EXPAND4(x, e0, e1, e2, e3);
        *reinterpret_cast<uint4 *>(o + (2 * w) * 128) = make_uint4(e0.x, e0.y, e1.x, e1.y);
        *reinterpret_cast<uint4 *>(o + (2 * w + 1) * 128) = make_uint4(e2.x, e2.y, e3.x, e3.y);
    }
}
// chunks 16 and 17 of row r (element 256 = tie-break key 31 - r % 32, elements 257..287 = 0) do not depend on the tile (tiles
// start at multiples of 32 rows): written once per ring stage when the CTA starts
__device__ __forceinline__ void write_key_chunks(uint8_t *tile, int r, bool key, bool zero)
{
    uint8_t *o = tile + (size_t)(r >> 3) * TC_SBO + (size_t)(r & 7) * 16;
    if (key) *reinterpret_cast<uint4 *>(o + 16 * 128) = make_uint4(31u - (uint32_t)(r & 31), 0, 0, 0);
    if (zero) *reinterpret_cast<uint4 *>(o + 17 * 128) = make_uint4(0, 0, 0, 0);
}
// one half of a train row (tiles of 64 rows: two expander threads per row): half 0 = descriptor bytes 0..15 (chunks 0..7),
// half 1 = bytes 16..31 (chunks 8..15)
__device__ __forceinline__ void expand_half_lut(uint8_t *tile, int r, int half, const uint4 d, bool valid, const uint2 *lut)
{
    uint8_t *row = tile + (size_t)(r >> 3) * TC_SBO + (size_t)(r & 7) * 16;
    uint8_t *o = row + (size_t)half * 8 * 128;
    if (!valid) {
#pragma unroll
        for (int c = 0; c < 8; c++) *reinterpret_cast<uint4 *>(o + c * 128) = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint32_t w32[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t x = w32[w];
        TC_EXPAND4(x, e0, e1, e2, e3);
        *reinterpret_cast<uint4 *>(o + (2 * w) * 128) = make_uint4(e0.x, e0.y, e1.x, e1.y);
        *reinterpret_cast<uint4 *>(o + (2 * w + 1) * 128) = make_uint4(e2.x, e2.y, e3.x, e3.y);
    }
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
// two K-adjacent core matrices (128 B), SBO = distance between 8-row groups (TC_SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// one thread of the (converged) warp, chosen by the hardware: unlike `lane == 0`, ptxas knows that the guarded region runs
// in a single thread and keeps the tcgen05 operands in uniform registers without a per-instruction election loop
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// ---- CTA pair (cta_group::2): barrier of the leader CTA (rank 0 of the cluster) addressed from either CTA
__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG> __device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar)
{
    if (CG == 1) { mbar_arrive(bar); return; }
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(raddr) : "r"(smem_u32(bar)));
    // default semantics (.release at CTA scope): a cluster-scope release compiles to MEMBAR.ALL.GPU + ERRBAR in front of every
    // arrival (ncu: 19 % of the stall samples); the data the leader's MMAs read was already handed to the async proxy by
    // fence.proxy.async, and TMEM reads are ordered by the tcgen05 fences
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// wait on a barrier that threads of the peer CTA arrive on as well
template <int CG> __device__ __forceinline__ void mbar_wait_cg(uint64_t *bar, uint32_t parity)
{
    if (CG == 1) { mbar_wait(bar, parity); return; }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int CG> __device__ __forceinline__ void tc_commit_cg(uint64_t *bar)
{
    if (CG == 1) { tc_commit(bar); return; }
    // the arrival lands on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mma_i8_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (lane = row, 4 int8 per 32-bit column), B from shared memory
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct __align__(16) TcShared {
    uint64_t fullA, full[TC_NSB_MAX], empty[TC_NSB_MAX], tfull[3][TC_MB_MAX], tempty[3][TC_MB_MAX];   // [TMEM stage][M block]
    uint32_t tmem_base;
    int part_best[TC_M], part_second[TC_M], part_idx[TC_M];   // MB = 1: partial results of the upper-half-column warps
};

#define TC_SMEM_BYTES (144 * 1024 + TC_LUT_BYTES + sizeof(TcShared) + 64)   // train ring: MB = 1: 4 x 36 KB; MB = 2: 6 x 18 KB; + table
extern __shared__ __align__(1024) uint8_t tc_smem[];

// CG = 2: a CTA PAIR (cluster of two, cta_group::2) shares every train tile.  Each CTA keeps its own MB query blocks in its own
// tensor memory and expands only HALF of the tile's rows into its own shared memory; the tcgen05.mma of the pair (issued by
// the leader CTA's threads, M = 256) feeds both tensor cores with both halves, and each CTA's accumulators cover all TN
// columns.  Per SM that halves the expansion work and the shared-memory reads of the train operand — the pipe that bound the
// single-CTA kernel (84 % busy at a 69 % busy tensor pipe).  Barriers the peer CTA has to signal (query blocks written, tile
// half expanded, accumulator read) live in the leader's shared memory and are reached with mapa + a cluster-scope arrive;
// completion of the MMAs is multicast to both CTAs by tcgen05.commit.
template <bool SECOND, int MB, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
hamming_i8mma(const PairDesc *__restrict__ pairs, int out_stride, int *__restrict__ best_idx, int *__restrict__ best_dist,
              int *__restrict__ second_dist)
{
    const PairDesc pd = pairs[blockIdx.y];
    constexpr int TC_ROWS = TC_M * MB;
    constexpr int TN = MB == 1 ? 128 : 64;             // train rows per tile (columns of an accumulator)
    constexpr int TL = TN / CG;                        // of which this CTA expands and holds TL
    constexpr int TB_BYTES = TL * TC_KB;
    constexpr int NSB = MB == 1 ? 4 : 6;               // shared-memory stages of the train ring (the expanders run ahead of the MMAs)
    constexpr int TS = (MB == 1 && CG == 2) ? 3 : 2;   // accumulator stages: a pair's hand-shakes cross the SM boundary twice per tile
                                                       // (tile expanded, accumulator read), a third stage hides that where TMEM has room
    constexpr int DCOLS = TS * MB * TN;                // TMEM: accumulators = TS stages x MB blocks x TN columns, then MB query blocks
    constexpr int PARTS = TC_EXP_THREADS / TL;         // expander threads per train row (1, 2 or 4)
    static_assert(DCOLS + MB * TC_ACOLS <= 512 && NSB <= TC_NSB_MAX && NSB * TB_BYTES <= 144 * 1024, "TMEM / shared-memory budget");
    static_assert(PARTS == 1 || PARTS == 2 || PARTS == 4, "expander threads per row");
    const int q0 = blockIdx.x * TC_ROWS;
    if ((int)(blockIdx.x / CG) * CG * TC_ROWS >= pd.nq) return;   // the whole pair has no query row (a lone CTA of a pair still expands its half)
    const uint32_t rank = CG > 1 ? cluster_rank() : 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *sB = tc_smem;                                  // NSB stages of TN rows
    uint2 *lut = reinterpret_cast<uint2 *>(tc_smem + NSB * TB_BYTES);
    TcShared *S = reinterpret_cast<TcShared *>(tc_smem + NSB * TB_BYTES + TC_LUT_BYTES);
    const int NT = (pd.nt + TN - 1) / TN;

    if (threadIdx.x == 0) {
        // fullA / full / tempty collect arrivals from both CTAs of a pair (only the leader's copies are used then)
        mbar_init(&S->fullA, 8 * CG);
        for (int s = 0; s < NSB; s++) { mbar_init(&S->full[s], (TC_EXP_THREADS / 32) * CG); mbar_init(&S->empty[s], MB); }   // one commit per issuer
        for (int s = 0; s < TS; s++)
            for (int b = 0; b < MB; b++) { mbar_init(&S->tfull[s][b], 1); mbar_init(&S->tempty[s][b], (8 / MB) * CG); }   // the epilogue warps that read this accumulator
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // TMEM: 512 columns (accumulators + the query blocks)
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S->tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S->tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    // global loads that the roles need first are issued before the prologue's barrier (table build, TMEM allocation):
    // expanders: the descriptors of the first TC_PF train tiles; epilogue warps: their query row
    constexpr int TC_PF = 3;   // train tiles in flight in registers: one L2 round trip (~700 clk) is longer than a tile's MMAs
    const int eg = (warp - TC_EXP_WARP0) >> 2;            // expander group: tiles t = eg mod 2 (a tile's LDS -> STS chains leave the
                                                          // warp waiting most of the time; two groups keep two tiles in the works)
    const int et = threadIdx.x - 32 * TC_EXP_WARP0 - eg * TC_EXP_THREADS;
    const int r = et & (TL - 1), part = et / TL;          // PARTS = 2 / 4: two / four expander threads share a row (16 / 8 descriptor bytes each)
    const uint4 *tq = reinterpret_cast<const uint4 *>(pd.t);
    uint4 lo[TC_PF], hi[TC_PF];
#pragma unroll
    for (int i = 0; i < TC_PF; i++) lo[i] = hi[i] = make_uint4(0, 0, 0, 0);
    auto fetch = [&](int t, uint4 &l, uint4 &h) {
        const int row = t * TN + (int)rank * TL + r;      // this CTA's part of the tile
        if (t < NT && row < pd.nt) {
            if (PARTS == 1) { l = __ldg(tq + 2 * (size_t)row); h = __ldg(tq + 2 * (size_t)row + 1); }
            else if (PARTS == 2) l = __ldg(tq + 2 * (size_t)row + part);
            else { const uint2 d = __ldg(reinterpret_cast<const uint2 *>(tq) + 4 * (size_t)row + part); l.x = d.x; l.y = d.y; }
        }
    };
    const int e_lane_base = 32 * (warp & 3), e_grp = (warp - TC_EPI_WARP0) >> 2, e_mb = MB > 1 ? e_grp : 0;
    const int e_row = q0 + e_mb * TC_M + e_lane_base + lane;
    uint4 dlo = make_uint4(0, 0, 0, 0), dhi = dlo;
    if (warp >= TC_EXP_WARP0) {
#pragma unroll
        for (int i = 0; i < TC_PF; i++) fetch(eg + 2 * i, lo[i], hi[i]);
    } else if (warp >= TC_EPI_WARP0 && e_row < pd.nq) {
        const uint4 *qq = reinterpret_cast<const uint4 *>(pd.q) + 2 * (size_t)e_row;
        if (MB > 1 || e_grp == 0) dlo = __ldg(qq);       // MB = 1: the two warps of a lane quarter split the row
        if (MB > 1 || e_grp == 1) dhi = __ldg(qq + 1);
    }
    for (int i = threadIdx.x; i < 256 * 16; i += TC_THREADS) {   // byte e -> eight +-1 bytes, replica i % 16
        const uint32_t e = (uint32_t)i >> 4;
        const uint32_t b0 = ((e & 0xF) * 0x00204081u) & 0x01010101u, b1 = ((e >> 4) * 0x00204081u) & 0x01010101u;
        lut[i] = make_uint2(0x01010101u + 0xFEu * b0, 0x01010101u + 0xFEu * b1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 1) __syncthreads(); else cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S->tmem_base;
    int best = -1000, second = -1000, bidx = -1;   // running (max dot, second, arg) of an epilogue thread

    if (warp >= TC_EXP_WARP0) {
        // ---- expanders: packed descriptors (global, read-only path) -> +-1 int8 rows in shared memory
        for (int s = 0; s < NSB; s++) write_key_chunks(sB + s * TB_BYTES, r, PARTS == 1 || part == 0, PARTS == 1 || part == 1);
#pragma unroll 1
        for (int t = eg; t < NT; t += 2) {
            const int s = t % NSB, ph = (t / NSB) & 1;
            mbar_wait(&S->empty[s], ph ^ 1);
            uint8_t *tile = sB + s * TB_BYTES;
            const bool rv = t * TN + (int)rank * TL + r < pd.nt;
            if (PARTS == 1) expand_row_lut(tile, r, lo[0], hi[0], rv, lut + (lane & 15));
            else if (PARTS == 2) expand_half_lut(tile, r, part, lo[0], rv, lut + (lane & 15));
            else expand_quarter_lut(tile, r, part, make_uint2(lo[0].x, lo[0].y), rv, lut + (lane & 15));
#pragma unroll
            for (int i = 0; i + 1 < TC_PF; i++) { lo[i] = lo[i + 1]; hi[i] = hi[i + 1]; }
            fetch(t + 2 * TC_PF, lo[TC_PF - 1], hi[TC_PF - 1]);   // in flight while the next tiles are expanded
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CG>(&S->full[s]);            // one (for the peer CTA: remote) arrival per warp
        }
    } else if (warp < (MB == 1 ? 1 : 2 * MB)) {
        if (rank == 0 && elect_one()) {
            // ---- MMA issuers.  A single thread gets a tcgen05.mma out every ~120 cycles (ncu: the issuing warp never waits
            //      on a barrier; every instruction needs ~9 dependent uniform-datapath instructions of set-up), while an
            //      M128 x N x K32 int8 step runs N/2 cycles on the tensor pipe.  So the tiles are dealt to 2 * MB issuing
            //      threads: thread (block mbi, parity par) issues the nine K steps of block mbi for the tiles t = par mod 2
            //      into TMEM stage t mod TS; tiles of different threads use different accumulators or are ordered by tempty.
            // idesc = D s32 | A s8 | B s8 | K-major both | N = TN | M = 128
            // (MB = 1: one thread issues every tile — with more than two accumulator stages a stage is no longer tied to a tile
            // parity, and a parity wait is only safe for a waiter that sees every phase of its barrier)
            const int mbi = warp % MB, par = warp / MB;
            constexpr int TSTEP = MB == 1 ? 1 : 2;
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((TC_M * CG) >> 4) << 24);
            mbar_wait_cg<CG>(&S->fullA, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int t = par; t < NT; t += TSTEP) {
                const int s = t % NSB, ph = (t / NSB) & 1;      // shared-memory stage
                const int ts = t % TS, tph = (t / TS) & 1;      // TMEM stage and its phase
                mbar_wait_cg<CG>(&S->full[s], ph);
                mbar_wait_cg<CG>(&S->tempty[ts][mbi], tph ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = make_desc(smem_u32(sB + s * TB_BYTES));
#pragma unroll
                for (int k = 0; k < TC_KB / 32; k++) {   // K step = 32 int8 = 8 TMEM columns of A = two 16-byte core-matrix columns (256 B) of B
                    if (CG == 1) mma_i8_ts(tmem + (uint32_t)((ts * MB + mbi) * TN), tmem + (uint32_t)(DCOLS + mbi * TC_ACOLS + 8 * k),
                                           bdesc + (uint64_t)((k * 256) >> 4), idesc, k > 0 ? 1u : 0u);
                    else mma_i8_ts2(tmem + (uint32_t)((ts * MB + mbi) * TN), tmem + (uint32_t)(DCOLS + mbi * TC_ACOLS + 8 * k),
                                    bdesc + (uint64_t)((k * 256) >> 4), idesc, k > 0 ? 1u : 0u);
                }
                tc_commit_cg<CG>(&S->empty[s]);          // shared-memory stage reusable once the MMAs of every block retire
                tc_commit_cg<CG>(&S->tfull[ts][mbi]);   // accumulator ready for the epilogue
            }
        }
    } else if (warp >= TC_EPI_WARP0) {
        // ---- epilogue warps 4..11: TMEM lanes 32*(warp%4) .. +31; warps 4..7 own the query rows of M block 0,
        //      warps 8..11 those of M block 1 — one thread per query row, all columns
        // (MB = 1: both warps of a lane quarter share the rows and split every tile's columns; merged at the end)
        const int lane_base = 32 * (warp & 3);
        const int grp = (warp - TC_EPI_WARP0) >> 2;
        const int mb = MB > 1 ? grp : 0;
        {
            // A operand: this thread's query row, +-32 with a 1 in element 256, into TMEM lane lane_base + lane.  Eight
            // columns (= 32 K elements = one 32-bit word of the descriptor) per tcgen05.st; group 8 carries the key element.
            const int row = q0 + mb * TC_M + lane_base + lane;
            const bool valid = row < pd.nq;
            const bool first_half = MB > 1 || grp == 0, second_half = MB > 1 || grp == 1;   // (dlo, dhi were fetched in the prologue)
            auto put = [&](int g, uint32_t bits, bool key) {
                uint32_t w[8];
                if (!key) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const uint32_t b = (((bits >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;   // bit i -> byte i
                        w[k] = valid ? 0x20202020u + 0xC0u * b : 0u;                                     // 0 -> +32, 1 -> -32; no carries
                    }
                } else {
                    w[0] = valid ? 1u : 0u;
#pragma unroll
                    for (int k = 1; k < 8; k++) w[k] = 0u;
                }
                const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + (uint32_t)(DCOLS + mb * TC_ACOLS + 8 * g);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            };
            if (first_half) { put(0, dlo.x, false); put(1, dlo.y, false); put(2, dlo.z, false); put(3, dlo.w, false); }
            if (second_half) { put(4, dhi.x, false); put(5, dhi.y, false); put(6, dhi.z, false); put(7, dhi.w, false); put(8, 0u, true); }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CG>(&S->fullA);
        }
        // MB = 2: the four warps of block mb read every tile (64 columns); MB = 1: both warp groups read every tile, 64 of its 128
        // columns each (merged at the end).  An accumulator is released as soon as it sits in registers.
        for (int t = 0; t < NT; t++) {
            const int s = t % TS, ph = (t / TS) & 1;
            mbar_wait(&S->tfull[s][mb], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int col0 = t * TN;
            constexpr int nh = 1;
            {
            constexpr int half = 0;
            const int cbeg = MB == 1 ? 64 * grp : 0;
            // the warp's 64 columns of this tile with one tcgen05.ld (a load + wait per 32 columns cost ~85 cycles each)
            uint32_t vv[64];
            {
                const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + (uint32_t)((s * MB + mb) * TN + cbeg);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                    "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                    "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                    : "=r"(vv[0]), "=r"(vv[1]), "=r"(vv[2]), "=r"(vv[3]), "=r"(vv[4]), "=r"(vv[5]), "=r"(vv[6]), "=r"(vv[7]),
                      "=r"(vv[8]), "=r"(vv[9]), "=r"(vv[10]), "=r"(vv[11]), "=r"(vv[12]), "=r"(vv[13]), "=r"(vv[14]), "=r"(vv[15]),
                      "=r"(vv[16]), "=r"(vv[17]), "=r"(vv[18]), "=r"(vv[19]), "=r"(vv[20]), "=r"(vv[21]), "=r"(vv[22]), "=r"(vv[23]),
                      "=r"(vv[24]), "=r"(vv[25]), "=r"(vv[26]), "=r"(vv[27]), "=r"(vv[28]), "=r"(vv[29]), "=r"(vv[30]), "=r"(vv[31]),
                      "=r"(vv[32]), "=r"(vv[33]), "=r"(vv[34]), "=r"(vv[35]), "=r"(vv[36]), "=r"(vv[37]), "=r"(vv[38]), "=r"(vv[39]),
                      "=r"(vv[40]), "=r"(vv[41]), "=r"(vv[42]), "=r"(vv[43]), "=r"(vv[44]), "=r"(vv[45]), "=r"(vv[46]), "=r"(vv[47]),
                      "=r"(vv[48]), "=r"(vv[49]), "=r"(vv[50]), "=r"(vv[51]), "=r"(vv[52]), "=r"(vv[53]), "=r"(vv[54]), "=r"(vv[55]),
                      "=r"(vv[56]), "=r"(vv[57]), "=r"(vv[58]), "=r"(vv[59]), "=r"(vv[60]), "=r"(vv[61]), "=r"(vv[62]), "=r"(vv[63])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            if (half == nh - 1) {   // the values are in registers: the issuer may overwrite this accumulator
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_leader<CG>(&S->tempty[s][mb]);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int c = cbeg + 32 * h;
                if (col0 + c >= pd.nt) break;   // warp-uniform
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = vv[32 * h + j];
                const int lim = pd.nt - (col0 + c);   // valid columns in this chunk
                if (!SECOND && lim >= 32) {
                    // best-only, full chunk: the accumulators already are keys (dot * 32 + 31 - j: larger key = larger dot,
                    // then lower column); reduce the 32 keys with a max3 tree: ~0.4 instructions per element
                    int k[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) k[j] = (int)v[j];     // = dot * 32 + 31 - j, straight from the tensor core
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
                    // best + second, full chunk.  First the chunk maximum alone (max3 tree, ~0.4 instructions per element): a
                    // chunk whose largest dot does not exceed the row's running second best changes nothing, and after the
                    // first few hundred columns that is almost every chunk (the running second is a record statistic), so
                    // the pair tree below only runs when some row of the warp needs it.
                    {
                        int k[11];
#pragma unroll
                        for (int j = 0; j < 30; j += 3) k[j / 3] = __vimax3_s32((int)v[j], (int)v[j + 1], (int)v[j + 2]);
                        k[10] = max((int)v[30], (int)v[31]);
                        const int a0 = __vimax3_s32(k[0], k[1], k[2]), a1 = __vimax3_s32(k[3], k[4], k[5]), a2 = __vimax3_s32(k[6], k[7], k[8]);
                        const int ck = max(__vimax3_s32(a0, a1, a2), max(k[9], k[10]));
                        if (!__any_sync(0xffffffffu, (ck >> 5) > second)) continue;
                    }
                    // same keys, reduced by a tree of (largest, second largest) pairs — 2 ops per leaf pair, 3 per merge
                    // (max, min, max3): ~2.4 ALU instructions per element instead of ~6
                    int hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int x = (int)v[2 * j], y = (int)v[2 * j + 1];   // keys dot * 32 + 31 - column
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
                    const int d = (int)v[j] >> 5;                        // key -> dot (arithmetic shift = floor, the key's low part is 0..31)
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
            }
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
    if (CG == 1) __syncthreads(); else cluster_sync_all();   // nobody leaves while the peer may still read this CTA's tile halves or signal its barriers
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
    if (MB == 1 && warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + 4) {
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
void tb_match_tc_free(tb_ctx *) {}

// CTA pairs (cta_group::2) are the default for the kNN-2 problem (one M block per CTA: halving the expansion per SM is what
// it needs; 200k x 200k 7.5 -> 6.9 ms); the batched frame-pair matcher already shares every tile between two M blocks of one
// CTA and is faster that way (0.40 ms against 0.51 ms as pairs).  TB_MATCH_CG = 1 / 2 forces one form for both (A/B).
static int tc_cta_group(bool knn2)
{
    static const int cg = getenv("TB_MATCH_CG") ? atoi(getenv("TB_MATCH_CG")) : 0;
    return cg == 1 ? 1 : cg == 2 ? 2 : knn2 ? 2 : 1;
}
template <typename K, typename... Args>
static cudaError_t launch_pairs(K kernel, dim3 grid, size_t smem, cudaStream_t st, Args... args)
{
    grid.x = (grid.x + 1) & ~1u;                                  // whole pairs; a CTA without query rows still expands its half
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}   // the tensor-core matcher keeps no workspace of its own any more (operands are expanded in the kernel)

// Fills best_idx/best_dist[2P][kp_cap] exactly like hamming_knn<false> does for the same pair table (match.cu: build_pairs).
int tb_match_knn_tc(tb_ctx *ctx, int n_pairs, const PairDesc *d_pairs, int kp_cap, int *best_idx, int *best_dist)
{
    ProfScope prof__(ctx, "hamming_i8mma");
    dim3 grid(div_up(kp_cap, TC_M * 2), 2 * n_pairs);
    if (tc_cta_group(false) == 2)   // pairs: one M block per CTA, 128-row tiles, three accumulator stages
        launch_pairs(hamming_i8mma<false, 1, 2>, dim3(div_up(kp_cap, TC_M), 2 * n_pairs), TC_SMEM_BYTES, ctx->stream, d_pairs, kp_cap, best_idx, best_dist, (int *)nullptr);
    else hamming_i8mma<false, 2, 1><<<grid, TC_THREADS, TC_SMEM_BYTES, ctx->stream>>>(d_pairs, kp_cap, best_idx, best_dist, nullptr);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

// best / second / argbest of nq query rows against nt train rows (tb_knn2_device with the tensor-core matcher);
// d_pairs = n_splits PairDesc {q, t + split offset, nq, rows of the split} already on the device: split s writes its partial
// result to idx / best / second + s * nq (n_splits == 1: the final arrays)
int tb_knn2_tc(tb_ctx *ctx, const PairDesc *d_pairs, int nq, int n_splits, int *idx, int *best, int *second)
{
    ProfScope prof__(ctx, "hamming_i8mma");
    if (tc_cta_group(true) == 2) launch_pairs(hamming_i8mma<true, 1, 2>, dim3(div_up(nq, TC_M), n_splits), TC_SMEM_BYTES, ctx->stream, d_pairs, nq, idx, best, second);
    else hamming_i8mma<true, 1, 1><<<dim3(div_up(nq, TC_M), n_splits), TC_THREADS, TC_SMEM_BYTES, ctx->stream>>>(d_pairs, nq, idx, best, second);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}
// How many slices of the train set make the launch fill whole waves of one 128-row CTA per SM: 25 000 query rows (a rank's
// share of config 4 on 8 GPUs) are 196 CTAs = 1.3 waves on 148 SMs; three slices make it 3.97 waves of a third of the length.
int tb_knn2_tc_splits(int nq, int nt, int num_sms)
{
    const int ctas = div_up(nq, TC_M);
    int best_s = 1;
    double best_cost = (double)div_up(ctas, num_sms);
    for (int s = 2; s <= 8; s++) {
        if (nt / s < 4096) break;
        const double cost = (double)div_up(ctas * s, num_sms) / s;
        if (cost < best_cost * 0.9) { best_cost = cost; best_s = s; }   // a slice more costs a query expansion per CTA and the merge
    }
    return best_s;
}
__global__ void knn2_merge_splits(const int *__restrict__ p_idx, const int *__restrict__ p_best, const int *__restrict__ p_second, int nq, int n_splits,
                                  int rows_per_split, int *__restrict__ idx, int *__restrict__ best, int *__restrict__ second)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nq) return;
    int bi = -1, bd = 0x7fffffff, sd = 0x7fffffff;
    for (int s = 0; s < n_splits; s++) {                          // slices in index order: the earlier one wins ties
        const int i = p_idx[(size_t)s * nq + r], d = p_best[(size_t)s * nq + r], d2 = p_second[(size_t)s * nq + r];
        if (i < 0) continue;
        if (d < bd) { sd = min(bd, d2); bd = d; bi = i + s * rows_per_split; }
        else sd = min(sd, d);
    }
    idx[r] = bi; best[r] = bd; second[r] = sd;
}
int tb_knn2_tc_merge(tb_ctx *ctx, const int *p_idx, const int *p_best, const int *p_second, int nq, int n_splits, int rows_per_split,
                     int *idx, int *best, int *second)
{
    knn2_merge_splits<<<div_up(nq, 256), 256, 0, ctx->stream>>>(p_idx, p_best, p_second, nq, n_splits, rows_per_split, idx, best, second);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

void tb_match_tc_init_device()
{
    const int smem = (int)(TC_SMEM_BYTES);
    cudaFuncSetAttribute(hamming_i8mma<false, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(hamming_i8mma<true, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(hamming_i8mma<false, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(hamming_i8mma<true, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}
