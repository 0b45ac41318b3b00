// octree.cu — K3: quadtree keypoint distribution, one CTA per (frame, level).
// Replaces ORBExtractor::DistributeOctTree + ExtractorNode::DivideNode
// (reference src/extractors/ORBextractor.cpp:494-733, :416-491).
//
// The reference walks a std::list of nodes that own std::vectors of keypoints.  Restated for a CTA:
//  * a node is its bounding box + key count; the list is an ARRAY kept in list order;
//  * keys never move: every key carries the index of its node (u16) and re-derives its child from
//    the box midpoints (halfX = ceil(w/2), float compares of integer coordinates => exact in int);
//  * one "pass" = (A) all keys of nodes with >1 key vote their child into 4 counters per node,
//    (B) nodes decide who splits and compute their new list positions with block-wide prefix sums
//    (children are push_front'ed in n1..n4 order, parents erased in place, so
//    new list = reverse(children in processing order) ++ untouched nodes in old order),
//    (C) keys remap to the new node indices.
//  * phase 1 (breadth-first, :555-624) processes in list order; phase 2 (:631-695) processes nodes
//    sorted by (size, UL.x, UL.y) descending (the pinned comparator) and stops at the
//    first node where the list already holds >= N entries — a prefix-sum cutoff.
//  * the per-leaf winner (max response, first in vKeys order on ties, :703-716) is an atomicMax of
//    (m, ~order) where order = (cell row, cell col, y, x) is the reference's candidate order.
//  * exit keys (AddPoints, :526-529, :467-480, :718-729) ride along as a second key set.
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>

#ifndef OCT_THREADS
#define OCT_THREADS 512
#endif
#ifdef OCT_MINB
#define OCT_BOUNDS __launch_bounds__(OCT_THREADS, OCT_MINB)
#else
#define OCT_BOUNDS __launch_bounds__(OCT_THREADS)
#endif


struct OctSmem {
    short4 *box[2];   // {ulx, urx, uly, bry}
    int *cnt[2];
    int *child;       // [node_cap*4] child votes of the current pass, turned into the children's new list positions
    int *child2;      // [node_cap*4] votes for the NEXT pass, cast while the keys are remapped (the two swap every pass)
    int *aux;         // [node_cap] per-node scratch (scan input / remap of unsplit nodes)
    int *aux2;        // [node_cap] nchild / split flags
    unsigned long long *skey; // [pow2(node_cap)] sort keys / best keys
};

// v / d for 0 <= v < 2^16 with magic = 2^32 / d rounded up (0 for d == 1)
__device__ __forceinline__ int oct_div(int v, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)v, magic) : v; }
// keypoint coordinates are integers (the reference compares them as floats: same result, without four conversions per key)
__device__ __forceinline__ int oct_child_i(short4 b, int x, int y)
{
    const int mx = b.x + ((b.y - b.x + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
    const int my = b.z + ((b.w - b.z + 1) >> 1);   // UL.y + ceil((BR.y-UL.y)/2)
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}
__device__ __forceinline__ int oct_child(short4 b, float x, float y)
{
    const int mx = b.x + ((b.y - b.x + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
    const int my = b.z + ((b.w - b.z + 1) >> 1);   // UL.y + ceil((BR.y-UL.y)/2)
    return (x < (float)mx ? 0 : 1) + (y < (float)my ? 0 : 2);
}
__device__ __forceinline__ short4 oct_child_box(short4 b, int c)
{
    const int mx = b.x + ((b.y - b.x + 1) >> 1);
    const int my = b.z + ((b.w - b.z + 1) >> 1);
    short4 r;
    r.x = (c & 1) ? mx : b.x; r.y = (c & 1) ? b.y : mx;
    r.z = (c & 2) ? my : b.z; r.w = (c & 2) ? b.w : my;
    return r;
}

// The key sweeps are bound by the latency of cand[k] (L2, one dependent load per iteration — ncu: 35 % of the stall samples sit
// on its first use): every sweep fetches OCT_MLP candidates per thread before it touches the first one.
#ifndef OCT_MLP
#define OCT_MLP 4
#endif

// In-place exclusive scan of a[0..n) (shared memory); returns the total. All threads must call.
__device__ int block_excl_scan(int *a, int n)
{
    __shared__ int s_warp[OCT_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int ipt = (n + OCT_THREADS - 1) / OCT_THREADS;
    const int b0 = tid * ipt, b1 = min(b0 + ipt, n);
    int sum = 0;
    for (int i = b0; i < b1; i++) sum += a[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        int w = tid < OCT_THREADS / 32 ? s_warp[tid] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (tid >= o) winc += t;
        }
        if (tid < OCT_THREADS / 32) s_warp[tid] = winc - w;
        if (tid == OCT_THREADS / 32 - 1) s_total = winc;
    }
    __syncthreads();
    int run = s_warp[tid >> 5] + inc - sum;
    for (int i = b0; i < b1; i++) { int v = a[i]; a[i] = run; run += v; }
    __syncthreads();
    return s_total;
}

// Descending bitonic sort of key[0..m), m a power of two.
__device__ void block_bitonic_desc(unsigned long long *key, int m)
{
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += OCT_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = key[i], b = key[ixj];
                    bool up = (i & k) == 0; // "up" blocks sorted descending
                    if (up ? (a < b) : (a > b)) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

extern __shared__ __align__(16) uint8_t oct_smem_raw[];

struct OctArgsAll { OctArgs lv[TB_MAX_LEVELS]; int nid_cache_keys; };

// grid = (frames, levels): all (frame, level) problems of a batch in ONE launch, so the short per-level
// tails overlap instead of serialising eight launches that are each bounded by their slowest CTA.
// NIDS: the per-key node indices of this (frame, level) fit the launch's shared-memory key cache.  Two instantiations so that
// the compiler knows the address space of nid[] in each (one pointer that is shared or global at run time is a generic
// pointer: 64-bit address arithmetic and generic loads in every sweep).
template <bool NIDS>
__device__ __forceinline__ void octree_body(const OctArgsAll &all)
{
    const OctArgs &a = all.lv[blockIdx.y];
    const int tid = threadIdx.x;
    const int f = blockIdx.x;
    const int NC = a.node_cap;
    int n = a.n_cand[(size_t)f * a.n_stride];
    if (n > a.cand_cap) n = a.cand_cap;
    const uint32_t *cand = a.cand + (size_t)f * a.cand_cap;
    uint16_t *nid = a.nid + (size_t)f * a.cand_cap;   // per-key node index; lives in shared memory when the keys fit (below)
    int16_t *enid = a.n_exit ? a.enid + (size_t)f * a.n_exit : nullptr;
    const int NE = a.n_exit;

    // carve shared memory
    OctSmem S;
    {
        uint8_t *p = oct_smem_raw;
        int m2 = 1; while (m2 < NC) m2 <<= 1;
        S.skey = reinterpret_cast<unsigned long long *>(p); p += sizeof(unsigned long long) * m2;
        S.box[0] = reinterpret_cast<short4 *>(p); p += sizeof(short4) * NC;
        S.box[1] = reinterpret_cast<short4 *>(p); p += sizeof(short4) * NC;
        S.cnt[0] = reinterpret_cast<int *>(p); p += sizeof(int) * NC;
        S.cnt[1] = reinterpret_cast<int *>(p); p += sizeof(int) * NC;
        S.child = reinterpret_cast<int *>(p); p += sizeof(int) * NC * 4;
        S.child2 = reinterpret_cast<int *>(p); p += sizeof(int) * NC * 4;
        S.aux = reinterpret_cast<int *>(p); p += sizeof(int) * NC;
        S.aux2 = reinterpret_cast<int *>(p); p += sizeof(int) * NC;
        // every pass reads and rewrites nid[k] of all keys (27 k at level 0 of a 1080p frame): keep it in shared memory
        // instead of L2 when it fits the launch's key cache
        if (NIDS) nid = reinterpret_cast<uint16_t *>(p);
    }
    __shared__ int s_L, s_exp, s_cut;

    if (n == 0 || a.nIni <= 0) {
        if (tid == 0) a.n_kept[(size_t)f * a.n_stride] = 0;
        return;
    }

    int cur = 0;
    // ---- initial nodes (:504-531)
    for (int i = tid; i < a.nIni; i += OCT_THREADS) {
        short4 b;
        b.x = (short)(int)__fmul_rn(a.hX, (float)i);
        b.y = (short)(int)__fmul_rn(a.hX, (float)(i + 1));
        b.z = 0; b.w = (short)a.H;
        S.box[0][i] = b; S.cnt[0][i] = 0;
    }
    __syncthreads();
    for (int k0 = tid; k0 < n; k0 += OCT_MLP * OCT_THREADS) {
        uint32_t cc[OCT_MLP];
#pragma unroll
        for (int j = 0; j < OCT_MLP; j++) { const int k = k0 + j * OCT_THREADS; cc[j] = k < n ? __ldg(&cand[k]) : 0u; }
#pragma unroll
        for (int j = 0; j < OCT_MLP; j++) {
            const int k = k0 + j * OCT_THREADS;
            if (k >= n) break;
            int ni = (int)__fdiv_rn((float)cand_x(cc[j]), a.hX);
            ni = min(ni, a.nIni - 1); // (the reference would index out of bounds; cannot happen for x < W)
            nid[k] = (uint16_t)ni;
            atomicAdd(&S.cnt[0][ni], 1);
        }
    }
    for (int e = tid; e < NE; e += OCT_THREADS) enid[e] = 0;
    __syncthreads();
    // drop empty initial nodes (:533-545): compaction in list order
    {
        for (int i = tid; i < a.nIni; i += OCT_THREADS) S.aux[i] = S.cnt[0][i] > 0;
        __syncthreads();
        int L = block_excl_scan(S.aux, a.nIni);
        for (int i = tid; i < a.nIni; i += OCT_THREADS)
            if (S.cnt[0][i] > 0) { S.box[1][S.aux[i]] = S.box[0][i]; S.cnt[1][S.aux[i]] = S.cnt[0][i]; }
        __syncthreads();
        for (int k = tid; k < n; k += OCT_THREADS) nid[k] = (uint16_t)S.aux[nid[k]];
        for (int e = tid; e < NE; e += OCT_THREADS) enid[e] = S.cnt[0][0] > 0 ? (int16_t)S.aux[0] : (int16_t)-1;
        cur = 1;
        if (tid == 0) s_L = L;
        __syncthreads();
    }

    bool phase2 = false;
    int *V = S.child, *W = S.child2;     // votes of this pass / of the next one
    // ---- (A) child votes of every node with >1 key, first pass only: later passes get theirs from the remap loop (C) of
    //      the pass before, which already has the key's coordinates in hand (one sweep over the keys per pass, not two)
    {
        const int L0 = s_L;
        const short4 *box = cur ? S.box[1] : S.box[0];   // constant indices: a dynamically indexed pointer pair puts the whole
        const int *cnt = cur ? S.cnt[1] : S.cnt[0];      // struct in local memory and every table access becomes a generic load

        for (int i = tid; i < 4 * L0; i += OCT_THREADS) V[i] = 0;
        __syncthreads();
        for (int k0 = tid; k0 < n; k0 += OCT_MLP * OCT_THREADS) {
            uint32_t cc[OCT_MLP];
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) { const int k = k0 + j * OCT_THREADS; cc[j] = k < n ? __ldg(&cand[k]) : 0u; }
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) {
                const int k = k0 + j * OCT_THREADS;
                const bool valid = k < n;
                const int ni = valid ? (int)nid[k] : 0;
                const int ch = oct_child_i(box[ni], cand_x(cc[j]), cand_y(cc[j]));
                if (valid && cnt[ni] > 1) atomicAdd(&V[4 * ni + ch], 1);
            }
        }
        __syncthreads();
    }
    for (;;) {
        const int L = s_L;
        short4 *box = cur ? S.box[1] : S.box[0], *nbox = cur ? S.box[0] : S.box[1];
        int *cnt = cur ? S.cnt[1] : S.cnt[0], *ncnt = cur ? S.cnt[0] : S.cnt[1];
        // the node tables live in shared memory: say so, or every access is a generic load with 64-bit address arithmetic
        // (the pointers come out of swaps and an indexed pair, which hides their address space from the compiler)
        // ---- (B) who splits, in which processing order
        // aux2[i] = number of non-empty children (0 for nodes that cannot split)
        for (int i = tid; i < L; i += OCT_THREADS) {
            int nch = 0;
            if (cnt[i] > 1) nch = (V[4 * i] > 0) + (V[4 * i + 1] > 0) + (V[4 * i + 2] > 0) + (V[4 * i + 3] > 0);
            S.aux2[i] = nch;
        }
        __syncthreads();
        int TC;           // total children created this pass
        if (!phase2) {
            // processing order = list order; P_i = exclusive scan of nchild
            for (int i = tid; i < L; i += OCT_THREADS) S.aux[i] = S.aux2[i];
            __syncthreads();
            TC = block_excl_scan(S.aux, L);
            // child positions: push order o among non-empty children; final = TC - P_i - 1 - o
            for (int i = tid; i < L; i += OCT_THREADS) {
                if (S.aux2[i] > 0) {
                    int P = S.aux[i], o = 0;
                    for (int c = 0; c < 4; c++) {
                        int cc = V[4 * i + c];
                        if (cc > 0) {
                            int pos = TC - P - 1 - o; o++;
                            nbox[pos] = oct_child_box(box[i], c); ncnt[pos] = cc;
                            V[4 * i + c] = pos;
                        } else V[4 * i + c] = -1;
                    }
                }
            }
            __syncthreads();
        } else {
            // processing order = (size, UL.x, UL.y) descending among nodes with >1 key
            int m2 = 1; while (m2 < L) m2 <<= 1;
            for (int i = tid; i < m2; i += OCT_THREADS) {
                unsigned long long key = 0;
                if (i < L && cnt[i] > 1)
                    key = ((unsigned long long)cnt[i] << 40) | ((unsigned long long)(unsigned short)box[i].x << 26) |
                          ((unsigned long long)(unsigned short)box[i].z << 13) | (unsigned long long)i;
                S.skey[i] = key;
            }
            __syncthreads();
            block_bitonic_desc(S.skey, m2);
            // growth of the list if sorted node j is split: nchild - 1; it is split iff the list size
            // before it is still < N (:690-691 break after the split that reaches N)
            for (int j = tid; j < L; j += OCT_THREADS) {
                unsigned long long key = S.skey[j];
                S.aux[j] = key ? S.aux2[(int)(key & 0x1FFF)] - 1 : 0;
            }
            __syncthreads();
            block_excl_scan(S.aux, L);
            if (tid == 0) s_cut = 0;
            __syncthreads();
            for (int j = tid; j < L; j += OCT_THREADS)
                if (S.skey[j] && L + S.aux[j] < a.N) atomicMax(&s_cut, j + 1);
            __syncthreads();
            const int ncut = s_cut; // sorted nodes [0, ncut) are split
            // P over the processing order
            for (int j = tid; j < L; j += OCT_THREADS) S.aux[j] = j < ncut ? S.aux2[(int)(S.skey[j] & 0x1FFF)] : 0;
            __syncthreads();
            TC = block_excl_scan(S.aux, L);
            for (int j = tid; j < ncut; j += OCT_THREADS) {
                int i = (int)(S.skey[j] & 0x1FFF);
                int P = S.aux[j], o = 0;
                for (int c = 0; c < 4; c++) {
                    int cc = V[4 * i + c];
                    if (cc > 0) {
                        int pos = TC - P - 1 - o; o++;
                        nbox[pos] = oct_child_box(box[i], c); ncnt[pos] = cc;
                        V[4 * i + c] = pos;
                    } else V[4 * i + c] = -1;
                }
            }
            __syncthreads();
            // nodes beyond the cut do not split: clear their nchild
            for (int j = ncut + tid; j < L; j += OCT_THREADS)
                if (S.skey[j]) S.aux2[(int)(S.skey[j] & 0x1FFF)] = 0;
            __syncthreads();
        }
        // untouched nodes keep their relative order behind the new children
        for (int i = tid; i < L; i += OCT_THREADS) S.aux[i] = S.aux2[i] > 0 ? 0 : 1;
        __syncthreads();
        const int nun = block_excl_scan(S.aux, L);
        const int newL = TC + nun;
        if (newL > NC) { if (tid == 0) { atomicExch(a.overflow, 2); a.n_kept[(size_t)f * a.n_stride] = 0; } return; }
        for (int i = tid; i < L; i += OCT_THREADS)
            if (S.aux2[i] == 0) { int pos = TC + S.aux[i]; nbox[pos] = box[i]; ncnt[pos] = cnt[i]; S.aux[i] = pos; }
        __syncthreads();
        // ---- (C) remap keys, and vote for the next pass in the same sweep
        for (int i = tid; i < 4 * newL; i += OCT_THREADS) W[i] = 0;
        __syncthreads();
        for (int k0 = tid; k0 < n; k0 += OCT_MLP * OCT_THREADS) {
            uint32_t cc[OCT_MLP];
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) { const int k = k0 + j * OCT_THREADS; cc[j] = k < n ? __ldg(&cand[k]) : 0u; }
            // branch-free and in stages (all loads of a stage for the OCT_MLP keys, then the next stage), so that the dependent
            // chains (nid -> aux2/box -> V -> ncnt/nbox -> vote: shared-memory round trips) overlap instead of running one after
            // the other; stores and votes come last because the compiler must assume they alias the tables. Keys past the end
            // read node 0 and store nothing.
            int ni[OCT_MLP], nn[OCT_MLP], px[OCT_MLP], py[OCT_MLP], a2[OCT_MLP], nc[OCT_MLP];
            short4 bx[OCT_MLP];
            bool valid[OCT_MLP];
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) {
                const int k = k0 + j * OCT_THREADS;
                valid[j] = k < n;
                ni[j] = valid[j] ? (int)nid[k] : 0;
                px[j] = cand_x(cc[j]); py[j] = cand_y(cc[j]);
            }
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) { a2[j] = S.aux2[ni[j]]; bx[j] = box[ni[j]]; }
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) {
                const int *src = a2[j] > 0 ? &V[4 * ni[j] + oct_child_i(bx[j], px[j], py[j])] : &S.aux[ni[j]];
                nn[j] = max(*src, 0);                 // (only a key past the end can meet an empty child's -1)
            }
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) { nc[j] = ncnt[nn[j]]; bx[j] = nbox[nn[j]]; }
#pragma unroll
            for (int j = 0; j < OCT_MLP; j++) {
                const int k = k0 + j * OCT_THREADS;
                if (valid[j]) nid[k] = (uint16_t)nn[j];
                if (valid[j] && nc[j] > 1) atomicAdd(&W[4 * nn[j] + oct_child_i(bx[j], px[j], py[j])], 1);
            }
        }
        for (int e = tid; e < NE; e += OCT_THREADS) {
            int ni = enid[e];
            if (ni >= 0) {
                if (S.aux2[ni] > 0) {
                    float2 p = a.exit_xy[e];
                    enid[e] = (int16_t)V[4 * ni + oct_child(box[ni], p.x, p.y)]; // -1 if that child is empty
                } else enid[e] = (int16_t)S.aux[ni];
            }
        }
        // expandable nodes for the next pass = new children with > 1 key
        if (tid == 0) s_exp = 0;
        __syncthreads();
        {
            int loc = 0;
            for (int i = tid; i < TC; i += OCT_THREADS) loc += ncnt[i] > 1;
            if (loc) atomicAdd(&s_exp, loc);
        }
        __syncthreads();
        cur ^= 1;
        { int *t = V; V = W; W = t; }
        const int nExp = s_exp;
        if (tid == 0) s_L = newL;
        __syncthreads();
        // ---- termination (:626-629, :693-694)
        if (newL >= a.N || newL == L) break;
        if (!phase2 && newL + 3 * nExp > a.N) phase2 = true;
    }

    // ---- best key per leaf (:700-716): max response, first in vKeys order
    const int L = s_L;
    const uint32_t magic_h = a.hCell > 1 ? 0xFFFFFFFFu / (uint32_t)a.hCell + 1u : 0u, magic_w = a.wCell > 1 ? 0xFFFFFFFFu / (uint32_t)a.wCell + 1u : 0u;
    for (int i = tid; i < L; i += OCT_THREADS) { S.skey[i] = 0; S.aux2[i] = 1; }
    __syncthreads();
    for (int k0 = tid; k0 < n; k0 += OCT_MLP * OCT_THREADS) {
        uint32_t cc[OCT_MLP];
#pragma unroll
        for (int j = 0; j < OCT_MLP; j++) { const int k = k0 + j * OCT_THREADS; cc[j] = k < n ? __ldg(&cand[k]) : 0u; }
#pragma unroll
        for (int j = 0; j < OCT_MLP; j++) {
            const int k = k0 + j * OCT_THREADS;
            if (k >= n) break;
            const uint32_t c = cc[j];
            int x = cand_x(c), y = cand_y(c);
            // (y - 3) / hCell and (x - 3) / wCell by multiply-high (exact: the dividends are below 4096, the cells a few dozen pixels)
            unsigned long long ord = ((unsigned long long)oct_div(y - 3, magic_h) << 32) | ((unsigned long long)oct_div(x - 3, magic_w) << 24) |
                                     ((unsigned long long)y << 12) | (unsigned long long)x;
            unsigned long long key = ((unsigned long long)cand_m(c) << 40) | (~ord & 0xFFFFFFFFFFull);
            atomicMax(&S.skey[nid[k]], key);
        }
    }
    __syncthreads();
    // exit-key exclusion (:718-729): 20-px radius around the winner, per leaf
    for (int e = tid; e < NE; e += OCT_THREADS) {
        int ni = enid[e];
        if (ni >= 0) {
            unsigned long long ord = ~S.skey[ni] & 0xFFFFFFFFFFull;
            float bx = (float)(int)(ord & 0xFFF), by = (float)(int)((ord >> 12) & 0xFFF);
            float2 p = a.exit_xy[e];
            float dx = __fsub_rn(p.x, bx), dy = __fsub_rn(p.y, by);
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < 400.f) S.aux2[ni] = 0;
        }
    }
    __syncthreads();
    for (int i = tid; i < L; i += OCT_THREADS) S.aux[i] = S.aux2[i];
    __syncthreads();
    const int nk = block_excl_scan(S.aux, L);
    uint32_t *kept = a.kept + (size_t)f * a.kept_cap;
    for (int i = tid; i < L; i += OCT_THREADS) {
        if (S.aux2[i]) {
            unsigned long long key = S.skey[i];
            unsigned long long ord = ~key & 0xFFFFFFFFFFull;
            int pos = S.aux[i];
            if (pos < a.kept_cap) kept[pos] = cand_pack((int)(ord & 0xFFF), (int)((ord >> 12) & 0xFFF), (int)(key >> 40));
        }
    }
    if (tid == 0) {
        a.n_kept[(size_t)f * a.n_stride] = min(nk, a.kept_cap);
        if (nk > a.kept_cap) atomicExch(a.overflow, 3);
    }
}

__global__ void OCT_BOUNDS
octree_distribute(const __grid_constant__ OctArgsAll all)
{
    const OctArgs &a = all.lv[blockIdx.y];
    int n = a.n_cand[(size_t)blockIdx.x * a.n_stride];
    if (n > a.cand_cap) n = a.cand_cap;
    if (n <= all.nid_cache_keys) octree_body<true>(all);
    else octree_body<false>(all);
}

size_t tb_octree_smem_bytes(int node_cap)
{
    int m2 = 1; while (m2 < node_cap) m2 <<= 1;
    return sizeof(unsigned long long) * m2 + (size_t)node_cap * (2 * sizeof(short4) + 2 * sizeof(int) + 8 * sizeof(int) + 2 * sizeof(int));
}

int tb_launch_octree(tb_ctx *ctx, const OctArgs *levels, int n_levels, int n_frames, cudaStream_t stream)
{
    OctArgsAll all;
    memset(&all, 0, sizeof all);
    int max_cap = 0;
    for (int l = 0; l < n_levels; l++) { all.lv[l] = levels[l]; if (levels[l].node_cap > max_cap) max_cap = levels[l].node_cap; }
    for (int l = 0; l < n_levels; l++) all.lv[l].node_cap = max_cap;   // one shared-memory carve-up for the launch
    size_t smem = tb_octree_smem_bytes(max_cap);
    if (smem > 200 * 1024 || max_cap > 8191) return tb_fail(ctx, TB_E_UNSUPPORTED, "octree: per-level budget too large for shared memory");
    // key cache: whatever keeps two CTAs per SM (the register file allows two), at most 32 Ki keys
    {
        int max_keys = 0;
        for (int l = 0; l < n_levels; l++) max_keys = std::max(max_keys, levels[l].cand_cap);
        const size_t room = smem < 108 * 1024 ? 108 * 1024 - smem : 0;
        all.nid_cache_keys = (int)std::min<size_t>(std::min<size_t>(room / sizeof(uint16_t), 32768), (size_t)max_keys);
        all.nid_cache_keys &= ~7;
        smem += (size_t)all.nid_cache_keys * sizeof(uint16_t);
    }
    ProfScope prof__(ctx, "octree_distribute");
    octree_distribute<<<dim3(n_frames, n_levels), OCT_THREADS, smem, stream>>>(all);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

void tb_octree_init_device() { cudaFuncSetAttribute(octree_distribute, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }
