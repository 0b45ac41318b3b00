// window.cu — K9/K10: search grid, windowed / projection-guided Hamming matching and the
// rotation-histogram consistency filter.  Replaces, in the reference:
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/types/Frame.cpp:187-265 (+ :30-31)
//   Matcher::searchByViolence                                     src/matchers/matcher.cpp:299-395
//   Matcher::searchByProjection(F1,F2)                            src/matchers/matcher.cpp:405-531
//   Matcher::searchByProjection(map,F1,r) + Frame::IsInFrustum    src/matchers/matcher.cpp:539-617, Frame.cpp:370-412
//   rotation histogram + ComputeThreeMaxima                       src/matchers/matcher.cpp:359-393, :493-529, :810-851
//   PinholeCamera::World2Cam / CameraModel::IsInFrame             src/camera/CameraModel.cpp:63-94, CameraModel.h:33-39
//   Frame::SetBow (DBoW2 vocabulary transform) + Matcher::searchByBow  src/types/Frame.cpp:267-270, matcher.cpp:619-721
//
// Grid = CSR over the reference's 120x36 cells (cell (ix,iy) at ix*36+iy, items in insertion order),
// so the candidates of one ix column and iy range are ONE contiguous item range, already in the
// reference's traversal order.  One warp per query: lanes take consecutive candidates, keep their
// two smallest (distance, traversal position) keys, and a shuffle tree merges the 32 pairs — this
// reproduces "first strict minimum in traversal order" and the reference's second-best exactly.
// All parity-critical float math uses explicit round-to-nearest intrinsics (no FMA contraction),
// mirroring the reference's unoptimised build.
#include "common.cuh"
#include "kernels.cuh"

#define GRID_COLS 120
#define GRID_ROWS 36
#define GRID_CELLS (GRID_COLS * GRID_ROWS)

struct GridDev {
    const int *cell_start; // [GRID_CELLS+1]
    const int *items;
    const tb_keypoint *kps;
    float wInv, hInv;      // mfGridElementWidthInv = 36/rows, mfGridElementHeightInv = 120/cols (sic, Frame.cpp:30-31)
};

// ---- grid build
__global__ void grid_cell_ids(const tb_keypoint *kps, int n, float wInv, float hInv, int *cellid, int *count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int px = (int)roundf(__fmul_rn(kps[i].x, wInv)), py = (int)roundf(__fmul_rn(kps[i].y, hInv)); // Frame.cpp:259-260
    int c = -1;
    if (!(px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS)) { c = px * GRID_ROWS + py; atomicAdd(&count[c], 1); }
    cellid[i] = c;
}
__global__ void __launch_bounds__(1024) grid_scan(const int *count, int *cell_start)
{
    __shared__ int s[1024];
    const int tid = threadIdx.x;
    const int ipt = (GRID_CELLS + 1023) / 1024;
    int b0 = min(tid * ipt, GRID_CELLS), b1 = min(b0 + ipt, GRID_CELLS);
    int sum = 0;
    for (int i = b0; i < b1; i++) sum += count[i];
    s[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { int v = tid >= o ? s[tid - o] : 0; __syncthreads(); s[tid] += v; __syncthreads(); }
    int run = s[tid] - sum;
    for (int i = b0; i < b1; i++) { cell_start[i] = run; run += count[i]; }
    if (tid == 1023) cell_start[GRID_CELLS] = s[1023];
}
__global__ void grid_fill(const int *cellid, int n, const int *count, const int *cell_start, int *items)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= GRID_CELLS || count[c] == 0) return;
    int pos = cell_start[c];
    for (int i = 0; i < n; i++) if (cellid[i] == c) items[pos++] = i; // insertion (index) order
}

// ---- candidate walk
struct Top2 { uint32_t k1, k2; int i1, i2; }; // keys = dist<<20 | traversal position; 0xFFFFFFFF = none
__device__ __forceinline__ void top2_push(Top2 &t, uint32_t key, int idx)
{
    if (key < t.k1) { t.k2 = t.k1; t.i2 = t.i1; t.k1 = key; t.i1 = idx; }
    else if (key < t.k2) { t.k2 = key; t.i2 = idx; }
}
__device__ __forceinline__ Top2 top2_warp_merge(Top2 t)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t ok1 = __shfl_xor_sync(0xffffffffu, t.k1, o), ok2 = __shfl_xor_sync(0xffffffffu, t.k2, o);
        int oi1 = __shfl_xor_sync(0xffffffffu, t.i1, o), oi2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
        top2_push(t, ok1, oi1);
        top2_push(t, ok2, oi2);
    }
    return t;
}
__device__ __forceinline__ int ham256_g(const uint32_t q[8], const uint8_t *d)
{
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(d)), b = __ldg(reinterpret_cast<const uint4 *>(d) + 1);
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
           __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

// GetFeaturesInArea + distance scan for one warp. skip[j] != 0 excludes candidate j (may be null).
// If out_idx != null the surviving candidate indices are also written in traversal order (lane-
// cooperative, used by tb_features_in_area) and *out_n receives their number.
__device__ Top2 window_scan(const GridDev &g, float x, float y, float r, int minLevel, int maxLevel,
                            const uint32_t q[8], const uint8_t *desc, const uint8_t *skip, int *out_idx, int *out_n)
{
    const int lane = threadIdx.x & 31;
    Top2 t = {0xFFFFFFFFu, 0xFFFFFFFFu, -1, -1};
    int nout = 0;
    const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(x, r), g.wInv)));
    const int x1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(x, r), g.wInv)));
    const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(y, r), g.hInv)));
    const int y1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(y, r), g.hInv)));
    if (x0 >= GRID_COLS || x1 < 0 || y0 >= GRID_ROWS || y1 < 0) { if (out_n) *out_n = 0; return t; }
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    int tpos = 0;
    for (int ix = x0; ix <= x1; ix++) {
        const int s = g.cell_start[ix * GRID_ROWS + y0], e = g.cell_start[ix * GRID_ROWS + y1 + 1];
        for (int k0 = s; k0 < e; k0 += 32) {
            const int k = k0 + lane;
            bool pass = false;
            int j = -1;
            if (k < e) {
                j = g.items[k];
                const tb_keypoint kp = g.kps[j];
                pass = true;
                if (check) {
                    if (kp.octave < minLevel) pass = false;
                    if (maxLevel >= 0 && kp.octave > maxLevel) pass = false;
                }
                const float dx = __fsub_rn(kp.x, x), dy = __fsub_rn(kp.y, y);
                if (!(fabsf(dx) < r && fabsf(dy) < r)) pass = false;
            }
            if (out_idx) {
                unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) out_idx[nout + __popc(bal & ((1u << lane) - 1))] = j;
                nout += __popc(bal);
            }
            if (pass && desc && !(skip && skip[j])) {
                const int d = ham256_g(q, desc + (size_t)32 * j);
                top2_push(t, ((uint32_t)d << 20) | (uint32_t)(tpos + (k - k0)), j);
            }
            tpos += 32;
        }
    }
    if (out_n) *out_n = nout;
    return top2_warp_merge(t);
}

__device__ __forceinline__ void load_desc(uint32_t q[8], const uint8_t *d)
{
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(d)), b = __ldg(reinterpret_cast<const uint4 *>(d) + 1);
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
}

// per-query result record
struct QRes { int ok, q, t, dist, bin; };

__device__ __forceinline__ int rot_bin(float a_from, float a_to, int histo)
{
    float rot = __fsub_rn(a_from, a_to);
    if (rot < 0) rot = __fadd_rn(rot, 360.f);
    const float factor = __fdiv_rn(1.f, (float)histo);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == histo) bin = 0;
    return bin;
}

// searchByViolence inner loop: one warp per F1 keypoint
__global__ void __launch_bounds__(256)
violence_query(GridDev g2, const uint8_t *d2, const tb_keypoint *k1, const uint8_t *d1, int n1,
               int min_level, int max_level, float search_r, tb_matcher_params p, QRes *res)
{
    const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i1 >= n1) return;
    uint32_t q[8];
    load_desc(q, d1 + (size_t)32 * i1);
    const tb_keypoint kp = k1[i1];
    Top2 t = window_scan(g2, kp.x, kp.y, search_r, min_level, max_level, q, d2, nullptr, nullptr, nullptr);
    if ((threadIdx.x & 31) == 0) {
        QRes r = {0, i1, -1, 0, 0};
        if (t.i1 >= 0) {
            const int bestDist = (int)(t.k1 >> 20);
            const int bestDist2 = t.i2 >= 0 ? (int)(t.k2 >> 20) : 0x7fffffff;
            if (bestDist <= p.TH_LOW && (float)bestDist < __fmul_rn((float)bestDist2, p.nRatio)) {
                r.ok = 1; r.t = t.i1; r.dist = bestDist;
                if (p.checkOrientation) r.bin = rot_bin(kp.angle, g2.kps[t.i1].angle, p.HISTO_LENGTH);
            }
        }
        res[i1] = r;
    }
}

// Eigen's fixed-size 3-term reduction order: x0 + (x1 + x2)  (see oracle/tb_oracle.c:dot3)
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return __fadd_rn(__fmul_rn(a0, b0), __fadd_rn(__fmul_rn(a1, b1), __fmul_rn(a2, b2)));
}
struct Pose { float T[16]; };
__device__ __forceinline__ void transform_pt(const Pose &P, const float *X, float *o)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
        o[r] = __fadd_rn(dot3(P.T[4 * r], X[0], P.T[4 * r + 1], X[1], P.T[4 * r + 2], X[2]), P.T[4 * r + 3]);
}
__device__ __forceinline__ void world2cam(const tb_pinhole &c, const float *xyz, float &u, float &v)
{
    const float x = __fdiv_rn(xyz[0], xyz[2]), y = __fdiv_rn(xyz[1], xyz[2]);
    if (!(fabs((double)c.d[0]) > 0.0000001)) {
        u = __fadd_rn(__fmul_rn(c.fx, x), c.cx);
        v = __fadd_rn(__fmul_rn(c.fy, y), c.cy);
    } else {
        const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), r4 = __fmul_rn(r2, r2), r6 = __fmul_rn(r4, r2);
        const float a1 = __fmul_rn(__fmul_rn(2.f, x), y);
        const float a2 = __fadd_rn(r2, __fmul_rn(__fmul_rn(2.f, x), x)), a3 = __fadd_rn(r2, __fmul_rn(__fmul_rn(2.f, y), y));
        const float cdist = __fadd_rn(__fadd_rn(__fadd_rn(1.f, __fmul_rn(c.d[0], r2)), __fmul_rn(c.d[1], r4)), __fmul_rn(c.d[4], r6));
        const float xd = __fadd_rn(__fadd_rn(__fmul_rn(x, cdist), __fmul_rn(c.d[2], a1)), __fmul_rn(c.d[3], a2));
        const float yd = __fadd_rn(__fadd_rn(__fmul_rn(y, cdist), __fmul_rn(c.d[2], a3)), __fmul_rn(c.d[3], a1));
        u = __fadd_rn(__fmul_rn(xd, c.fx), c.cx);
        v = __fadd_rn(__fmul_rn(yd, c.fy), c.cy);
    }
}
__device__ __forceinline__ bool is_in_frame(const tb_pinhole &c, float u, float v)
{
    const int iu = (int)u, iv = (int)v;
    return iu >= 0 && iu < (int)((float)c.width * 1.f) && iv >= 0 && iv < (int)((float)c.height * 1.f);
}

// searchByProjection(F1,F2) inner loop: one warp per F2 keypoint carrying a map point
struct SfTab { float sf[TB_MAX_LEVELS]; };
__global__ void __launch_bounds__(256)
projection_query(GridDev g1, const uint8_t *d1, const uint8_t *f1_matched, SfTab sf1, tb_pinhole cam, Pose T,
                 const tb_keypoint *k2, int n2, const uint8_t *mp_valid, const float *mp_xyz, const uint8_t *mp_desc,
                 tb_matcher_params p, QRes *res)
{
    const int i2 = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i2 >= n2) return;
    QRes r = {0, -1, i2, 0, 0};
    bool go = mp_valid[i2] != 0;
    float u = 0, v = 0;
    if (go) {
        float xc[3];
        transform_pt(T, mp_xyz + 3 * (size_t)i2, xc);
        const float invzc = __fdiv_rn(1.0f, xc[2]);
        if (invzc < 0) go = false;
        else {
            world2cam(cam, xc, u, v);
            if (!is_in_frame(cam, u, v)) go = false;
        }
    }
    if (go) {
        const tb_keypoint kp2 = k2[i2];
        const int oct = kp2.octave;
        const float radius = __fmul_rn(p.nRatio, sf1.sf[oct]);
        uint32_t q[8];
        load_desc(q, mp_desc + (size_t)32 * i2);
        Top2 t = window_scan(g1, u, v, radius, oct - 1, oct + 1, q, d1, f1_matched, nullptr, nullptr);
        if (t.i1 >= 0) {
            const int bestDist = (int)(t.k1 >> 20);
            if (bestDist <= p.TH_HIGH) { // bestDist starts at 256 in the reference: a candidate at 256 cannot win
                if (bestDist < 256) {
                    r.ok = 1; r.q = t.i1; r.dist = bestDist;
                    if (p.checkOrientation) r.bin = rot_bin(kp2.angle, g1.kps[t.i1].angle, p.HISTO_LENGTH);
                }
            }
        }
    }
    if ((threadIdx.x & 31) == 0) res[i2] = r;
}

// searchByProjection(map,F1,radio) inner loop: one warp per map point
__global__ void __launch_bounds__(256)
projection_map_query(GridDev g1, const uint8_t *d1, const uint8_t *f1_matched, float sf0, tb_pinhole cam, Pose T,
                     int n_mp, const uint8_t *mp_bad, const float *mp_xyz, const float *mp_normal, const uint8_t *mp_desc,
                     tb_matcher_params p, float radio, QRes *res)
{
    const int im = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (im >= n_mp) return;
    QRes r = {0, -1, im, 0, 0};
    bool go = mp_bad[im] == 0;
    float u = 0, v = 0, viewCos = 0;
    if (go) {
        const float *P = mp_xyz + 3 * (size_t)im;
        float Pc[3];
        transform_pt(T, P, Pc);
        if (Pc[2] < 0.0f) go = false;
        else {
            world2cam(cam, Pc, u, v);
            if (!is_in_frame(cam, u, v)) go = false;
        }
        if (go) {
            float Ow[3];
#pragma unroll
            for (int c = 0; c < 3; c++) Ow[c] = dot3(-T.T[0 + c], T.T[3], -T.T[4 + c], T.T[7], -T.T[8 + c], T.T[11]); // Frame.cpp:57
            const float PO[3] = {__fsub_rn(P[0], Ow[0]), __fsub_rn(P[1], Ow[1]), __fsub_rn(P[2], Ow[2])};
            const float dist3 = __fsqrt_rn(dot3(PO[0], PO[0], PO[1], PO[1], PO[2], PO[2]));
            if (dist3 < 1.f || dist3 > 1000.f) go = false; // MapPoint.cpp:207-217
            else {
                const float *Pn = mp_normal + 3 * (size_t)im;
                viewCos = __fdiv_rn(dot3(PO[0], Pn[0], PO[1], Pn[1], PO[2], Pn[2]), dist3);
                if (viewCos < 0.5f) go = false;
            }
        }
    }
    if (go) {
        float rr = 4.f;
        if ((double)viewCos > 0.998) rr = 2.5f;
        if ((double)p.nRatio != 1.0) rr = __fmul_rn(rr, p.nRatio);
        uint32_t q[8];
        load_desc(q, mp_desc + (size_t)32 * im);
        Top2 t = window_scan(g1, u, v, __fmul_rn(rr, sf0), -1, 0, q, d1, f1_matched, nullptr, nullptr);
        if (t.i1 >= 0) {
            const int bestDist = (int)(t.k1 >> 20);
            if (bestDist < 256 && bestDist <= p.TH_HIGH) {
                const int bestLevel = g1.kps[t.i1].octave;
                int bestLevel2 = -1, bestDist2 = 256;
                if (t.i2 >= 0 && (int)(t.k2 >> 20) < 256) { bestLevel2 = g1.kps[t.i2].octave; bestDist2 = (int)(t.k2 >> 20); }
                if (!(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(radio, (float)bestDist2))) {
                    r.ok = 1; r.q = t.i1; r.dist = bestDist;
                }
            }
        }
    }
    if ((threadIdx.x & 31) == 0) res[im] = r;
}

__global__ void __launch_bounds__(32)
area_query(GridDev g, float x, float y, float r, int minLevel, int maxLevel, int *out, int *out_n)
{
    uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n = 0;
    window_scan(g, x, y, r, minLevel, maxLevel, q, nullptr, nullptr, out, &n);
    if (threadIdx.x == 0) *out_n = n;
}

// ---- emission: ordered compaction of per-query results, optional rotation-histogram regrouping.
__device__ int scan256w(int *a, int n)
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

// One CTA. With check_orientation: output = raw matches of the (up to) three maximal bins,
// grouped by bin ascending, raw order inside a bin (matcher.cpp:377-392). Else raw order.
__global__ void __launch_bounds__(256)
emit_matches(const QRes *res, int n, int check_orientation, int histo, tb_match *out, int capacity, int *out_n)
{
    __shared__ int s_hist[256];
    __shared__ int s_sel[3], s_nsel;
    __shared__ int s_chunk[2048];
    __shared__ int s_run;
    const int tid = threadIdx.x;
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    if (check_orientation) {
        for (int i = tid; i < n; i += 256) if (res[i].ok) atomicAdd(&s_hist[res[i].bin], 1);
        __syncthreads();
        if (tid == 0) {
            int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0; // ComputeThreeMaxima :810-851
            for (int i = 0; i < histo; i++) {
                const int s = s_hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
            int ns = 0;
            for (int b = 0; b < histo; b++) if (b == ind1 || b == ind2 || b == ind3) s_sel[ns++] = b;
            s_nsel = ns;
        }
    } else if (tid == 0) { s_nsel = 1; s_sel[0] = -1; }
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int sidx = 0; sidx < s_nsel; sidx++) {
        const int bsel = s_sel[sidx];
        for (int base = 0; base < n; base += 2048) {
            const int m = min(2048, n - base);
            for (int i = tid; i < m; i += 256) {
                const QRes r = res[base + i];
                s_chunk[i] = (r.ok && (bsel < 0 || r.bin == bsel)) ? 1 : 0;
            }
            __syncthreads();
            // remember flags before the scan overwrites them
            int mine[8]; int nm = 0;
            for (int i = tid; i < m; i += 256) mine[nm++] = s_chunk[i];
            __syncthreads();
            const int tot = scan256w(s_chunk, m);
            const int run = s_run;
            nm = 0;
            for (int i = tid; i < m; i += 256) {
                if (mine[nm++]) {
                    const int pos = run + s_chunk[i];
                    if (pos < capacity) {
                        const QRes r = res[base + i];
                        tb_match mt = {r.q, r.t, -1, (float)r.dist};
                        out[pos] = mt;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) s_run = run + tot;
            __syncthreads();
        }
    }
    if (tid == 0) *out_n = s_run;
}

// ------------------------------------------------------------------------------------------ host side
static int build_grid(tb_ctx *ctx, ArenaScope &buf, const tb_keypoint *d_kps, int n, int cols, int rows, GridDev &g)
{
    int *cellid = buf.alloc<int>(n), *count = buf.alloc<int>(GRID_CELLS), *start = buf.alloc<int>(GRID_CELLS + 1), *items = buf.alloc<int>(n);
    if (!cellid || !count || !start || !items) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    volatile float hInv = (float)GRID_COLS / (float)cols, wInv = (float)GRID_ROWS / (float)rows; // Frame.cpp:30-31
    g.wInv = wInv; g.hInv = hInv; g.kps = d_kps; g.cell_start = start; g.items = items;
    TB_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * GRID_CELLS, ctx->stream));
    if (n > 0) { grid_cell_ids<<<div_up(n, 256), 256, 0, ctx->stream>>>(d_kps, n, g.wInv, g.hInv, cellid, count); TB_LAUNCH_CHECK(ctx); }
    grid_scan<<<1, 1024, 0, ctx->stream>>>(count, start); TB_LAUNCH_CHECK(ctx);
    if (n > 0) { grid_fill<<<div_up(GRID_CELLS, 128), 128, 0, ctx->stream>>>(cellid, n, count, start, items); TB_LAUNCH_CHECK(ctx); }
    return TB_OK;
}

static int emit_and_download(tb_ctx *ctx, ArenaScope &buf, const QRes *res, int nq, int check, int histo, tb_match *out, int capacity, int *n)
{
    if (histo > 256 || histo < 1) return tb_fail(ctx, TB_E_UNSUPPORTED, "HISTO_LENGTH must be in [1,256]");
    if (check && histo < 19) return tb_fail(ctx, TB_E_INVALID, "checkOrientation needs HISTO_LENGTH >= 19: bins reach round(360/HISTO_LENGTH) and the reference asserts bin < HISTO_LENGTH");
    tb_match *d_out = buf.alloc<tb_match>(nq);
    int *d_n = buf.alloc<int>(1);
    if (!d_out || !d_n) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    emit_matches<<<1, 256, 0, ctx->stream>>>(res, nq, check, histo, d_out, nq, d_n);
    TB_LAUNCH_CHECK(ctx);
    int cnt = 0;
    TB_CUDA(ctx, cudaMemcpyAsync(&cnt, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    int m = cnt < capacity ? cnt : capacity;
    if (m > 0) TB_CUDA(ctx, cudaMemcpy(out, d_out, sizeof(tb_match) * m, cudaMemcpyDeviceToHost));
    return cnt > capacity ? tb_fail(ctx, TB_E_CAPACITY, "match output capacity too small") : TB_OK;
}

extern "C" int tb_features_in_area(tb_ctx *ctx, const tb_keypoint *kps, int n, int cols, int rows, float x, float y, float r,
                                   int minLevel, int maxLevel, int32_t *out, int capacity, int *n_out)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || n < 0 || !n_out || cols <= 0 || rows <= 0) return TB_E_INVALID;
    ArenaScope buf(ctx);
    tb_keypoint *dk = buf.upload(kps, n);
    int *d_out = buf.alloc<int>(n + 32), *d_n = buf.alloc<int>(1);
    if (!dk || !d_out || !d_n) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    GridDev g;
    int rc = build_grid(ctx, buf, dk, n, cols, rows, g);
    if (rc) return rc;
    area_query<<<1, 32, 0, ctx->stream>>>(g, x, y, r, minLevel, maxLevel, d_out, d_n);
    TB_LAUNCH_CHECK(ctx);
    int cnt = 0;
    TB_CUDA(ctx, cudaMemcpyAsync(&cnt, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = cnt;
    int m = cnt < capacity ? cnt : capacity;
    if (m > 0) TB_CUDA(ctx, cudaMemcpy(out, d_out, sizeof(int) * m, cudaMemcpyDeviceToHost));
    return cnt > capacity ? TB_E_CAPACITY : TB_OK;
}

extern "C" int tb_search_by_violence(tb_ctx *ctx, const tb_matcher_params *p, const tb_keypoint *k1, const uint8_t *d1, int n1,
                                     const tb_keypoint *k2, const uint8_t *d2, int n2, int cols2, int rows2,
                                     int min_level, int max_level, float search_r, tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || n1 < 0 || n2 < 0 || !n) return TB_E_INVALID;
    *n = 0;
    if (n1 == 0) return TB_OK;
    ArenaScope buf(ctx);
    tb_keypoint *dk1 = buf.upload(k1, n1), *dk2 = buf.upload(k2, n2);
    uint8_t *dd1 = buf.upload(d1, (size_t)32 * n1), *dd2 = buf.upload(d2, (size_t)32 * n2);
    QRes *res = buf.alloc<QRes>(n1);
    if (!dk1 || !dk2 || !dd1 || !dd2 || !res) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    GridDev g;
    int rc = build_grid(ctx, buf, dk2, n2, cols2, rows2, g);
    if (rc) return rc;
    violence_query<<<div_up(n1, 8), 256, 0, ctx->stream>>>(g, dd2, dk1, dd1, n1, min_level, max_level, search_r, *p, res);
    TB_LAUNCH_CHECK(ctx);
    return emit_and_download(ctx, buf, res, n1, p->checkOrientation, p->HISTO_LENGTH, out, capacity, n);
}

extern "C" int tb_search_by_projection(tb_ctx *ctx, const tb_matcher_params *p, const tb_pinhole *cam, const float *Tcw1,
                                       const tb_keypoint *k1, const uint8_t *d1, int n1, const uint8_t *f1_matched,
                                       const float *sf1, int n_levels1, int cols1, int rows1,
                                       const tb_keypoint *k2, int n2, const uint8_t *mp_valid, const float *mp_xyz,
                                       const uint8_t *mp_desc, tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || !cam || !Tcw1 || !sf1 || n1 < 0 || n2 < 0 || !n || n_levels1 <= 0 || n_levels1 > TB_MAX_LEVELS) return TB_E_INVALID;
    *n = 0;
    if (n2 == 0) return TB_OK;
    for (int i = 0; i < n2; i++) if (k2[i].octave < 0 || k2[i].octave >= n_levels1) return tb_fail(ctx, TB_E_INVALID, "octave outside the scale table");
    ArenaScope buf(ctx);
    tb_keypoint *dk1 = buf.upload(k1, n1), *dk2 = buf.upload(k2, n2);
    uint8_t *dd1 = buf.upload(d1, (size_t)32 * n1);
    uint8_t *dfm = f1_matched ? buf.upload(f1_matched, (size_t)n1) : nullptr;
    uint8_t *dval = buf.upload(mp_valid, (size_t)n2);
    float *dxyz = buf.upload(mp_xyz, (size_t)3 * n2);
    uint8_t *dmd = buf.upload(mp_desc, (size_t)32 * n2);
    QRes *res = buf.alloc<QRes>(n2);
    if (!dk1 || !dk2 || !dd1 || !dval || !dxyz || !dmd || !res) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    GridDev g;
    int rc = build_grid(ctx, buf, dk1, n1, cols1, rows1, g);
    if (rc) return rc;
    SfTab sf; memset(&sf, 0, sizeof sf);
    for (int i = 0; i < n_levels1; i++) sf.sf[i] = sf1[i];
    Pose T; memcpy(T.T, Tcw1, sizeof T.T);
    projection_query<<<div_up(n2, 8), 256, 0, ctx->stream>>>(g, dd1, dfm, sf, *cam, T, dk2, n2, dval, dxyz, dmd, *p, res);
    TB_LAUNCH_CHECK(ctx);
    return emit_and_download(ctx, buf, res, n2, p->checkOrientation, p->HISTO_LENGTH, out, capacity, n);
}

extern "C" int tb_search_by_projection_map(tb_ctx *ctx, const tb_matcher_params *p, const tb_pinhole *cam, const float *Tcw1,
                                           const tb_keypoint *k1, const uint8_t *d1, int n1, const uint8_t *f1_matched,
                                           const float *sf1, int cols1, int rows1,
                                           int n_mp, const uint8_t *mp_bad, const float *mp_xyz, const float *mp_normal,
                                           const uint8_t *mp_desc, float radio, tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || !cam || !Tcw1 || !sf1 || n1 < 0 || n_mp < 0 || !n) return TB_E_INVALID;
    *n = 0;
    if (n_mp == 0) return TB_OK;
    ArenaScope buf(ctx);
    tb_keypoint *dk1 = buf.upload(k1, n1);
    uint8_t *dd1 = buf.upload(d1, (size_t)32 * n1);
    uint8_t *dfm = f1_matched ? buf.upload(f1_matched, (size_t)n1) : nullptr;
    uint8_t *dbad = buf.upload(mp_bad, (size_t)n_mp);
    float *dxyz = buf.upload(mp_xyz, (size_t)3 * n_mp), *dnrm = buf.upload(mp_normal, (size_t)3 * n_mp);
    uint8_t *dmd = buf.upload(mp_desc, (size_t)32 * n_mp);
    QRes *res = buf.alloc<QRes>(n_mp);
    if (!dk1 || !dd1 || !dbad || !dxyz || !dnrm || !dmd || !res) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    GridDev g;
    int rc = build_grid(ctx, buf, dk1, n1, cols1, rows1, g);
    if (rc) return rc;
    Pose T; memcpy(T.T, Tcw1, sizeof T.T);
    projection_map_query<<<div_up(n_mp, 8), 256, 0, ctx->stream>>>(g, dd1, dfm, sf1[0], *cam, T, n_mp, dbad, dxyz, dnrm, dmd, *p, radio, res);
    TB_LAUNCH_CHECK(ctx);
    // no orientation check in this overload (matcher.cpp:539-617)
    return emit_and_download(ctx, buf, res, n_mp, 0, p->HISTO_LENGTH > 0 ? p->HISTO_LENGTH : 1, out, capacity, n);
}

// ------------------------------------------------------------------------------------------
// Device-resident, batched searchByProjection over frames of one extract result (BASELINE config 5:
// test_vo-style tracking).  Nothing leaves the GPU: grids are built per frame by one CTA, one warp per
// (pair, F2 keypoint) walks the window, one CTA per pair applies the rotation histogram and compacts.
// ------------------------------------------------------------------------------------------

// CSR grid of every frame: cell_start[f][GRID_CELLS+1], items[f][kp_cap] (insertion = index order)
__global__ void __launch_bounds__(1024)
grid_build_batched(const tb_keypoint *__restrict__ kps, int kp_cap, const int *__restrict__ counts, float wInv, float hInv,
                   int *__restrict__ cell_start, int *__restrict__ items)
{
    __shared__ int s_cnt[GRID_CELLS + 1];
    __shared__ int s_part[1024];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(counts[f], kp_cap);
    const tb_keypoint *k = kps + (size_t)f * kp_cap;
    auto cell_of = [&](int i) {
        const int px = (int)roundf(__fmul_rn(k[i].x, wInv)), py = (int)roundf(__fmul_rn(k[i].y, hInv));
        return (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? -1 : px * GRID_ROWS + py;
    };
    for (int i = tid; i <= GRID_CELLS; i += 1024) s_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int c = cell_of(i);
        if (c >= 0) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan of the 4320 cell counts
    const int ipt = (GRID_CELLS + 1023) / 1024;
    const int b0 = min(tid * ipt, GRID_CELLS), b1 = min(b0 + ipt, GRID_CELLS);
    int sum = 0;
    for (int i = b0; i < b1; i++) sum += s_cnt[i];
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { int v = tid >= o ? s_part[tid - o] : 0; __syncthreads(); s_part[tid] += v; __syncthreads(); }
    int run = s_part[tid] - sum;
    int *cs = cell_start + (size_t)f * (GRID_CELLS + 1);
    for (int i = b0; i < b1; i++) { const int c = s_cnt[i]; s_cnt[i] = run; cs[i] = run; run += c; }
    if (tid == 1023) cs[GRID_CELLS] = s_part[1023];
    __syncthreads();
    // stable placement in index order (the reference's push_back order, Frame.cpp:193-199): one warp walks the keypoints 32
    // at a time; lanes of the same cell rank themselves with match_any, the cell's cursor advances by the group size
    if (tid < 32) {
        int *it = items + (size_t)f * kp_cap;
        for (int base = 0; base < n; base += 32) {
            const int i = base + tid;
            const int c = i < n ? cell_of(i) : -1;
            const unsigned act = __ballot_sync(0xffffffffu, c >= 0);
            if (c >= 0) {
                const unsigned same = __match_any_sync(act, c);
                const int rank = __popc(same & ((1u << tid) - 1));
                const int start = s_cnt[c];
                it[start + rank] = i;
                __syncwarp(act);
                if (rank == 0) s_cnt[c] = start + __popc(same);
            }
            __syncwarp();
        }
    }
}

struct ProjPair { int f1, f2; float T[16]; };

__global__ void __launch_bounds__(256)
projection_query_batched(const ProjPair *__restrict__ pairs, const tb_keypoint *__restrict__ kps, const uint8_t *__restrict__ desc,
                         int kp_cap, const int *__restrict__ counts, const int *__restrict__ cell_start, const int *__restrict__ items,
                         float wInv, float hInv, const uint8_t *__restrict__ f1_matched, SfTab sf1, tb_pinhole cam,
                         const uint8_t *__restrict__ mp_valid, const float *__restrict__ mp_xyz, tb_matcher_params p, QRes *__restrict__ res)
{
    const ProjPair &pr = pairs[blockIdx.y];
    const int i2 = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int n2 = min(counts[pr.f2], kp_cap);
    if (i2 >= n2) return;
    GridDev g1;
    g1.cell_start = cell_start + (size_t)pr.f1 * (GRID_CELLS + 1);
    g1.items = items + (size_t)pr.f1 * kp_cap;
    g1.kps = kps + (size_t)pr.f1 * kp_cap;
    g1.wInv = wInv; g1.hInv = hInv;
    const uint8_t *d1 = desc + (size_t)pr.f1 * kp_cap * 32;
    const uint8_t *matched = f1_matched ? f1_matched + (size_t)pr.f1 * kp_cap : nullptr;
    const size_t o2 = (size_t)pr.f2 * kp_cap + i2;
    QRes r = {0, -1, i2, 0, 0};
    bool go = mp_valid[o2] != 0;
    float u = 0, v = 0;
    Pose T;
#pragma unroll
    for (int k = 0; k < 16; k++) T.T[k] = pr.T[k];
    if (go) {
        float xc[3];
        transform_pt(T, mp_xyz + 3 * o2, xc);
        const float invzc = __fdiv_rn(1.0f, xc[2]);
        if (invzc < 0) go = false;
        else {
            world2cam(cam, xc, u, v);
            if (!is_in_frame(cam, u, v)) go = false;
        }
    }
    if (go) {
        const tb_keypoint kp2 = kps[o2];
        const int oct = kp2.octave;
        const float radius = __fmul_rn(p.nRatio, sf1.sf[oct]);
        uint32_t q[8];
        load_desc(q, desc + o2 * 32);   // the map point carries the descriptor of the keypoint that created it
        Top2 t = window_scan(g1, u, v, radius, oct - 1, oct + 1, q, d1, matched, nullptr, nullptr);
        if (t.i1 >= 0) {
            const int bestDist = (int)(t.k1 >> 20);
            if (bestDist <= p.TH_HIGH && bestDist < 256) {
                r.ok = 1; r.q = t.i1; r.dist = bestDist;
                if (p.checkOrientation) r.bin = rot_bin(kp2.angle, g1.kps[t.i1].angle, p.HISTO_LENGTH);
            }
        }
    }
    if ((threadIdx.x & 31) == 0) res[(size_t)blockIdx.y * kp_cap + i2] = r;
}

__global__ void __launch_bounds__(256)
emit_matches_batched(const ProjPair *__restrict__ pairs, const QRes *__restrict__ res, int kp_cap, const int *__restrict__ counts,
                     int check_orientation, int histo, tb_match *__restrict__ out, int capacity, int *__restrict__ out_n, int queries_from_f1)
{
    // same logic as emit_matches, one CTA per pair
    __shared__ int s_hist[256];
    __shared__ int s_sel[3], s_nsel;
    __shared__ int s_chunk[2048];
    __shared__ int s_run;
    const int tid = threadIdx.x, pidx = blockIdx.x;
    const int n = min(counts[queries_from_f1 ? pairs[pidx].f1 : pairs[pidx].f2], kp_cap);
    const QRes *rs = res + (size_t)pidx * kp_cap;
    tb_match *o = out + (size_t)pidx * capacity;
    s_hist[tid] = 0;
    __syncthreads();
    if (check_orientation) {
        for (int i = tid; i < n; i += 256) if (rs[i].ok) atomicAdd(&s_hist[rs[i].bin], 1);
        __syncthreads();
        if (tid == 0) {
            int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < histo; i++) {
                const int s = s_hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
            int ns = 0;
            for (int b = 0; b < histo; b++) if (b == ind1 || b == ind2 || b == ind3) s_sel[ns++] = b;
            s_nsel = ns;
        }
    } else if (tid == 0) { s_nsel = 1; s_sel[0] = -1; }
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int sidx = 0; sidx < s_nsel; sidx++) {
        const int bsel = s_sel[sidx];
        for (int base = 0; base < n; base += 2048) {
            const int m = min(2048, n - base);
            int mine[8]; int nm = 0;
            for (int i = tid; i < m; i += 256) {
                const QRes r = rs[base + i];
                const int fl = (r.ok && (bsel < 0 || r.bin == bsel)) ? 1 : 0;
                s_chunk[i] = fl; mine[nm++] = fl;
            }
            __syncthreads();
            const int tot = scan256w(s_chunk, m);
            const int run = s_run;
            nm = 0;
            for (int i = tid; i < m; i += 256) {
                if (mine[nm++]) {
                    const int pos = run + s_chunk[i];
                    if (pos < capacity) { const QRes r = rs[base + i]; tb_match mt = {r.q, r.t, -1, (float)r.dist}; o[pos] = mt; }
                }
            }
            __syncthreads();
            if (tid == 0) s_run = run + tot;
            __syncthreads();
        }
    }
    if (tid == 0) out_n[pidx] = s_run;
}

// searchByViolence over frames of one frontend: pair p matches the keypoints of frame f1 inside windows on the grid of f2
__global__ void __launch_bounds__(256)
violence_query_batched(const ProjPair *__restrict__ pairs, const tb_keypoint *__restrict__ kps, const uint8_t *__restrict__ desc,
                       int kp_cap, const int *__restrict__ counts, const int *__restrict__ cell_start, const int *__restrict__ items,
                       float wInv, float hInv, int min_level, int max_level, float search_r, tb_matcher_params p, QRes *__restrict__ res)
{
    const ProjPair &pr = pairs[blockIdx.y];
    const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int n1 = min(counts[pr.f1], kp_cap);
    if (i1 >= n1) return;
    GridDev g2;
    g2.cell_start = cell_start + (size_t)pr.f2 * (GRID_CELLS + 1);
    g2.items = items + (size_t)pr.f2 * kp_cap;
    g2.kps = kps + (size_t)pr.f2 * kp_cap;
    g2.wInv = wInv; g2.hInv = hInv;
    const uint8_t *d2 = desc + (size_t)pr.f2 * kp_cap * 32;
    const size_t o1 = (size_t)pr.f1 * kp_cap + i1;
    uint32_t q[8];
    load_desc(q, desc + o1 * 32);
    const tb_keypoint kp = kps[o1];
    Top2 t = window_scan(g2, kp.x, kp.y, search_r, min_level, max_level, q, d2, nullptr, nullptr, nullptr);
    if ((threadIdx.x & 31) == 0) {
        QRes r = {0, i1, -1, 0, 0};
        if (t.i1 >= 0) {
            const int bestDist = (int)(t.k1 >> 20);
            const int bestDist2 = t.i2 >= 0 ? (int)(t.k2 >> 20) : 0x7fffffff;
            if (bestDist <= p.TH_LOW && (float)bestDist < __fmul_rn((float)bestDist2, p.nRatio)) {
                r.ok = 1; r.t = t.i1; r.dist = bestDist;
                if (p.checkOrientation) r.bin = rot_bin(kp.angle, g2.kps[t.i1].angle, p.HISTO_LENGTH);
            }
        }
        res[(size_t)blockIdx.y * kp_cap + i1] = r;
    }
}

struct WinWs { int *cell_start = nullptr, *items = nullptr; QRes *res = nullptr; ProjPair *pairs = nullptr; size_t frames = 0, pairs_cap = 0, kp_cap = 0; };

static int win_ws_reserve(tb_ctx *ctx, int n_frames, int n_pairs, int kp_cap, bool need_grid)
{
    if (!ctx->wws) ctx->wws = new WinWs();
    WinWs &w = *ctx->wws;
    if ((size_t)kp_cap != w.kp_cap || (need_grid && (size_t)n_frames > w.frames)) {
        cudaFree(w.cell_start); cudaFree(w.items); cudaFree(w.res); cudaFree(w.pairs);
        w.cell_start = nullptr; w.items = nullptr; w.res = nullptr; w.pairs = nullptr; w.frames = 0; w.pairs_cap = 0; w.kp_cap = kp_cap;
    }
    int rc;
    if (need_grid && (size_t)n_frames > w.frames) {
        w.frames = 0;
        if ((rc = tb_regrow(ctx, w.cell_start, (size_t)n_frames * (GRID_CELLS + 1))) || (rc = tb_regrow(ctx, w.items, (size_t)n_frames * kp_cap))) return rc;
        w.frames = n_frames;
    }
    if ((size_t)n_pairs > w.pairs_cap) {
        w.pairs_cap = 0;
        if ((rc = tb_regrow(ctx, w.res, (size_t)n_pairs * kp_cap)) || (rc = tb_regrow(ctx, w.pairs, (size_t)n_pairs))) return rc;
        w.pairs_cap = n_pairs;
    }
    return TB_OK;
}

int tb_window_projection_batched(tb_ctx *ctx, const tb_matcher_params *p, const tb_pinhole *cam, int n_pairs, const int32_t *f1_frames,
                                 const int32_t *f2_frames, const float *Tcw1, int n_frames, const tb_keypoint *kps, const uint8_t *desc,
                                 int kp_cap, const int *counts, const uint8_t *f1_matched, const uint8_t *mp_valid, const float *mp_xyz,
                                 const float *sf1, int n_levels, int cols, int rows, tb_match *matches, int capacity, int *out_counts)
{
    if (n_pairs <= 0) return TB_OK;
    if (p->HISTO_LENGTH < 1 || p->HISTO_LENGTH > 256) return tb_fail(ctx, TB_E_UNSUPPORTED, "HISTO_LENGTH must be in [1,256]");
    if (p->checkOrientation && p->HISTO_LENGTH < 19) return tb_fail(ctx, TB_E_INVALID, "checkOrientation needs HISTO_LENGTH >= 19: bins reach round(360/HISTO_LENGTH) and the reference asserts bin < HISTO_LENGTH");
    int rc = win_ws_reserve(ctx, n_frames, n_pairs, kp_cap, true);
    if (rc) return rc;
    WinWs &w = *ctx->wws;
    std::vector<ProjPair> hp(n_pairs);
    for (int i = 0; i < n_pairs; i++) { hp[i].f1 = f1_frames[i]; hp[i].f2 = f2_frames[i]; memcpy(hp[i].T, Tcw1 + 16 * (size_t)i, sizeof(float) * 16); }
    if ((rc = tb_stage_upload(ctx, w.pairs, hp.data(), sizeof(ProjPair) * n_pairs))) return rc;   // through the pinned ring: no host wait
    volatile float hInv = (float)GRID_COLS / (float)cols, wInv = (float)GRID_ROWS / (float)rows;   // Frame.cpp:30-31
    {
        ProfScope prof__(ctx, "grid_build_batched");
        grid_build_batched<<<n_frames, 1024, 0, ctx->stream>>>(kps, kp_cap, counts, wInv, hInv, w.cell_start, w.items);
        TB_LAUNCH_CHECK(ctx);
    }
    SfTab sf; memset(&sf, 0, sizeof sf);
    for (int i = 0; i < n_levels && i < TB_MAX_LEVELS; i++) sf.sf[i] = sf1[i];
    {
        ProfScope prof__(ctx, "projection_query_batched");
        projection_query_batched<<<dim3(div_up(kp_cap, 8), n_pairs), 256, 0, ctx->stream>>>(w.pairs, kps, desc, kp_cap, counts, w.cell_start, w.items,
                                                                                              wInv, hInv, f1_matched, sf, *cam, mp_valid, mp_xyz, *p, w.res);
        TB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope prof__(ctx, "emit_matches_batched");
        emit_matches_batched<<<n_pairs, 256, 0, ctx->stream>>>(w.pairs, w.res, kp_cap, counts, p->checkOrientation, p->HISTO_LENGTH, matches, capacity, out_counts, 0);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

// ------------------------------------------------------------------------------------------ bag of words
// Frame::SetBow (src/types/Frame.cpp:267-270) = ORBVocabulary::transform(features, bowVec, featVec, 4)
// (third_part/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194): every descriptor walks the vocabulary tree, at each
// node taking the child with the smallest Hamming distance (first minimum in child order, :1236-1248), until a leaf;
// the FeatureVector bucket is the node passed at level L-levelsup.  Matcher::searchByBow (matcher.cpp:619-721)
// then compares only features that fell into the same bucket.
// The tree lives in HBM as CSR (children in ascending node id = loadFromTextFile order); one thread walks one
// descriptor (6 levels x 10 children for ORBvoc: 60 distances, the 35 MB of node descriptors stay in L2).
struct tb_vocab {
    tb_ctx *ctx;
    int n_nodes, L;
    int *child_start, *children, *word_id;
    uint8_t *desc;
    double *weight;
};

__global__ void __launch_bounds__(128)
bow_transform(const int *__restrict__ child_start, const int *__restrict__ children, const int *__restrict__ word_of,
              const uint8_t *__restrict__ vdesc, const double *__restrict__ vweight, int L, int levelsup,
              const uint8_t *__restrict__ desc, long long n, int32_t *__restrict__ word_id, int32_t *__restrict__ node_id,
              double *__restrict__ weight)
{
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    uint32_t q[8];
    load_desc(q, desc + 32 * (size_t)i);
    const int nid_level = L - levelsup;
    int nid = nid_level <= 0 ? 0 : -1;
    int cur = 0, level = 0;
    int cs = child_start[0], ce = child_start[1];
    do {
        level++;
        int best = children[cs];
        int best_d = ham256_g(q, vdesc + 32 * (size_t)best);
        for (int c = cs + 1; c < ce; c++) {
            const int id = children[c];
            const int d = ham256_g(q, vdesc + 32 * (size_t)id);
            if (d < best_d) { best_d = d; best = id; }
        }
        cur = best;
        if (level == nid_level) nid = cur;
        cs = child_start[cur]; ce = child_start[cur + 1];
    } while (ce > cs);
    if (nid < 0) nid = cur;   // leaf above the bucket level: *nid is left unset by the reference (UB), pinned to the leaf
    if (word_id) word_id[i] = word_of[cur];
    if (node_id) node_id[i] = nid;
    if (weight) weight[i] = vweight[cur];
}

extern "C" int tb_vocab_create(tb_ctx *ctx, int n_nodes, int depth_levels, const int32_t *parent, const uint8_t *is_leaf,
                               const uint8_t *desc, const double *weight, tb_vocab **out)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !out || n_nodes < 2 || !parent || !is_leaf || !desc || !weight) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_vocab_create: bad arguments") : TB_E_INVALID;
    // children of a node in ascending id (the order loadFromTextFile pushes them); word ids to leaves in id order
    std::vector<int> start(n_nodes + 1, 0), kids(n_nodes - 1), word(n_nodes, -1);
    for (int i = 1; i < n_nodes; i++) {
        if (parent[i] < 0 || parent[i] >= i) return tb_fail(ctx, TB_E_INVALID, "tb_vocab_create: parent id must refer to an earlier node");
        start[parent[i] + 1]++;
    }
    for (int i = 0; i < n_nodes; i++) start[i + 1] += start[i];
    { std::vector<int> fill(start.begin(), start.end() - 1); for (int i = 1; i < n_nodes; i++) kids[fill[parent[i]]++] = i; }
    if (start[1] == start[0]) return tb_fail(ctx, TB_E_INVALID, "tb_vocab_create: the root has no children");
    int nw = 0;
    for (int i = 1; i < n_nodes; i++) {
        const bool leaf = start[i + 1] == start[i];
        if (is_leaf[i]) word[i] = nw++;        // word ids follow the file's leaf flag (TemplatedVocabulary.h:1403-1410)
        else if (leaf) word[i] = 0;            // childless node not flagged as a word: Node() default word_id 0 (:315)
    }
    tb_vocab *v = new tb_vocab();
    v->ctx = ctx; v->n_nodes = n_nodes; v->L = depth_levels;
    v->child_start = v->children = v->word_id = nullptr; v->desc = nullptr; v->weight = nullptr;
    bool ok = cudaMalloc(&v->child_start, sizeof(int) * (n_nodes + 1)) == cudaSuccess && cudaMalloc(&v->children, sizeof(int) * n_nodes) == cudaSuccess &&
              cudaMalloc(&v->word_id, sizeof(int) * n_nodes) == cudaSuccess && cudaMalloc(&v->desc, (size_t)32 * n_nodes) == cudaSuccess &&
              cudaMalloc(&v->weight, sizeof(double) * n_nodes) == cudaSuccess;
    if (ok) ok = cudaMemcpy(v->child_start, start.data(), sizeof(int) * (n_nodes + 1), cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(v->children, kids.data(), sizeof(int) * (n_nodes - 1), cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(v->word_id, word.data(), sizeof(int) * n_nodes, cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(v->desc, desc, (size_t)32 * n_nodes, cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(v->weight, weight, sizeof(double) * n_nodes, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        cudaFree(v->child_start); cudaFree(v->children); cudaFree(v->word_id); cudaFree(v->desc); cudaFree(v->weight);
        delete v;
        return tb_fail(ctx, TB_E_NOMEM, "tb_vocab_create: device allocation or upload failed");
    }
    *out = v;
    return TB_OK;
}

extern "C" void tb_vocab_destroy(tb_vocab *v)
{
    if (!v) return;
    cudaFree(v->child_start); cudaFree(v->children); cudaFree(v->word_id); cudaFree(v->desc); cudaFree(v->weight);
    delete v;
}

extern "C" int tb_bow_transform_device(tb_ctx *ctx, const tb_vocab *v, const uint8_t *d_desc, long long n, int levelsup,
                                       int32_t *d_word_id, int32_t *d_node_id, double *d_weight)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !v || n < 0 || (n > 0 && !d_desc)) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_bow_transform_device: bad arguments") : TB_E_INVALID;
    if (n == 0) return TB_OK;
    ProfScope prof__(ctx, "bow_transform");
    bow_transform<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(v->child_start, v->children, v->word_id, v->desc, v->weight, v->L, levelsup,
                                                                          d_desc, n, d_word_id, d_node_id, d_weight);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

extern "C" int tb_bow_transform(tb_ctx *ctx, const tb_vocab *v, const uint8_t *desc, int n, int levelsup,
                                int32_t *word_id, int32_t *node_id, double *weight)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !v || n < 0 || (n > 0 && !desc)) return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_bow_transform: bad arguments") : TB_E_INVALID;
    if (n == 0) return TB_OK;
    ArenaScope buf(ctx);
    uint8_t *dd = buf.upload(desc, (size_t)32 * n);
    int32_t *dw = buf.alloc<int32_t>(n), *dn = buf.alloc<int32_t>(n);
    double *dwt = buf.alloc<double>(n);
    if (!dd || !dw || !dn || !dwt) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    int rc = tb_bow_transform_device(ctx, v, dd, n, levelsup, dw, dn, dwt);
    if (rc) return rc;
    if (word_id) TB_CUDA(ctx, cudaMemcpyAsync(word_id, dw, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (node_id) TB_CUDA(ctx, cudaMemcpyAsync(node_id, dn, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (weight) TB_CUDA(ctx, cudaMemcpyAsync(weight, dwt, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

// searchByBow inner loops: one warp per F1 feature.  The reference walks the two FeatureVectors (std::map node ->
// feature indices in insertion order) in ascending node id; the slot of feature i1 in that walk is its rank under
// (node, index) among the features that have a bucket, counted here directly.  Candidates are F2's features of the same
// bucket in index order: "first strict minimum" = smallest (distance, index) key.
__global__ void __launch_bounds__(256)
bow_query(const tb_keypoint *__restrict__ k1, const uint8_t *__restrict__ d1, const int32_t *__restrict__ node1, int n1,
          const tb_keypoint *__restrict__ k2, const uint8_t *__restrict__ d2, const int32_t *__restrict__ node2, int n2,
          const uint8_t *__restrict__ mp_valid2, int map_point_only, tb_matcher_params p, QRes *__restrict__ res)
{
    const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i1 >= n1) return;
    const int node = node1[i1];
    if (node < 0) return;                       // not in the FeatureVector (stopped word)
    int before = 0;
    for (int i = lane; i < n1; i += 32) {
        const int nd = node1[i];
        before += (nd >= 0 && (nd < node || (nd == node && i < i1)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    uint32_t q[8];
    load_desc(q, d1 + (size_t)32 * i1);
    Top2 t = {0xFFFFFFFFu, 0xFFFFFFFFu, -1, -1};
    for (int j = lane; j < n2; j += 32) {
        if (node2[j] != node) continue;
        if (map_point_only && !(mp_valid2 && mp_valid2[j])) continue;
        const int d = ham256_g(q, d2 + (size_t)32 * j);
        if (d < 256) top2_push(t, ((uint32_t)d << 20) | (uint32_t)j, j);   // bestDist1/2 start at 256 (matcher.cpp:645-647)
    }
    t = top2_warp_merge(t);
    if (lane == 0) {
        QRes r = {0, i1, -1, 0, 0};
        if (t.i1 >= 0) {
            const int bestDist1 = (int)(t.k1 >> 20);
            const int bestDist2 = t.i2 >= 0 ? (int)(t.k2 >> 20) : 256;
            if (bestDist1 < p.TH_LOW && (float)bestDist1 < __fmul_rn(p.nRatio, (float)bestDist2)) {
                r.ok = 1; r.t = t.i1; r.dist = bestDist1;
                if (p.checkOrientation) r.bin = rot_bin(k1[i1].angle, k2[t.i1].angle, p.HISTO_LENGTH);
            }
        }
        res[before] = r;
    }
}

extern "C" int tb_search_by_bow(tb_ctx *ctx, const tb_matcher_params *p, const tb_keypoint *k1, const uint8_t *d1, const int32_t *node1, int n1,
                                const tb_keypoint *k2, const uint8_t *d2, const int32_t *node2, int n2, const uint8_t *mp_valid2,
                                int map_point_only, tb_match *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || n1 < 0 || n2 < 0 || !n) return TB_E_INVALID;
    *n = 0;
    if (n1 == 0 || n2 == 0) return TB_OK;
    if (n2 >= (1 << 20)) return tb_fail(ctx, TB_E_UNSUPPORTED, "searchByBow supports fewer than 2^20 features per frame");
    ArenaScope buf(ctx);
    tb_keypoint *dk1 = buf.upload(k1, n1), *dk2 = buf.upload(k2, n2);
    uint8_t *dd1 = buf.upload(d1, (size_t)32 * n1), *dd2 = buf.upload(d2, (size_t)32 * n2);
    int32_t *dn1 = buf.upload(node1, n1), *dn2 = buf.upload(node2, n2);
    uint8_t *dv = mp_valid2 ? buf.upload(mp_valid2, n2) : nullptr;
    QRes *res = buf.alloc<QRes>(n1);
    if (!dk1 || !dk2 || !dd1 || !dd2 || !dn1 || !dn2 || !res || (mp_valid2 && !dv)) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    TB_CUDA(ctx, cudaMemsetAsync(res, 0, sizeof(QRes) * n1, ctx->stream));   // slots past the bucketed features stay ok = 0
    {
        ProfScope prof__(ctx, "bow_query");
        bow_query<<<div_up(n1, 8), 256, 0, ctx->stream>>>(dk1, dd1, dn1, n1, dk2, dd2, dn2, n2, dv, map_point_only, *p, res);
        TB_LAUNCH_CHECK(ctx);
    }
    return emit_and_download(ctx, buf, res, n1, p->checkOrientation, p->HISTO_LENGTH, out, capacity, n);
}

// ---- device-resident, batched bag of words over the frames of a frontend
// FeatureVector of every frame: node[f][i] = bucket of keypoint i, -1 for a stopped word or i >= counts[f]
__global__ void __launch_bounds__(128)
bow_transform_frames(const int *__restrict__ child_start, const int *__restrict__ children, const int *__restrict__ word_of,
                     const uint8_t *__restrict__ vdesc, const double *__restrict__ vweight, int L, int levelsup,
                     const uint8_t *__restrict__ desc, int kp_cap, const int *__restrict__ counts,
                     int32_t *__restrict__ node_id, int32_t *__restrict__ word_id, double *__restrict__ weight)
{
    const int f = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (i >= kp_cap) return;
    const size_t o = (size_t)f * kp_cap + i;
    if (i >= min(counts[f], kp_cap)) { node_id[o] = -1; if (word_id) word_id[o] = -1; if (weight) weight[o] = 0.0; return; }
    uint32_t q[8];
    load_desc(q, desc + 32 * o);
    const int nid_level = L - levelsup;
    int nid = nid_level <= 0 ? 0 : -1;
    int cur = 0, level = 0;
    int cs = child_start[0], ce = child_start[1];
    do {
        level++;
        int best = children[cs];
        int best_d = ham256_g(q, vdesc + 32 * (size_t)best);
        for (int c = cs + 1; c < ce; c++) {
            const int id = children[c];
            const int d = ham256_g(q, vdesc + 32 * (size_t)id);
            if (d < best_d) { best_d = d; best = id; }
        }
        cur = best;
        if (level == nid_level) nid = cur;
        cs = child_start[cur]; ce = child_start[cur + 1];
    } while (ce > cs);
    if (nid < 0) nid = cur;
    const double w = vweight[cur];
    node_id[o] = w > 0 ? nid : -1;           // "if(w > 0) fv.addFeature(nid, i_feature)" (TemplatedVocabulary.h:1157-1161)
    if (word_id) word_id[o] = word_of[cur];
    if (weight) weight[o] = w;
}

int tb_window_set_bow_batched(tb_ctx *ctx, const tb_vocab *v, int n_frames, int levelsup, const uint8_t *desc, int kp_cap, const int *counts,
                              int32_t *node_id, int32_t *word_id, double *weight)
{
    if (n_frames <= 0) return TB_OK;
    ProfScope prof__(ctx, "bow_transform_frames");
    bow_transform_frames<<<dim3(div_up(kp_cap, 128), n_frames), 128, 0, ctx->stream>>>(v->child_start, v->children, v->word_id, v->desc, v->weight, v->L,
                                                                                        levelsup, desc, kp_cap, counts, node_id, word_id, weight);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

__global__ void __launch_bounds__(256)
bow_query_batched(const ProjPair *__restrict__ pairs, const tb_keypoint *__restrict__ kps, const uint8_t *__restrict__ desc, int kp_cap,
                  const int *__restrict__ counts, const int32_t *__restrict__ node, const uint8_t *__restrict__ mp_valid, int map_point_only,
                  tb_matcher_params p, QRes *__restrict__ res)
{
    const int f1 = pairs[blockIdx.y].f1, f2 = pairs[blockIdx.y].f2;
    const int n1 = min(counts[f1], kp_cap), n2 = min(counts[f2], kp_cap);
    const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i1 >= n1) return;
    QRes *rs = res + (size_t)blockIdx.y * kp_cap;
    const int32_t *node1 = node + (size_t)f1 * kp_cap, *node2 = node + (size_t)f2 * kp_cap;
    const int nd1 = node1[i1];
    int before = 0, bucketed = 0;
    for (int i = lane; i < n1; i += 32) {
        const int nd = node1[i];
        bucketed += (nd >= 0);
        before += (nd >= 0 && nd1 >= 0 && (nd < nd1 || (nd == nd1 && i < i1)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, o); bucketed += __shfl_xor_sync(0xffffffffu, bucketed, o); }
    if (nd1 < 0) {
        // features without a bucket fill the slots behind the bucketed ones with empty records (any order)
        int skipped = 0;
        for (int i = lane; i < i1; i += 32) skipped += (node1[i] < 0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) skipped += __shfl_xor_sync(0xffffffffu, skipped, o);
        if (lane == 0) { QRes r = {0, i1, -1, 0, 0}; rs[bucketed + skipped] = r; }
        return;
    }
    uint32_t q[8];
    load_desc(q, desc + 32 * ((size_t)f1 * kp_cap + i1));
    const uint8_t *d2 = desc + 32 * (size_t)f2 * kp_cap;
    const uint8_t *mv = mp_valid ? mp_valid + (size_t)f2 * kp_cap : nullptr;
    Top2 t = {0xFFFFFFFFu, 0xFFFFFFFFu, -1, -1};
    for (int j = lane; j < n2; j += 32) {
        if (node2[j] != nd1) continue;
        if (map_point_only && !(mv && mv[j])) continue;
        const int d = ham256_g(q, d2 + (size_t)32 * j);
        if (d < 256) top2_push(t, ((uint32_t)d << 20) | (uint32_t)j, j);
    }
    t = top2_warp_merge(t);
    if (lane == 0) {
        QRes r = {0, i1, -1, 0, 0};
        if (t.i1 >= 0) {
            const int bestDist1 = (int)(t.k1 >> 20);
            const int bestDist2 = t.i2 >= 0 ? (int)(t.k2 >> 20) : 256;
            if (bestDist1 < p.TH_LOW && (float)bestDist1 < __fmul_rn(p.nRatio, (float)bestDist2)) {
                r.ok = 1; r.t = t.i1; r.dist = bestDist1;
                if (p.checkOrientation) r.bin = rot_bin(kps[(size_t)f1 * kp_cap + i1].angle, kps[(size_t)f2 * kp_cap + t.i1].angle, p.HISTO_LENGTH);
            }
        }
        rs[before] = r;
    }
}

int tb_window_bow_batched(tb_ctx *ctx, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames, const int32_t *f2_frames, int n_frames,
                          const tb_keypoint *kps, const uint8_t *desc, int kp_cap, const int *counts, const int32_t *node,
                          const uint8_t *mp_valid, int map_point_only, tb_match *matches, int capacity, int *out_counts)
{
    (void)n_frames;
    if (n_pairs <= 0) return TB_OK;
    if (p->HISTO_LENGTH < 1 || p->HISTO_LENGTH > 256) return tb_fail(ctx, TB_E_UNSUPPORTED, "HISTO_LENGTH must be in [1,256]");
    if (p->checkOrientation && p->HISTO_LENGTH < 19) return tb_fail(ctx, TB_E_INVALID, "checkOrientation needs HISTO_LENGTH >= 19: bins reach round(360/HISTO_LENGTH) and the reference asserts bin < HISTO_LENGTH");
    int rc = win_ws_reserve(ctx, 0, n_pairs, kp_cap, false);
    if (rc) return rc;
    WinWs &w = *ctx->wws;
    std::vector<ProjPair> hp(n_pairs);
    memset(hp.data(), 0, sizeof(ProjPair) * n_pairs);
    for (int i = 0; i < n_pairs; i++) { hp[i].f1 = f1_frames[i]; hp[i].f2 = f2_frames[i]; }
    if ((rc = tb_stage_upload(ctx, w.pairs, hp.data(), sizeof(ProjPair) * n_pairs))) return rc;   // through the pinned ring: no host wait
    {
        ProfScope prof__(ctx, "bow_query_batched");
        bow_query_batched<<<dim3(div_up(kp_cap, 8), n_pairs), 256, 0, ctx->stream>>>(w.pairs, kps, desc, kp_cap, counts, node, mp_valid, map_point_only, *p, w.res);
        TB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope prof__(ctx, "emit_matches_batched");
        emit_matches_batched<<<n_pairs, 256, 0, ctx->stream>>>(w.pairs, w.res, kp_cap, counts, p->checkOrientation, p->HISTO_LENGTH, matches, capacity, out_counts, 1);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

int tb_window_violence_batched(tb_ctx *ctx, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames, const int32_t *f2_frames,
                               int n_frames, const tb_keypoint *kps, const uint8_t *desc, int kp_cap, const int *counts, int cols, int rows,
                               int min_level, int max_level, float search_r, tb_match *matches, int capacity, int *out_counts)
{
    if (n_pairs <= 0) return TB_OK;
    if (p->HISTO_LENGTH < 1 || p->HISTO_LENGTH > 256) return tb_fail(ctx, TB_E_UNSUPPORTED, "HISTO_LENGTH must be in [1,256]");
    if (p->checkOrientation && p->HISTO_LENGTH < 19) return tb_fail(ctx, TB_E_INVALID, "checkOrientation needs HISTO_LENGTH >= 19: bins reach round(360/HISTO_LENGTH) and the reference asserts bin < HISTO_LENGTH");
    int rc = win_ws_reserve(ctx, n_frames, n_pairs, kp_cap, true);
    if (rc) return rc;
    WinWs &w = *ctx->wws;
    std::vector<ProjPair> hp(n_pairs);
    memset(hp.data(), 0, sizeof(ProjPair) * n_pairs);
    for (int i = 0; i < n_pairs; i++) { hp[i].f1 = f1_frames[i]; hp[i].f2 = f2_frames[i]; }
    if ((rc = tb_stage_upload(ctx, w.pairs, hp.data(), sizeof(ProjPair) * n_pairs))) return rc;   // through the pinned ring: no host wait
    volatile float hInv = (float)GRID_COLS / (float)cols, wInv = (float)GRID_ROWS / (float)rows;   // Frame.cpp:30-31
    {
        ProfScope prof__(ctx, "grid_build_batched");
        grid_build_batched<<<n_frames, 1024, 0, ctx->stream>>>(kps, kp_cap, counts, wInv, hInv, w.cell_start, w.items);
        TB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope prof__(ctx, "violence_query_batched");
        violence_query_batched<<<dim3(div_up(kp_cap, 8), n_pairs), 256, 0, ctx->stream>>>(w.pairs, kps, desc, kp_cap, counts, w.cell_start, w.items,
                                                                                            wInv, hInv, min_level, max_level, search_r, *p, w.res);
        TB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope prof__(ctx, "emit_matches_batched");
        emit_matches_batched<<<n_pairs, 256, 0, ctx->stream>>>(w.pairs, w.res, kp_cap, counts, p->checkOrientation, p->HISTO_LENGTH, matches, capacity, out_counts, 1);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// ORB stereo row-band matcher: the block-commented body of LocalBA::AddMapPointsByStereo
// (src/mapping/LocalBA.cpp:69-288).  One CTA per (left, right) frame pair.  The reference's row table lists a right key
// under every row of [floor(y-r), ceil(y+r)]; here each right key is counting-sorted ONCE by floor(y) in shared memory
// with its row interval packed beside it, and a left key on row v scans the bins [v-R, v+R] (one contiguous range of the
// sorted list) testing the interval — same candidate set, one entry per key instead of up to 2r+2.  First-minimum
// (`dist < best`, candidates in index order) = minimum of the key dist<<16 | iR, so the scan order is free.
struct StereoSf { float inv[TB_MAX_LEVELS]; };

__global__ void __launch_bounds__(512)
stereo_rowband_batched(const ProjPair *__restrict__ pairs, const tb_keypoint *__restrict__ kps, const uint8_t *__restrict__ desc,
                       int kp_cap, const int *__restrict__ counts, StereoSf sf, int n_levels, int n_rows, int R, tb_stereo_params p,
                       tb_match *__restrict__ matches, int capacity, int *__restrict__ out_counts, float *__restrict__ depth)
{
    extern __shared__ __align__(16) unsigned char st_smem[];
    int *start = reinterpret_cast<int *>(st_smem);                 // [n_rows + 1] bin starts (after the scan)
    int *cur = start + (n_rows + 1);                               // [n_rows + 1] counts, then fill cursors
    uint2 *ent = reinterpret_cast<uint2 *>(cur + (n_rows + 1));                               // [kp_cap] (2(n_rows+1) ints before it: 8-byte aligned)
    __shared__ int warp_sum[16];
    __shared__ int s_base;
    const ProjPair pr = pairs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nL = min(counts[pr.f1], kp_cap), nR = min(counts[pr.f2], kp_cap);
    const tb_keypoint *kL = kps + (size_t)pr.f1 * kp_cap, *kR = kps + (size_t)pr.f2 * kp_cap;
    const uint8_t *dL = desc + (size_t)pr.f1 * kp_cap * 32, *dR = desc + (size_t)pr.f2 * kp_cap * 32;

    for (int i = tid; i <= n_rows; i += 512) cur[i] = 0;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i = tid; i < nR; i += 512) {
        const float y = __ldg(&kR[i].y);
        const int b = min(max((int)floorf(y), 0), n_rows - 1);
        atomicAdd(&cur[b], 1);
    }
    __syncthreads();
    // exclusive scan of cur[0..n_rows) -> start[0..n_rows]: a serial chunk per thread, warp scans over the chunk sums
    {
        const int per = (n_rows + 511) / 512, b0 = min(tid * per, n_rows), b1 = min(b0 + per, n_rows);
        int sum = 0;
        for (int i = b0; i < b1; i++) sum += cur[i];
        int inc = sum;
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) warp_sum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int v = lane < 16 ? warp_sum[lane] : 0, w = v;
            for (int o = 1; o < 16; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
            if (lane < 16) warp_sum[lane] = w - v;
        }
        __syncthreads();
        int run = warp_sum[wid] + inc - sum;
        for (int i = b0; i < b1; i++) { const int c = cur[i]; start[i] = run; cur[i] = run; run += c; }
        if (tid == 511) start[n_rows] = run;   // thread 511 owns the tail (b1 == n_rows), run == nR
    }
    __syncthreads();
    for (int i = tid; i < nR; i += 512) {
        const float y = __ldg(&kR[i].y), x = __ldg(&kR[i].x);
        const int oc = __ldg(&kR[i].octave);
        const float r = __fmul_rn(p.band, sf.inv[min(max(oc, 0), n_levels - 1)]);
        int max_r = (int)ceilf(__fadd_rn(y, r)), min_r = (int)floorf(__fsub_rn(y, r));   // LocalBA.cpp:87-89
        min_r = max(min_r, 0); max_r = min(max_r, n_rows - 1);
        const int b = min(max((int)floorf(y), 0), n_rows - 1);
        const int ok = (x >= p.min_u && x <= p.max_u && min_r <= max_r) ? 1 : 0;         // :136
        const int slot = atomicAdd(&cur[b], 1);
        ent[slot] = make_uint2((uint32_t)i | ((uint32_t)(oc & 0x7f) << 16) | ((uint32_t)ok << 24) | (oc < 0 || oc > 127 ? 0x80000000u : 0u),
                               (uint32_t)max(min_r, 0) | ((uint32_t)max(max_r, 0) << 16));
    }
    __syncthreads();

    tb_match *out = matches + (size_t)blockIdx.x * capacity;
    float *dep = depth ? depth + (size_t)blockIdx.x * kp_cap : nullptr;
    for (int i0 = 0; i0 < nL; i0 += 512) {
        const int iL = i0 + tid;
        uint32_t best = 0xFFFFFFFFu;
        float uL = 0.f;
        if (iL < nL) {
            const float vL = __ldg(&kL[iL].y);
            uL = __ldg(&kL[iL].x);
            const int levelL = __ldg(&kL[iL].octave);
            if (vL > -1.f && vL < (float)n_rows) {
                const int row = (int)vL;                                                   // (size_t)vL, :110
                const int e0 = start[max(row - R, 0)], e1 = start[min(row + R, n_rows - 1) + 1];
                uint32_t q[8];
                {
                    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(dL + (size_t)iL * 32)), b = __ldg(reinterpret_cast<const uint4 *>(dL + (size_t)iL * 32) + 1);
                    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
                }
                for (int e = e0; e < e1; e++) {
                    const uint2 en = ent[e];
                    const int lo = en.y & 0xffff, hi = en.y >> 16;
                    if (row < lo || row > hi || !(en.x & 0x01000000u)) continue;
                    if (!(en.x & 0x80000000u)) {
                        const int oc = (en.x >> 16) & 0x7f;
                        if (oc < levelL - p.octave_tol || oc > levelL + p.octave_tol) continue;   // :131
                    } else {
                        const int oc = __ldg(&kR[en.x & 0xffff].octave);
                        if (oc < levelL - p.octave_tol || oc > levelL + p.octave_tol) continue;
                    }
                    const int iR = en.x & 0xffff;
                    const int d = ham256_g(q, dR + (size_t)iR * 32);
                    if (d > 0) best = min(best, ((uint32_t)d << 16) | (uint32_t)iR);       // :140
                }
            }
        }
        const int hit = (iL < nL) && (int)(best >> 16) < p.th_dist;                         // :148
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_sum[wid] = __popc(bal);
        __syncthreads();
        int off = s_base + __popc(bal & ((1u << lane) - 1));
        for (int w = 0; w < wid; w++) off += warp_sum[w];
        if (iL < nL) {
            float dz = -1.0f;
            if (hit) {
                const int iR = best & 0xffff;
                const float dx = fabsf(__fsub_rn(uL, __ldg(&kR[iR].x)));                    // :168-171
                if (dx > 0.f) dz = __fdiv_rn(p.bf, dx);
                if (off < capacity) { tb_match mt = {iL, iR, -1, (float)(best >> 16)}; out[off] = mt; }
            }
            if (dep) dep[iL] = dz;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 16; w++) t += warp_sum[w]; s_base += t; }
        __syncthreads();
    }
    if (dep) for (int i = nL + tid; i < kp_cap; i += 512) dep[i] = -1.0f;
    if (tid == 0) out_counts[blockIdx.x] = s_base;
}

int tb_window_stereo_batched(tb_ctx *ctx, const tb_stereo_params *p, int n_pairs, const int32_t *left_frames, const int32_t *right_frames,
                             const tb_keypoint *kps, const uint8_t *desc, int kp_cap, const int *counts, const float *inv_sf, int n_levels,
                             int n_rows, tb_match *matches, int capacity, int *out_counts, float *depth)
{
    if (n_pairs <= 0) return TB_OK;
    if (kp_cap > 65535) return tb_fail(ctx, TB_E_UNSUPPORTED, "stereo row-band matcher supports up to 65535 keypoints per frame");
    if (n_levels < 1 || n_levels > TB_MAX_LEVELS || n_rows < 1 || n_rows > 65535) return tb_fail(ctx, TB_E_UNSUPPORTED, "stereo row-band matcher: 1..16 levels, 1..65535 rows");
    if (!(p->band >= 0.f) || p->th_dist > 257) return tb_fail(ctx, TB_E_INVALID, "stereo row-band matcher: band >= 0, th_dist <= 257");
    StereoSf sf;
    float rmax = 0.f;
    for (int l = 0; l < n_levels; l++) { sf.inv[l] = inv_sf[l]; volatile float r = p->band * inv_sf[l]; if (r > rmax) rmax = r; }
    if (!(rmax < 32768.f)) return tb_fail(ctx, TB_E_INVALID, "stereo row-band matcher: band * inverse scale factor out of range");
    const int R = (int)ceilf(rmax) + 1;
    const size_t smem = sizeof(int) * (2 * (size_t)(n_rows + 1) + 1) + sizeof(uint2) * (size_t)kp_cap;
    if (smem > 200 * 1024) return tb_fail(ctx, TB_E_UNSUPPORTED, "stereo row-band matcher: rows*8 + keypoint capacity*8 must fit 200 KB of shared memory");
    int rc = win_ws_reserve(ctx, 0, n_pairs, kp_cap, false);
    if (rc) return rc;
    WinWs &w = *ctx->wws;
    std::vector<ProjPair> hp(n_pairs);
    memset(hp.data(), 0, sizeof(ProjPair) * n_pairs);
    for (int i = 0; i < n_pairs; i++) { hp[i].f1 = left_frames[i]; hp[i].f2 = right_frames[i]; }
    if ((rc = tb_stage_upload(ctx, w.pairs, hp.data(), sizeof(ProjPair) * n_pairs))) return rc;   // through the pinned ring: no host wait
    {
        ProfScope prof__(ctx, "stereo_rowband_batched");
        stereo_rowband_batched<<<n_pairs, 512, smem, ctx->stream>>>(w.pairs, kps, desc, kp_cap, counts, sf, n_levels, n_rows, R, *p,
                                                                     matches, capacity, out_counts, depth);
        TB_LAUNCH_CHECK(ctx);
    }
    return TB_OK;
}

void tb_window_init_device() { cudaFuncSetAttribute(stereo_rowband_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }   // per device

// Host-buffer form: left keys/descriptors become frame 0, right ones frame 1 of a temporary two-frame layout.
extern "C" int tb_stereo_match(tb_ctx *ctx, const tb_stereo_params *p, const tb_keypoint *kL, const uint8_t *dL, int nL,
                               const tb_keypoint *kR, const uint8_t *dR, int nR, const float *inv_sf, int n_levels, int n_rows,
                               tb_match *out, int capacity, int *n, float *depth)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || nL < 0 || nR < 0 || (nL && (!kL || !dL)) || (nR && (!kR || !dR)) || !inv_sf || !n || capacity < 0 || (capacity && !out)) return TB_E_INVALID;
    *n = 0;
    if (nL == 0) return TB_OK;
    const int cap = std::max(std::max(nL, nR), 1);
    ArenaScope buf(ctx);
    tb_keypoint *d_k = buf.alloc<tb_keypoint>(2 * (size_t)cap);
    uint8_t *d_d = buf.alloc<uint8_t>(64 * (size_t)cap);
    int *d_cnt = buf.alloc<int>(3);
    tb_match *d_m = buf.alloc<tb_match>(nL);
    float *d_z = buf.alloc<float>(cap);
    if (!d_k || !d_d || !d_cnt || !d_m || !d_z) return tb_fail(ctx, TB_E_NOMEM, "cudaMalloc failed");
    const int cnt[2] = {nL, nR};
    TB_CUDA(ctx, cudaMemcpyAsync(d_k, kL, sizeof(tb_keypoint) * nL, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(ctx, cudaMemcpyAsync(d_d, dL, 32 * (size_t)nL, cudaMemcpyHostToDevice, ctx->stream));
    if (nR) {
        TB_CUDA(ctx, cudaMemcpyAsync(d_k + cap, kR, sizeof(tb_keypoint) * nR, cudaMemcpyHostToDevice, ctx->stream));
        TB_CUDA(ctx, cudaMemcpyAsync(d_d + 32 * (size_t)cap, dR, 32 * (size_t)nR, cudaMemcpyHostToDevice, ctx->stream));
    }
    TB_CUDA(ctx, cudaMemcpyAsync(d_cnt, cnt, sizeof(cnt), cudaMemcpyHostToDevice, ctx->stream));
    const int32_t f1 = 0, f2 = 1;
    int rc = tb_window_stereo_batched(ctx, p, 1, &f1, &f2, d_k, d_d, cap, d_cnt, inv_sf, n_levels, n_rows, d_m, nL, d_cnt + 2, d_z);
    if (rc) return rc;
    int m = 0;
    TB_CUDA(ctx, cudaMemcpyAsync(&m, d_cnt + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n = m;
    const int take = std::min(m, capacity);
    if (take > 0) TB_CUDA(ctx, cudaMemcpy(out, d_m, sizeof(tb_match) * take, cudaMemcpyDeviceToHost));
    if (depth) TB_CUDA(ctx, cudaMemcpy(depth, d_z, sizeof(float) * nL, cudaMemcpyDeviceToHost));
    return m > capacity ? tb_fail(ctx, TB_E_CAPACITY, "stereo match buffer too small") : TB_OK;
}

void tb_window_ws_free(tb_ctx *ctx)
{
    if (!ctx->wws) return;
    cudaFree(ctx->wws->cell_start); cudaFree(ctx->wws->items); cudaFree(ctx->wws->res); cudaFree(ctx->wws->pairs);
    delete ctx->wws;
    ctx->wws = nullptr;
}
