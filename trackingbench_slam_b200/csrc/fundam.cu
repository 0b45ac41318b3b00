// fundam.cu — Matcher::rejectWithF (reference src/matchers/matcher.cpp:853-881): the tracks with status != 0 go through
// cv::findFundamentalMat(cur, last, cv::FM_RANSAC, 1.0, 0.99, mask) and the outliers' status is cleared.  Every reference
// call site of searchByOPFlow passes reject = true (test_vo.cpp:716, LocalBA.cpp:54).
//
// OpenCV's RANSAC (calib3d fundam.cpp / ptsetreg.cpp, 4.x) is sequential only in its bookkeeping: the sample of iteration i
// comes from cv::RNG((uint64)-1) and never depends on an earlier model, so
//   * one thread replays the generator and getSubset (duplicate test, collinearity test, redraws) for a chunk of 32
//     iterations — the points of the pair sit in shared memory;
//   * 32 threads solve one 7-point problem each (Hartley normalisation, null space of the 7x9 system, det = 0 cubic, up to
//     three models) in double precision;
//   * all warps count the inliers of the chunk's <= 96 models (symmetric epipolar error in double, rounded to float and
//     compared with threshold^2 as OpenCV does);
//   * one thread replays OpenCV's loop over the chunk: best-so-far model, RANSACUpdateNumIters, stop at niters.
// One CTA per frame pair, no host round trip; the models of the iterations past the stopping point are simply dropped.
// The null space comes from a Gram-Schmidt orthogonalisation of the seven rows (in registers) instead of OpenCV's SVD: the
// candidate models are the members of the pencil spanned by the two null vectors with det = 0, which do not depend on the basis.
// Inlier masks are identical to cv2 4.13 on every case tested (tests/test_gpu_flow.py, tests/test_oracle_cv2.py).
#include "common.cuh"
#include "kernels.cuh"
#include <cfloat>

#define RF_THREADS 256
#define RF_CHUNK 32          // RANSAC iterations per round
#define RF_SMEM_PTS 6144     // valid tracks held in shared memory (more: they stay in the global scratch)

struct RfPair {
    const float2 *cur;       // tracked points (findFundamentalMat's first set)
    const uint8_t *last;     // points they were tracked from: float2 at last + i * last_stride
    int last_stride;
    const int32_t *count;    // number of points (device)
    uint8_t *status;         // in / out
};
struct RfWs {
    RfPair *pairs = nullptr; int pairs_cap = 0;
    float2 *m = nullptr; int32_t *id = nullptr; size_t cap = 0;   // [pair][2][pts_cap] points, [pair][pts_cap] indices
};

__device__ __forceinline__ unsigned rf_rng_next(unsigned long long &st)
{
    st = (unsigned long long)(unsigned)st * 4164903690ull + (unsigned)(st >> 32);
    return (unsigned)st;
}
// haveCollinearPoints(m, 7): the last point against every pair of earlier ones
__device__ __noinline__ bool rf_collinear7(const float2 *p)
{
    for (int j = 0; j < 6; j++) {
        const double dx1 = (double)p[j].x - (double)p[6].x, dy1 = (double)p[j].y - (double)p[6].y;
        for (int k = 0; k < j; k++) {
            const double dx2 = (double)p[k].x - (double)p[6].x, dy2 = (double)p[k].y - (double)p[6].y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}
// cv::solveCubic for c[0] x^3 + c[1] x^2 + c[2] x + c[3]
__device__ __noinline__ int rf_solve_cubic(const double *c, double *x)
{
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    if (a0 == 0) {
        if (a1 == 0) { if (a2 == 0) return 0; x[0] = -a3 / a2; return 1; }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0) return 0;
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; } else { x[0] = q2 / a1; x[1] = a3 / q2; }
        return d > 0 ? 2 : 1;
    }
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9), R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qcubed)), t0 = -2 * sqrt(Q), t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * cos(t1) - t2;
        x[1] = t0 * cos(t1 + (2. * 3.1415926535897932384626433832795 / 3)) - t2;
        x[2] = t0 * cos(t1 + (4. * 3.1415926535897932384626433832795 / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
        else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
        return x[0] == x[1] ? 1 : 2;
    }
    d = sqrt(-d);
    double e = pow(d + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}
// run7Point (fundam.cpp): up to three models of 9 doubles for the 7 correspondences m1[i] <-> m2[i]
__device__ __noinline__ int rf_run7(const float2 *m1, const float2 *m2, double *F)
{
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0, scale1 = 0, scale2 = 0;
    for (int i = 0; i < 7; i++) { c1x += m1[i].x; c1y += m1[i].y; c2x += m2[i].x; c2y += m2[i].y; }
    const double t = 1. / 7;
    c1x *= t; c1y *= t; c2x *= t; c2y *= t;
    for (int i = 0; i < 7; i++) {
        double dx = m1[i].x - c1x, dy = m1[i].y - c1y; scale1 += sqrt(dx * dx + dy * dy);
        dx = m2[i].x - c2x; dy = m2[i].y - c2y; scale2 += sqrt(dx * dx + dy * dy);
    }
    scale1 *= t; scale2 *= t;
    if (scale1 < FLT_EPSILON || scale2 < FLT_EPSILON) return 0;
    scale1 = sqrt(2.) / scale1; scale2 = sqrt(2.) / scale2;
    // rows of the 7 x 9 system (m2[i], 1)^T F (m1[i], 1) = 0 on the normalised points, orthonormalised in registers (modified
    // Gram-Schmidt, every row twice): all loops are unrolled and no array is indexed dynamically, so nothing lives in local
    // memory (a Jacobi sweep over A^T A there cost milliseconds per round of 32 hypotheses)
    double A[7][9];
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const double x0 = (m1[i].x - c1x) * scale1, y0 = (m1[i].y - c1y) * scale1;
        const double x1 = (m2[i].x - c2x) * scale2, y1 = (m2[i].y - c2y) * scale2;
        A[i][0] = x1 * x0; A[i][1] = x1 * y0; A[i][2] = x1; A[i][3] = y1 * x0; A[i][4] = y1 * y0; A[i][5] = y1; A[i][6] = x0; A[i][7] = y0; A[i][8] = 1;
    }
#pragma unroll
    for (int i = 0; i < 7; i++) {
#pragma unroll
        for (int pass = 0; pass < 2; pass++)
#pragma unroll
            for (int j = 0; j < i; j++) {
                double d = 0;
#pragma unroll
                for (int k = 0; k < 9; k++) d += A[j][k] * A[i][k];
#pragma unroll
                for (int k = 0; k < 9; k++) A[i][k] -= d * A[j][k];
            }
        double nn = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) nn += A[i][k] * A[i][k];
        const double inv = nn > 1e-280 ? 1. / sqrt(nn) : 0.;     // a dependent row drops out (the null space is then larger than 2)
#pragma unroll
        for (int k = 0; k < 9; k++) A[i][k] *= inv;
    }
    // two orthonormal vectors of the complement: project the unit vector that sticks out most, twice
    double f1[9], f2[9], c[4], r[3] = {0, 0, 0};
#pragma unroll
    for (int which = 0; which < 2; which++) {
        double res[9];
#pragma unroll
        for (int k = 0; k < 9; k++) {
            double q = 1.;
#pragma unroll
            for (int i = 0; i < 7; i++) q -= A[i][k] * A[i][k];
            if (which == 1) q -= f2[k] * f2[k];
            res[k] = q;
        }
        int km = 0;
        double best = res[0];
#pragma unroll
        for (int k = 1; k < 9; k++) if (res[k] > best) { best = res[k]; km = k; }
        double v[9];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = k == km ? 1. : 0.;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
#pragma unroll
            for (int i = 0; i < 7; i++) {
                double d = 0;
#pragma unroll
                for (int k = 0; k < 9; k++) d += A[i][k] * v[k];
#pragma unroll
                for (int k = 0; k < 9; k++) v[k] -= d * A[i][k];
            }
            if (which == 1) {
                double d = 0;
#pragma unroll
                for (int k = 0; k < 9; k++) d += f2[k] * v[k];
#pragma unroll
                for (int k = 0; k < 9; k++) v[k] -= d * f2[k];
            }
        }
        double nn = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) nn += v[k] * v[k];
        const double inv = 1. / sqrt(nn);
#pragma unroll
        for (int k = 0; k < 9; k++) { if (which == 0) f2[k] = v[k] * inv; else f1[k] = v[k] * inv; }
    }
    // det(lambda * f1 + (1 - lambda) * f2) = 0
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    const int n = rf_solve_cubic(c, r);
    if (n < 1 || n > 3) return 0;
    const double T1[9] = {scale1, 0, -scale1 * c1x, 0, scale1, -scale1 * c1y, 0, 0, 1};
    const double T2[9] = {scale2, 0, -scale2 * c2x, 0, scale2, -scale2 * c2y, 0, 0, 1};
    for (int k = 0; k < n; k++, F += 9) {
        double lambda = r[k], mu = 1., G[9], H[9];
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; G[8] = 1.; } else G[8] = 0.;
        for (int i = 0; i < 8; i++) G[i] = f1[i] * lambda + f2[i] * mu;
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double v = 0; for (int m = 0; m < 3; m++) v += T2[m * 3 + a] * G[m * 3 + b]; H[a * 3 + b] = v; }
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double v = 0; for (int m = 0; m < 3; m++) v += H[a * 3 + m] * T1[m * 3 + b]; F[a * 3 + b] = v; }
        if (fabs(F[8]) > FLT_EPSILON) { const double inv = 1. / F[8]; for (int i = 0; i < 9; i++) F[i] *= inv; }
    }
    return n;
}
// FMEstimatorCallback::computeError of one correspondence, rounded to float as OpenCV stores it
__device__ __forceinline__ float rf_error(const double *F, float2 p1, float2 p2)
{
    const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
    const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
    a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
    const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
    const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    return (float)(e1 > e2 ? e1 : e2);
}
__device__ __noinline__ int rf_update_iters(double p, double ep, int max_iters)
{
    p = p < 0 ? 0 : p > 1 ? 1 : p; ep = ep < 0 ? 0 : ep > 1 ? 1 : ep;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN, denom = 1. - pow(1. - ep, 7.);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

extern __shared__ __align__(16) uint8_t rf_smem[];

__global__ void __launch_bounds__(RF_THREADS, 1)
reject_with_f(const RfPair *__restrict__ pairs, float thr2, double confidence, int max_iters, int pts_cap, float2 *__restrict__ scratch_m,
              int32_t *__restrict__ scratch_id, double *__restrict__ F_out, int32_t *__restrict__ iters_out)
{
    const RfPair pd = pairs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(*pd.count, pts_cap);
    __shared__ int s_scan[RF_THREADS / 32], s_done, s_niters, s_iter, s_maxgood, s_nmodels[RF_CHUNK], s_good[RF_CHUNK * 3], s_nhyp;
    __shared__ unsigned long long s_rng;
    __shared__ double s_best[9], s_models[RF_CHUNK * 27];
    __shared__ float2 s_sub[RF_CHUNK][14];   // the 7 + 7 points of every hypothesis of the chunk
    if (iters_out && tid == 0) iters_out[blockIdx.x] = 0;
    if (n <= 8) return;                                          // matcher.cpp:870: findFundamentalMat is not called
    // ---- compaction of the tracks with status != 0, order preserved (matcher.cpp:859-867)
    float2 *gm1 = scratch_m + (size_t)blockIdx.x * 2 * pts_cap, *gm2 = gm1 + pts_cap;
    int32_t *gid = scratch_id + (size_t)blockIdx.x * pts_cap;
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += RF_THREADS) {
        const int i = i0 + tid;
        const bool v = i < n && pd.status[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; w++) off += s_scan[w];
        if (v) {
            const int pos = off + __popc(bal & ((1u << lane) - 1));
            gm1[pos] = pd.cur[i];
            const float *lp = reinterpret_cast<const float *>(pd.last + (size_t)i * pd.last_stride);   // 28-byte keypoints: 4-byte aligned only
            gm2[pos] = make_float2(lp[0], lp[1]);
            gid[pos] = i;
        }
        for (int w = 0; w < RF_THREADS / 32; w++) base += s_scan[w];
        __syncthreads();
    }
    const int count = base;
    if (count < 15) return;                                      // OpenCV runs LMedS (or nothing) below 15 points: status unchanged
    // points into shared memory when they fit (the sampler's random reads and the inlier sweeps go there)
    const bool in_smem = count <= RF_SMEM_PTS;
    float2 *m1 = in_smem ? reinterpret_cast<float2 *>(rf_smem) : gm1, *m2 = in_smem ? m1 + count : gm2;
    if (in_smem) for (int i = tid; i < count; i += RF_THREADS) { m1[i] = gm1[i]; m2[i] = gm2[i]; }
    if (tid == 0) { s_rng = 0xffffffffffffffffull; s_niters = max_iters > 1 ? max_iters : 1; s_iter = 0; s_maxgood = 0; s_done = 0; }
    __syncthreads();
    while (true) {
        // ---- one thread: getSubset for the next RF_CHUNK iterations (the generator never depends on a model)
        if (tid == 0) {
            unsigned long long st = s_rng;
            int nh = 0;
            for (; nh < RF_CHUNK; nh++) {
                bool found = false;
                for (int attempt = 0; attempt < 10000 && !found; attempt++) {
                    int idx[7];
                    for (int i = 0; i < 7; i++) {
                        int v; bool dup;
                        do {
                            v = (int)(rf_rng_next(st) % (unsigned)count);
                            dup = false;
                            for (int k = 0; k < i; k++) dup |= idx[k] == v;
                        } while (dup);
                        idx[i] = v;
                        s_sub[nh][i] = m1[v]; s_sub[nh][7 + i] = m2[v];
                    }
                    found = !rf_collinear7(s_sub[nh]) && !rf_collinear7(s_sub[nh] + 7);
                }
                if (!found) break;                               // getSubset failed: the loop ends here (ptsetreg.cpp)
            }
            s_rng = st; s_nhyp = nh;
        }
        __syncthreads();
        const int nh = s_nhyp;
        // ---- 32 threads: one 7-point problem each
        if (tid < nh) s_nmodels[tid] = rf_run7(s_sub[tid], s_sub[tid] + 7, s_models + 27 * tid);
        __syncthreads();
        // ---- all warps: inlier counts of the chunk's models
        for (int m = warp; m < nh * 3; m += RF_THREADS / 32) {
            const int h = m / 3, k = m - 3 * h;
            if (k >= s_nmodels[h]) continue;                     // warp-uniform
            const double *F = s_models + 27 * h + 9 * k;
            int good = 0;
            for (int i = lane; i < count; i += 32) good += rf_error(F, m1[i], m2[i]) <= thr2;
            good = __reduce_add_sync(0xffffffffu, good);
            if (lane == 0) s_good[m] = good;
        }
        __syncthreads();
        // ---- one thread: OpenCV's loop over the chunk
        if (tid == 0) {
            int iter = s_iter, niters = s_niters, maxgood = s_maxgood;
            int h = 0;
            for (; h < nh && iter < niters; h++, iter++)
                for (int k = 0; k < s_nmodels[h]; k++) {
                    const int good = s_good[3 * h + k];
                    if (good > max(maxgood, 6)) {
                        for (int i = 0; i < 9; i++) s_best[i] = s_models[27 * h + 9 * k + i];
                        maxgood = good;
                        niters = rf_update_iters(confidence, (double)(count - good) / count, niters);
                    }
                }
            s_iter = iter; s_niters = niters; s_maxgood = maxgood;
            s_done = iter >= niters || nh < RF_CHUNK;
        }
        __syncthreads();
        if (s_done) break;
    }
    // ---- outliers of the best model lose their status (matcher.cpp:874-880)
    if (s_maxgood > 0) {
        for (int i = tid; i < count; i += RF_THREADS)
            if (!(rf_error(s_best, m1[i], m2[i]) <= thr2)) pd.status[gid[i]] = 0;
        if (F_out && tid < 9) F_out[(size_t)blockIdx.x * 9 + tid] = s_best[tid];
    }
    if (iters_out && tid == 0) iters_out[blockIdx.x] = s_iter;
}

void tb_fund_ws_free(tb_ctx *ctx)
{
    if (!ctx->rfws) return;
    RfWs &W = *ctx->rfws;
    if (W.pairs) cudaFree(W.pairs);
    if (W.m) cudaFree(W.m);
    if (W.id) cudaFree(W.id);
    delete ctx->rfws;
    ctx->rfws = nullptr;
}
void tb_fund_init_device()
{
    cudaFuncSetAttribute(reject_with_f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(float2) * RF_SMEM_PTS));
}

// pairs described on the host (device pointers inside), launched without a host wait
int tb_reject_with_f_batched(tb_ctx *ctx, const tb_ransac_params *p, int n_pairs, const std::vector<RfPairHost> &hp, int pts_cap,
                             double *F_out, int32_t *iters_out)
{
    if (n_pairs == 0) return TB_OK;
    if (!ctx->rfws) ctx->rfws = new RfWs();
    RfWs &W = *ctx->rfws;
    if (n_pairs > W.pairs_cap) { W.pairs_cap = 0; int rc = tb_regrow(ctx, W.pairs, (size_t)n_pairs); if (rc) return rc; W.pairs_cap = n_pairs; }
    const size_t need = (size_t)n_pairs * pts_cap;
    if (need > W.cap) {
        W.cap = 0;
        int rc = tb_regrow(ctx, W.m, 2 * need); if (rc) return rc;
        rc = tb_regrow(ctx, W.id, need); if (rc) return rc;
        W.cap = need;
    }
    std::vector<RfPair> t(n_pairs);
    for (int i = 0; i < n_pairs; i++) { t[i].cur = (const float2 *)hp[i].cur; t[i].last = (const uint8_t *)hp[i].last; t[i].last_stride = hp[i].last_stride; t[i].count = hp[i].count; t[i].status = hp[i].status; }
    int rc = tb_stage_upload(ctx, W.pairs, t.data(), sizeof(RfPair) * n_pairs);
    if (rc) return rc;
    float thr = p->threshold > 0 ? p->threshold : 3.f;
    double conf = p->confidence;
    if (conf < DBL_EPSILON || conf > 1 - DBL_EPSILON) conf = 0.99;
    const size_t smem = 2 * sizeof(float2) * (size_t)std::min(pts_cap, RF_SMEM_PTS);
    ProfScope prof__(ctx, "reject_with_f");
    reject_with_f<<<n_pairs, RF_THREADS, smem, ctx->stream>>>(W.pairs, (float)((double)thr * (double)thr), conf, p->max_iters > 0 ? p->max_iters : 1000,
                                                              pts_cap, W.m, W.id, F_out, iters_out);
    TB_LAUNCH_CHECK(ctx);
    return TB_OK;
}

extern "C" int tb_reject_with_f_device(tb_ctx *ctx, const tb_ransac_params *p, int n_pairs, const void *cur_pts, const void *last_pts,
                                       const void *counts, int pts_cap, void *status, void *F_out, void *iters_out)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || n_pairs < 0 || pts_cap < 0 || (n_pairs && (!cur_pts || !last_pts || !counts || !status)))
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_reject_with_f_device: bad arguments") : TB_E_INVALID;
    std::vector<RfPairHost> hp(n_pairs);
    for (int i = 0; i < n_pairs; i++) {
        hp[i].cur = (const float *)cur_pts + 2 * (size_t)i * pts_cap;
        hp[i].last = (const uint8_t *)last_pts + sizeof(float2) * (size_t)i * pts_cap;
        hp[i].last_stride = (int)sizeof(float2);
        hp[i].count = (const int32_t *)counts + i;
        hp[i].status = (uint8_t *)status + (size_t)i * pts_cap;
    }
    return tb_reject_with_f_batched(ctx, p, n_pairs, hp, pts_cap, (double *)F_out, (int32_t *)iters_out);
}

extern "C" int tb_reject_with_f_host(tb_ctx *ctx, const tb_ransac_params *p, const float *cur_xy, const float *last_xy, int n, uint8_t *status,
                                     double *F9, int *iterations)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !p || n < 0 || (n && (!cur_xy || !last_xy || !status)))
        return ctx ? tb_fail(ctx, TB_E_INVALID, "tb_reject_with_f_host: bad arguments") : TB_E_INVALID;
    if (iterations) *iterations = 0;
    if (n == 0) return TB_OK;
    ArenaScope buf(ctx);
    float2 *d_cur = buf.upload((const float2 *)cur_xy, n), *d_last = buf.upload((const float2 *)last_xy, n);
    uint8_t *d_st = buf.upload(status, n);
    int32_t *d_cnt = buf.upload(&n, 1), *d_it = buf.alloc<int32_t>(1);
    double *d_F = buf.alloc<double>(9);
    if (!d_cur || !d_last || !d_st || !d_cnt || !d_it || !d_F) return tb_fail(ctx, TB_E_CUDA, "tb_reject_with_f_host: out of device memory");
    TB_CUDA(ctx, cudaMemsetAsync(d_F, 0, sizeof(double) * 9, ctx->stream));
    int rc = tb_reject_with_f_device(ctx, p, 1, d_cur, d_last, d_cnt, n, d_st, d_F, d_it);
    if (rc) return rc;
    int it = 0;
    TB_CUDA(ctx, cudaMemcpyAsync(status, d_st, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (F9) TB_CUDA(ctx, cudaMemcpyAsync(F9, d_F, sizeof(double) * 9, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaMemcpyAsync(&it, d_it, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (iterations) *iterations = it;
    return TB_OK;
}
