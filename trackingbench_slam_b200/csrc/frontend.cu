// frontend.cu — context + tb_frontend: owns the device-resident batch (pyramid, blurred pyramid,
// FAST candidates, octree scratch, keypoints/descriptors) and sequences the kernels.
// Host code here is bookkeeping only (sizes, budgets, launches); all image / descriptor
// arithmetic runs in the CUDA kernels of pyramid.cu, fast.cu, octree.cu, describe.cu, match.cu.
#include "common.cuh"
#include "kernels.cuh"
#include <cmath>
#include <cstdlib>
#include <algorithm>

// ------------------------------------------------------------------------------------------ ctx
extern "C" const char *tb_version(void) { return "trackingbench_slam_b200 0.1 (sm_100a)"; }

extern "C" int tb_ctx_create(int device, void *stream, tb_ctx **out)
{
    if (!out) return TB_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return TB_E_CUDA; // no CPU fallback: fail loudly
    if (device < 0 || device >= ndev) return TB_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return TB_E_CUDA;
    tb_ctx *c = new tb_ctx();
    c->device = device;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return TB_E_CUDA; }
        c->own_stream = true;
    }
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool ev_ok = true;
    for (int l = 0; l < TB_MAX_LEVELS; l++) ev_ok = ev_ok && cudaEventCreateWithFlags(&c->ev_lvl[l], cudaEventDisableTiming) == cudaSuccess;
    for (int l = 0; l < TB_MAX_LEVELS; l++) ev_ok = ev_ok && cudaEventCreateWithFlags(&c->ev_pyr[l], cudaEventDisableTiming) == cudaSuccess;
    const int side_prio = getenv("TB_GAUSS_PRIO") ? (atoi(getenv("TB_GAUSS_PRIO")) > 0 ? prio_hi : prio_lo) : 0;   // tuning knob (default priority measured best)
    if (!ev_ok || cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, side_prio) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->side2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||   // its few CTAs go first
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming) != cudaSuccess) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return TB_E_CUDA;
    }
    tb_fast_init_device(); tb_fund_init_device(); tb_octree_init_device(); tb_pyramid_init_device(); tb_match_tc_init_device(); tb_describe_init_device();
    tb_window_init_device(); tb_flow_init_device();
    *out = c;
    return TB_OK;
}
extern "C" void tb_ctx_destroy(tb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    tb_match_ws_free(ctx);
    tb_match_tc_free(ctx);
    tb_window_ws_free(ctx);
    tb_ingest_ws_free(ctx);
    tb_flow_ws_free(ctx);
    tb_fund_ws_free(ctx);
    ctx->arena.destroy();
    ctx->stage.destroy();
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->side2) cudaStreamDestroy(ctx->side2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    for (int l = 0; l < TB_MAX_LEVELS; l++) if (ctx->ev_lvl[l]) cudaEventDestroy(ctx->ev_lvl[l]);
    for (int l = 0; l < TB_MAX_LEVELS; l++) if (ctx->ev_pyr[l]) cudaEventDestroy(ctx->ev_pyr[l]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" int tb_sync(tb_ctx *ctx)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx) return TB_E_INVALID;
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}
extern "C" const char *tb_last_error(const tb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" int64_t tb_launch_count(const tb_ctx *ctx) { return ctx ? ctx->launches : 0; }

// 0 = LOP3+POPC kernel (K7), 1 = tcgen05 int8 +-1 formulation (K8) for the batched brute-force matcher.
extern "C" int tb_set_match_impl(tb_ctx *ctx, int impl)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || (impl != 0 && impl != 1)) return TB_E_INVALID;
    ctx->match_impl = impl;
    return TB_OK;
}

extern "C" int tb_profile_enable(tb_ctx *ctx, int on)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx) return TB_E_INVALID;
    ctx->profiling = on != 0;
    return TB_OK;
}
// Synchronises, then writes one line per kernel name: "name launches total_ms\n"; clears the log.
extern "C" int tb_profile_report(tb_ctx *ctx, char *buf, size_t cap)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !buf || cap == 0) return TB_E_INVALID;
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<std::string> names; std::vector<int> cnt; std::vector<double> ms;
    for (auto &e : ctx->prof) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e.a, e.b);
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
        size_t k = 0;
        for (; k < names.size(); k++) if (names[k] == e.name) break;
        if (k == names.size()) { names.push_back(e.name); cnt.push_back(0); ms.push_back(0); }
        cnt[k]++; ms[k] += t;
    }
    ctx->prof.clear();
    std::string out;
    char line[256];
    for (size_t k = 0; k < names.size(); k++) { snprintf(line, sizeof line, "%s %d %.6f\n", names[k].c_str(), cnt[k], ms[k]); out += line; }
    snprintf(buf, cap, "%s", out.c_str());
    return TB_OK;
}

// Frame::Frame scale tables + ComputePyramid sizes (reference src/types/Frame.cpp:18-29, :424)
extern "C" int tb_pyramid_geometry(int width, int height, int n_levels, float scale, float *sf, int32_t *lw, int32_t *lh)
{
    if (n_levels <= 0 || n_levels > TB_MAX_LEVELS || width <= 0 || height <= 0) return TB_E_INVALID;
    std::vector<float> s(n_levels, 1.f);
    for (int i = 1; i < n_levels; i++) {
        volatile float v = s[i - 1] * scale; // float32 chain
        s[i] = v;
    }
    for (int i = 0; i < n_levels; i++) {
        if (sf) sf[i] = s[i];
        volatile float fw = (float)width * s[i], fh = (float)height * s[i];
        if (lw) lw[i] = i == 0 ? width : (int)fw;
        if (lh) lh[i] = i == 0 ? height : (int)fh;
    }
    return TB_OK;
}

// ------------------------------------------------------------------------------------------ frontend
struct tb_frontend {
    tb_ctx *ctx = nullptr;
    int L = 0, B = 0, max_target = 0, kp_cap = 0, kept_cap = 0;
    LevelGeom g[TB_MAX_LEVELS];
    uint8_t *pyr[TB_MAX_LEVELS] = {}, *blur[TB_MAX_LEVELS] = {};
    int4 *xtab[TB_MAX_LEVELS] = {}, *ytab[TB_MAX_LEVELS] = {};
    bool area2[TB_MAX_LEVELS] = {};
    int box_pitch[TB_MAX_LEVELS] = {}, box_rows[TB_MAX_LEVELS] = {};   // K1 shared-memory source box
    CUtensorMap pyr_tmap[TB_MAX_LEVELS];                               // K1 TMA: tensor map over level l-1 (source of level l)
    bool has_tmap[TB_MAX_LEVELS] = {};
    bool span8[TB_MAX_LEVELS] = {};                                    // K1: the 8-columns-per-thread kernel applies
    uint32_t *cand[TB_MAX_LEVELS] = {}, *kept[TB_MAX_LEVELS] = {};
    uint16_t *nid[TB_MAX_LEVELS] = {};
    int *n_cand = nullptr, *n_kept = nullptr, *n_kp = nullptr, *overflow = nullptr;
    tb_keypoint *kps = nullptr;
    uint8_t *desc = nullptr;
    float2 *exit_xy = nullptr; int16_t *enid = nullptr; int exit_cap = 0;
    int budget[TB_MAX_LEVELS] = {};
    bool has_budget = false;
    bool levels_exposed = false;   // a raw pointer to a level >= 1 was handed out (tb_frontend_level_layout): the pyramid is then
                                   // always built on the main stream, so that a wait on the caller's stream covers it
    bool blur_valid = false;   // the dense blurred levels hold the Gaussian of the current pyramid (else made on demand)
    uint8_t *raw = nullptr; size_t raw_bytes = 0;   // staging for distorted frames (tb_frontend_upload_frames_undistort)
    // CUDA graphs of the launch sequences of small batches (the single-frame host entries are launch bound: ~50 launches
    // for well under a millisecond of kernels): captured once per parameter set, replayed afterwards
    struct GraphSlot { cudaGraphExec_t exec = nullptr; int n_frames = 0, target = 0; float initTh = 0, minTh = 0; float sf[TB_MAX_LEVELS] = {}; int64_t launches = 0; };
    GraphSlot g_extract, g_pyramid;
    std::vector<void *> allocs;
};

template <typename T> static int dev_alloc(tb_frontend *fe, T **p, size_t count)
{
    void *q = nullptr;
    if (cudaMalloc(&q, count * sizeof(T) + 256) != cudaSuccess) return tb_fail(fe->ctx, TB_E_NOMEM, "cudaMalloc failed");
    fe->allocs.push_back(q);
    *p = (T *)q;
    return TB_OK;
}

static void level_geometry(LevelGeom &g, int w, int h, int B)
{
    // ORBextractor.cpp:743-763 (cell grid) and :498-500 (initial octree nodes)
    g.w = w; g.h = h;
    g.pitch = (int)align_up((size_t)w, 64);
    g.frame_stride = (size_t)g.pitch * h;
    const int minB = 16;
    g.maxBX = w - 16; g.maxBY = h - 16;
    const float width = (float)(g.maxBX - minB), height = (float)(g.maxBY - minB);
    g.nCols = (int)(width / 30.f); g.nRows = (int)(height / 30.f);
    if (g.nCols <= 0 || g.nRows <= 0) { g.nCols = g.nRows = 0; g.wCell = g.hCell = 1; g.nColsRun = g.nRowsRun = 0; g.cand_cap = 64; g.nIni = 0; g.hX = 1.f; return; }
    g.wCell = (int)ceil(width / (float)g.nCols); g.hCell = (int)ceil(height / (float)g.nRows);
    g.nColsRun = 0; for (int j = 0; j < g.nCols; j++) if ((float)(minB + j * g.wCell) < (float)g.maxBX - 6.f) g.nColsRun++;
    g.nRowsRun = 0; for (int i = 0; i < g.nRows; i++) if ((float)minB + (float)i * (float)g.hCell < (float)g.maxBY - 3.f) g.nRowsRun++;
    int wi = w - 38 > 0 ? w - 38 : 0, hi = h - 38 > 0 ? h - 38 : 0;
    g.cand_cap = ((wi + g.nCols + 1) / 2) * ((hi + g.nRows + 1) / 2) + 64; // NMS survivors can never exceed this
    g.nIni = (int)roundf((float)(g.maxBX - minB) / (g.maxBY - minB));
    volatile float hx = (float)(g.maxBX - minB) / g.nIni;
    g.hX = g.nIni > 0 ? hx : 1.f;
    (void)B;
}

extern "C" void tb_frontend_destroy(tb_frontend *fe)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe) return;
    cudaSetDevice(fe->ctx->device);
    cudaStreamSynchronize(fe->ctx->stream);
    if (fe->g_extract.exec) cudaGraphExecDestroy(fe->g_extract.exec);
    if (fe->g_pyramid.exec) cudaGraphExecDestroy(fe->g_pyramid.exec);
    for (void *p : fe->allocs) cudaFree(p);
    cudaFree(fe->raw);
    delete fe;
}

extern "C" int tb_frontend_create(tb_ctx *ctx, int n_levels, const int32_t *lw, const int32_t *lh, int max_batch, int max_target, tb_frontend **out)
{
    TB_DEVICE_GUARD(ctx);
    if (!ctx || !out || !lw || !lh || n_levels <= 0 || n_levels > TB_MAX_LEVELS || max_batch <= 0 || max_target <= 0) return TB_E_INVALID;
    *out = nullptr;
    for (int l = 0; l < n_levels; l++) {
        if (lw[l] <= 0 || lh[l] <= 0) return tb_fail(ctx, TB_E_INVALID, "level size must be positive");
        if (lw[l] > 4096 || lh[l] > 4096) return tb_fail(ctx, TB_E_UNSUPPORTED, "level dimensions above 4096 are not supported (12-bit packed coordinates)");
    }
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    tb_frontend *fe = new tb_frontend();
    fe->ctx = ctx; fe->L = n_levels; fe->B = max_batch; fe->max_target = max_target;
    fe->kept_cap = max_target + 8;
    fe->kp_cap = max_target + 4 * n_levels + 16;
    int rc = TB_OK;
    const size_t B = (size_t)max_batch;
    for (int l = 0; l < n_levels && !rc; l++) {
        LevelGeom &g = fe->g[l];
        level_geometry(g, lw[l], lh[l], max_batch);
        if ((rc = dev_alloc(fe, &fe->pyr[l], g.frame_stride * B))) break;
        if ((rc = dev_alloc(fe, &fe->blur[l], g.frame_stride * B))) break;
        if ((rc = dev_alloc(fe, &fe->cand[l], (size_t)g.cand_cap * B))) break;
        if ((rc = dev_alloc(fe, &fe->nid[l], (size_t)g.cand_cap * B))) break;
        if ((rc = dev_alloc(fe, &fe->kept[l], (size_t)fe->kept_cap * B))) break;
        // padding bytes of every row must be defined (4-byte loads may touch them)
        cudaMemsetAsync(fe->pyr[l], 0, g.frame_stride * B, ctx->stream);
        if (l > 0) {
            fe->area2[l] = tb_resize_is_area2(lw[l - 1], lh[l - 1], lw[l], lh[l]);
            std::vector<ResizeTab> tx, ty; std::vector<int32_t> x1, y1;
            tb_build_resize_tab(lw[l - 1], lw[l], true, tx, x1);
            tb_build_resize_tab(lh[l - 1], lh[l], false, ty, y1);
            tb_resize_box(tx, x1, ty, y1, &fe->box_pitch[l], &fe->box_rows[l]);
            std::vector<int4> hx(lw[l]), hy(lh[l]);
            for (int i = 0; i < lw[l]; i++) hx[i] = make_int4(tx[i].ofs, x1[i], tx[i].c0, tx[i].c1);
            for (int i = 0; i < lh[l]; i++) hy[i] = make_int4(ty[i].ofs, y1[i], ty[i].c0, ty[i].c1);
            if ((rc = dev_alloc(fe, &fe->xtab[l], hx.size()))) break;
            if ((rc = dev_alloc(fe, &fe->ytab[l], hy.size()))) break;
            cudaMemcpy(fe->xtab[l], hx.data(), hx.size() * sizeof(int4), cudaMemcpyHostToDevice);
            cudaMemcpy(fe->ytab[l], hy.data(), hy.size() * sizeof(int4), cudaMemcpyHostToDevice);
            fe->span8[l] = tb_resize_span8_ok(tx, x1);
            if (!fe->area2[l] && tb_resize_span_ok(tx, x1))
                fe->has_tmap[l] = tb_resize_tensor_map(&fe->pyr_tmap[l], fe->pyr[l - 1], lw[l - 1], lh[l - 1], B, fe->g[l - 1].pitch,
                                                       fe->g[l - 1].frame_stride, fe->box_rows[l]);
        }
    }
    if (!rc) rc = dev_alloc(fe, &fe->n_cand, B * n_levels);
    if (!rc) rc = dev_alloc(fe, &fe->n_kept, B * n_levels);
    if (!rc) rc = dev_alloc(fe, &fe->n_kp, B);
    if (!rc) rc = dev_alloc(fe, &fe->overflow, 4);
    if (!rc) rc = dev_alloc(fe, &fe->kps, B * fe->kp_cap);
    if (!rc) rc = dev_alloc(fe, &fe->desc, B * fe->kp_cap * 32);
    if (!rc) {
        cudaMemsetAsync(fe->n_kp, 0, sizeof(int) * B, ctx->stream);
        cudaMemsetAsync(fe->n_kept, 0, sizeof(int) * B * n_levels, ctx->stream);
        cudaMemsetAsync(fe->n_cand, 0, sizeof(int) * B * n_levels, ctx->stream);
        cudaMemsetAsync(fe->overflow, 0, sizeof(int) * 4, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = tb_fail(ctx, TB_E_CUDA, "frontend init failed");
    }
    if (rc) { tb_frontend_destroy(fe); return rc; }
    *out = fe;
    return TB_OK;
}

extern "C" int tb_frontend_kp_capacity(const tb_frontend *fe) { return fe ? fe->kp_cap : 0; }

// ORBExtractor::mnFeaturesPerLevel (ORBextractor.h:88) lives in the frontend; a caller that re-creates its frontend (other
// pyramid shape, larger target) carries it across with this pair, so that AddPoints still finds the budget of the last operator().
extern "C" int tb_frontend_get_budget(const tb_frontend *fe, int32_t *per_level, int *n_levels)
{
    if (!fe || !per_level || !n_levels) return TB_E_INVALID;
    *n_levels = fe->has_budget ? fe->L : 0;
    for (int l = 0; l < *n_levels; l++) per_level[l] = fe->budget[l];
    return TB_OK;
}
extern "C" int tb_frontend_set_budget(tb_frontend *fe, const int32_t *per_level, int n_levels)
{
    if (!fe || !per_level || n_levels != fe->L) return TB_E_INVALID;
    for (int l = 0; l < n_levels; l++) {
        if (per_level[l] < 0 || per_level[l] > fe->max_target + 8) return tb_fail(fe->ctx, TB_E_INVALID, "tb_frontend_set_budget: a level's budget exceeds the frontend's max_target");
        fe->budget[l] = per_level[l];
    }
    fe->has_budget = true;
    return TB_OK;
}

extern "C" int tb_frontend_level_layout(const tb_frontend *fe, int level, int blurred, void **base, size_t *pitch, size_t *frame_stride)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || level < 0 || level >= fe->L) return TB_E_INVALID;
    if (level >= 1) const_cast<tb_frontend *>(fe)->levels_exposed = true;
    if (base) *base = blurred ? fe->blur[level] : fe->pyr[level];
    if (pitch) *pitch = fe->g[level].pitch;
    if (frame_stride) *frame_stride = fe->g[level].frame_stride;
    return TB_OK;
}

extern "C" int tb_frontend_upload_level(tb_frontend *fe, int frame, int level, const uint8_t *host, size_t host_stride)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !host || frame < 0 || frame >= fe->B || level < 0 || level >= fe->L) return TB_E_INVALID;
    const LevelGeom &g = fe->g[level];
    TB_CUDA(fe->ctx, cudaMemcpy2DAsync(fe->pyr[level] + (size_t)frame * g.frame_stride, g.pitch, host, host_stride, g.w, g.h,
                                       cudaMemcpyHostToDevice, fe->ctx->stream));
    return TB_OK;
}

extern "C" int tb_frontend_upload_frames(tb_frontend *fe, int first, int n, const uint8_t *host, size_t host_stride)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !host || first < 0 || n < 0 || first + n > fe->B) return TB_E_INVALID;
    const LevelGeom &g = fe->g[0];
    // frames are back to back on the host and at frame_stride on the device: one strided copy
    TB_CUDA(fe->ctx, cudaMemcpy2DAsync(fe->pyr[0] + (size_t)first * g.frame_stride, g.pitch, host, host_stride, g.w, (size_t)g.h * n,
                                       cudaMemcpyHostToDevice, fe->ctx->stream));
    return TB_OK;
}

// PinholeCamera::UndistortImage in front of the pyramid (test_vo.cpp:178): raw frames -> level 0 through the remap
extern "C" int tb_frontend_undistort_device(tb_frontend *fe, const tb_undistort *und, int first, int n, const void *raw, int raw_w, int raw_h,
                                            size_t raw_pitch, size_t raw_frame_stride)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !und || !raw || first < 0 || n < 0 || first + n > fe->B) return TB_E_INVALID;
    const LevelGeom &g = fe->g[0];
    // the remap writes und->h rows of und->w pixels per frame: a map built for another size would run into the next frame
    if (tb_undistort_width(und) != g.w || tb_undistort_height(und) != g.h)
        return tb_fail(fe->ctx, TB_E_INVALID, "tb_frontend_undistort_device: the undistort map's size differs from pyramid level 0");
    if (raw_w <= 0 || raw_h <= 0 || raw_pitch < (size_t)raw_w || (n > 1 && raw_frame_stride < raw_pitch * (size_t)raw_h))
        return tb_fail(fe->ctx, TB_E_INVALID, "tb_frontend_undistort_device: raw frame layout (pitch < width or frame stride < pitch*height)");
    return tb_undistort_image_device(fe->ctx, und, raw, raw_w, raw_h, raw_pitch, raw_frame_stride, fe->pyr[0] + (size_t)first * g.frame_stride,
                                     g.pitch, g.frame_stride, n);
}

extern "C" int tb_frontend_upload_frames_undistort(tb_frontend *fe, const tb_undistort *und, int first, int n, const uint8_t *host,
                                                   int raw_w, int raw_h, size_t host_stride)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !und || !host || first < 0 || n < 0 || first + n > fe->B || raw_w <= 0 || raw_h <= 0 || host_stride < (size_t)raw_w) return TB_E_INVALID;
    if (n == 0) return TB_OK;
    const size_t pitch = align_up((size_t)raw_w, 64), need = pitch * raw_h * n;
    if (need > fe->raw_bytes) {
        TB_CUDA(fe->ctx, cudaStreamSynchronize(fe->ctx->stream));
        cudaFree(fe->raw); fe->raw = nullptr; fe->raw_bytes = 0;
        TB_CUDA(fe->ctx, cudaMalloc(&fe->raw, need));
        fe->raw_bytes = need;
    }
    TB_CUDA(fe->ctx, cudaMemcpy2DAsync(fe->raw, pitch, host, host_stride, raw_w, (size_t)raw_h * n, cudaMemcpyHostToDevice, fe->ctx->stream));
    return tb_frontend_undistort_device(fe, und, first, n, fe->raw, raw_w, raw_h, pitch, pitch * raw_h);
}

extern "C" int tb_frontend_download_level(tb_frontend *fe, int frame, int level, int blurred, uint8_t *host, size_t host_stride)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !host || frame < 0 || frame >= fe->B || level < 0 || level >= fe->L) return TB_E_INVALID;
    const LevelGeom &g = fe->g[level];
    if (blurred && !fe->blur_valid) {
        // blur on demand: ORBExtractor's GaussianBlur(7x7, sigma 2) of this frame's level (ORBextractor.cpp:959-960)
        int rc = tb_launch_gauss(fe->ctx, fe->ctx->stream, fe->pyr[level] + (size_t)frame * g.frame_stride, fe->blur[level] + (size_t)frame * g.frame_stride,
                                 g.w, g.h, g.pitch, g.frame_stride, 1);
        if (rc) return rc;
    }
    const uint8_t *src = (blurred ? fe->blur[level] : fe->pyr[level]) + (size_t)frame * g.frame_stride;
    TB_CUDA(fe->ctx, cudaMemcpy2DAsync(host, host_stride, src, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost, fe->ctx->stream));
    TB_CUDA(fe->ctx, cudaStreamSynchronize(fe->ctx->stream));
    return TB_OK;
}

// batches up to this size replay captured graphs (TB_GRAPHS=0 disables; profiling always launches directly)
static int graph_max_frames()
{
    static const int v = getenv("TB_GRAPHS") ? atoi(getenv("TB_GRAPHS")) : 16;
    return v;
}
// Graph replay pays where a call is launch bound: small batches of small frames (a KITTI frame is ~50 launches for well under
// a millisecond of kernels).  Above ~9 Mpx of level-0 pixels per call (one 4K frame, four 1080p frames) the kernels outlast the
// launches and the call is enqueued directly.
static bool graph_worthwhile(const tb_frontend *fe, int n_frames)
{
    return n_frames <= graph_max_frames() && (size_t)n_frames * fe->g[0].w * fe->g[0].h <= (size_t)9 << 20;
}
// capture `enqueue` on the ctx stream into slot (or replay it if the slot already holds this parameter set)
template <typename F>
static int run_captured(tb_ctx *ctx, tb_frontend::GraphSlot &slot, bool same, F enqueue)
{
    if (!same || !slot.exec) {
        if (slot.exec) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
        const int64_t l0 = ctx->launches;
        TB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess || !graph) return tb_fail(ctx, TB_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        const cudaError_t ei = cudaGraphInstantiate(&slot.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { slot.exec = nullptr; return tb_fail(ctx, TB_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei)); }
        slot.launches = ctx->launches - l0;
        ctx->launches = l0;
    }
    TB_CUDA(ctx, cudaGraphLaunch(slot.exec, ctx->stream));
    ctx->launches += slot.launches;
    return TB_OK;
}

static int pyramid_enqueue(tb_frontend *fe, int n_frames)
{
    for (int l = 1; l < fe->L; l++) {
        const LevelGeom &s = fe->g[l - 1], &d = fe->g[l];
        int rc = tb_launch_resize(fe->ctx, fe->pyr[l - 1], s.pitch, s.frame_stride, fe->pyr[l], d.pitch, d.frame_stride,
                                  d.w, d.h, fe->xtab[l], fe->ytab[l], fe->area2[l], n_frames, fe->box_pitch[l], fe->box_rows[l],
                                  fe->has_tmap[l] ? &fe->pyr_tmap[l] : nullptr, fe->span8[l]);
        if (rc) return rc;
    }
    return TB_OK;
}

// Large batches: the level chain runs on the side stream and records an event per level; tb_frontend_extract lets FAST(l)
// wait for level l only, so FAST of the fine levels overlaps the resizes of the coarse ones (both kernels are bound by
// instruction issue on different pipes and by their launch tails: 6.16 -> 5.96 ms per 256-frame 1080p step). Every other
// entry point waits for the whole pyramid first (TB_DEVICE_GUARD). TB_PYR_ASYNC=0 builds on the main stream.
static int pyramid_enqueue_async(tb_frontend *fe, int n_frames)
{
    tb_ctx *ctx = fe->ctx;
    TB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    TB_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->side;
    int rc = TB_OK;
    for (int l = 1; l < fe->L && rc == TB_OK; l++) {
        const LevelGeom &s = fe->g[l - 1], &d = fe->g[l];
        rc = tb_launch_resize(ctx, fe->pyr[l - 1], s.pitch, s.frame_stride, fe->pyr[l], d.pitch, d.frame_stride,
                              d.w, d.h, fe->xtab[l], fe->ytab[l], fe->area2[l], n_frames, fe->box_pitch[l], fe->box_rows[l],
                              fe->has_tmap[l] ? &fe->pyr_tmap[l] : nullptr, fe->span8[l]);
        if (cudaEventRecord(ctx->ev_pyr[l], ctx->side) != cudaSuccess && rc == TB_OK) rc = TB_E_CUDA;
    }
    ctx->stream = main_stream;
    if (fe->L > 1) { ctx->pyr_pending = true; ctx->pyr_last = fe->L - 1; }   // (also after a failed launch: what was enqueued is waited for)
    return rc;
}

extern "C" int tb_frontend_build_pyramid(tb_frontend *fe, int n_frames)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || n_frames < 0 || n_frames > fe->B) return TB_E_INVALID;
    if (n_frames == 0) return TB_OK;
    tb_ctx *ctx = fe->ctx;
    static const bool pyr_async = !(getenv("TB_PYR_ASYNC") && atoi(getenv("TB_PYR_ASYNC")) == 0);   // A/B knob: 0 = build on the main stream
    if (pyr_async && !ctx->profiling && n_frames >= 32 && !fe->levels_exposed) return pyramid_enqueue_async(fe, n_frames);
    if (ctx->profiling || !graph_worthwhile(fe, n_frames)) return pyramid_enqueue(fe, n_frames);
    const bool same = fe->g_pyramid.n_frames == n_frames;
    fe->g_pyramid.n_frames = n_frames;
    return run_captured(ctx, fe->g_pyramid, same, [&]() { return pyramid_enqueue(fe, n_frames); });
}

// ORBextractor.cpp:919-930
static void orb_budget(int L, float factor, int target, int *per_level)
{
    volatile float denom = 1 - (float)pow((double)factor, (double)L);
    volatile float numer = target * (1 - factor);
    volatile float nDesired = numer / denom;
    int sum = 0;
    for (int level = 0; level < L - 1; level++) {
        per_level[level] = (int)lrintf(nDesired);
        sum += per_level[level];
        nDesired = nDesired * factor;
    }
    per_level[L - 1] = target - sum > 0 ? target - sum : 0;
}

static int extract_enqueue(tb_frontend *fe, int n_frames, const float *sf, float initTh, float minTh, bool add_points, int n_exit);

extern "C" int tb_frontend_extract(tb_frontend *fe, int n_frames, const float *sf, int target, float initTh, float minTh,
                                   const tb_keypoint *exit_kps, int n_exit)
{
    TB_DEVICE_GUARD_NOJOIN(fe ? fe->ctx : nullptr);   // extract_enqueue waits for the pyramid level by level
    if (!fe || !sf || n_frames < 0 || n_frames > fe->B || target <= 0 || n_exit < 0) return TB_E_INVALID;
    tb_ctx *ctx = fe->ctx;
    if (target > fe->max_target) return tb_fail(ctx, TB_E_INVALID, "target exceeds the frontend's max_target");
    if (n_frames == 0) return TB_OK;
    const int L = fe->L;
    const bool add_points = exit_kps != nullptr;
    if (add_points) {
        if (!fe->has_budget) return tb_fail(ctx, TB_E_INVALID, "AddPoints before operator(): the reference reads an unset mnFeaturesPerLevel here");
    } else {
        if (L < 2) { fe->budget[0] = target; } // (the reference indexes mvScaleFactor[1]; single level => all features)
        else orb_budget(L, sf[1], target, fe->budget);
        fe->has_budget = true;
    }
    if (add_points && n_exit > 0) {
        if (n_exit > fe->exit_cap) {
            int rc = dev_alloc(fe, &fe->exit_xy, (size_t)n_exit);
            if (!rc) rc = dev_alloc(fe, &fe->enid, (size_t)n_exit * fe->B * fe->L);   // per (level, frame): all levels run in one launch
            if (rc) return rc;
            fe->exit_cap = n_exit;
        }
        std::vector<float2> xy(n_exit);
        for (int i = 0; i < n_exit; i++) xy[i] = make_float2(exit_kps[i].x, exit_kps[i].y);
        TB_CUDA(ctx, cudaMemcpyAsync(fe->exit_xy, xy.data(), sizeof(float2) * n_exit, cudaMemcpyHostToDevice, ctx->stream));
        TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (add_points || ctx->profiling || !graph_worthwhile(fe, n_frames))
        return extract_enqueue(fe, n_frames, sf, initTh, minTh, add_points, n_exit);
    TB_DEVICE_GUARD(ctx);   // a replayed graph cannot wait level by level: join a pyramid that is still being built
    tb_frontend::GraphSlot &gs = fe->g_extract;
    bool same = gs.n_frames == n_frames && gs.target == target && gs.initTh == initTh && gs.minTh == minTh;
    for (int l = 0; l < L && same; l++) same = gs.sf[l] == sf[l];
    gs.n_frames = n_frames; gs.target = target; gs.initTh = initTh; gs.minTh = minTh;
    for (int l = 0; l < L; l++) gs.sf[l] = sf[l];
    return run_captured(ctx, gs, same, [&]() { return extract_enqueue(fe, n_frames, sf, initTh, minTh, false, 0); });
}

// the launch sequence of one extract (captured into a graph for small batches)
static int extract_enqueue(tb_frontend *fe, int n_frames, const float *sf, float initTh, float minTh, bool add_points, int n_exit)
{
    tb_ctx *ctx = fe->ctx;
    const int L = fe->L;
    TB_CUDA(ctx, cudaMemsetAsync(fe->n_cand, 0, sizeof(int) * (size_t)n_frames * L, ctx->stream));
    TB_CUDA(ctx, cudaMemsetAsync(fe->n_kept, 0, sizeof(int) * (size_t)n_frames * L, ctx->stream));
    TB_CUDA(ctx, cudaMemsetAsync(fe->overflow, 0, sizeof(int) * 4, ctx->stream));

    // Three streams inside one extract (joined before the describe launch):
    //   main : FAST level 0..L-1 (instruction bound, the critical path)
    //   side : the Gaussian levels — they only need the pyramid, so they start right away and fill FAST's issue gaps
    //   side2: the octree of level l as soon as FAST(l) is done — latency bound (one CTA per frame), it hides entirely
    //          behind FAST of the following levels
    // (With per-kernel profiling on, everything stays on the main stream so the event pairs stay meaningful.)
    const bool overlap = !ctx->profiling;
    const bool dense_blur = !tb_describe_fused();   // the fused describe kernel blurs on demand; the dense levels are made lazily
    fe->blur_valid = dense_blur;
    // per-level octree launches only pay when FAST of the next level is long enough to hide them; for small batches (the
    // single-frame host entries) one launch with all levels side by side has the lower latency
    const bool per_level = overlap && n_frames >= 32;
    cudaStream_t gs = overlap ? ctx->side : ctx->stream, os = per_level ? ctx->side2 : ctx->stream;
    if (overlap) {
        TB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        TB_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        for (int l = 0; l < L && dense_blur; l++) {
            const LevelGeom &g = fe->g[l];
            int rc = tb_launch_gauss(ctx, gs, fe->pyr[l], fe->blur[l], g.w, g.h, g.pitch, g.frame_stride, n_frames);
            if (rc) return rc;
        }
        TB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
    }
    DescArgs d;
    memset(&d, 0, sizeof d);
    d.n_levels = L; d.n_kept = fe->n_kept; d.kps = fe->kps; d.desc = fe->desc; d.kp_cap = fe->kp_cap; d.n_kp = fe->n_kp;
    d.overflow = fe->overflow;
    for (int l = 0; l < L; l++) {
        const LevelGeom &g = fe->g[l];
        d.lv[l].raw = fe->pyr[l]; d.lv[l].blur = fe->blur[l]; d.lv[l].pitch = g.pitch; d.lv[l].frame_stride = g.frame_stride;
        d.lv[l].w = g.w; d.lv[l].h = g.h;
        d.lv[l].kept = fe->kept[l]; d.lv[l].kept_cap = fe->kept_cap;
        d.lv[l].sf = sf[l];
        volatile float ps = 31 * sf[l];
        d.lv[l].size = (float)(int)ps; // (int)(PATCH_SIZE*mvScaleFactor[level]) :812
    }
    // with blur on demand the describe kernel only needs the pyramid and the octree's winners: for large batches it follows the
    // octree of its level on the side stream and hides behind FAST of the next levels, as the octree does
    const bool desc_per_level = per_level && tb_describe_fused();
    int max_budget = 0;
    OctArgs oct[TB_MAX_LEVELS];
    for (int l = 0; l < L; l++) {
        const LevelGeom &g = fe->g[l];
        if (g.nColsRun > 0 && g.nRowsRun > 0) {
            FastArgs a;
            a.img = fe->pyr[l]; a.pitch = g.pitch; a.frame_stride = g.frame_stride; a.w = g.w; a.h = g.h;
            a.nCols = g.nCols; a.nRows = g.nRows; a.wCell = g.wCell; a.hCell = g.hCell; a.maxBX = g.maxBX; a.maxBY = g.maxBY;
            static const int gpx = getenv("TB_FAST_GPX") ? atoi(getenv("TB_FAST_GPX")) : 256;   // tile width in pixels (tuning knob)
            a.G = std::min(gpx, TB_FAST_TPITCH - 16) / g.wCell; if (a.G < 1) a.G = 1; if (a.G > g.nColsRun) a.G = g.nColsRun; if (a.G > 32) a.G = 32;
            a.scw = a.G * g.wCell;
            if (a.scw > TB_FAST_TPITCH - 16 || (g.hCell + 6) * TB_FAST_TPITCH > 65535) return tb_fail(ctx, TB_E_UNSUPPORTED, "FAST cell larger than the tile");
            a.tpitch = TB_FAST_TPITCH;   // 8-byte aligned origin + halo + spare for the E-side read; multiple of 16 keeps every shared array 16-byte aligned
            {
                // full-width tile: x0 = 16 + k*G*wCell; when that is a multiple of 8 for every k the tile geometry is fixed
                auto magic = [](int n) -> uint32_t { return n > 1 ? 0xFFFFFFFFu / (uint32_t)n + 1u : 0u; };
                const int ox = 16 & 7, cs = ox + 3, ce = cs + a.scw;     // valid when (G*wCell) % 8 == 0; otherwise the kernel recomputes
                const bool fixed = ((a.G * g.wCell) & 7) == 0;
                a.magic_groups = fixed ? magic(((ce + 15) >> 4) - (cs >> 4)) : 0xFFFFFFFFu;   // 16-pixel prefilter groups per tile row
                a.magic_w8 = fixed ? magic((ox + a.scw + 6 + 7) >> 3) : 0xFFFFFFFFu;
                a.magic_wcell = magic(g.wCell);
            }
            a.initTh = (int)initTh; a.minTh = (int)minTh;
            a.one = 1;
            a.cand = fe->cand[l]; a.cand_cap = g.cand_cap;
            a.n_cand = fe->n_cand + l; a.n_cand_stride = L;
            a.overflow = fe->overflow;
            {
                static const int forced = getenv("TB_FAST_LISTCAP") ? atoi(getenv("TB_FAST_LISTCAP")) : 0;
                // per-warp corner lists: as long as TB_FAST_CTAS CTAs still fit an SM (227 KB / CTAs - 1 KB reserved - static), never more
                // than a warp's share of the tile could fill (TB_FAST_LISTCAP overrides: tests force the overflow path with a tiny list)
                const int nw = TB_FAST_NT / 32;
                const int budget = 232448 / TB_FAST_CTAS - 1024 - 128;
                const int base = (int)tb_fast_smem_bytes(g.hCell, a.scw, a.tpitch, 0);
                static const int list_div = getenv("TB_FAST_LIST_DIV") ? atoi(getenv("TB_FAST_LIST_DIV")) : 0;   // developer knob: list = tile pixels / this
                int cap = list_div > 0 ? (g.hCell * a.scw) / list_div : std::min(g.hCell * a.scw + nw * 32, ((budget - base) / (2 * nw)) * nw);
                a.list_cap = forced > 0 ? forced : std::max(256, cap);
            }
            static const int smem_pad = getenv("TB_FAST_SMEM_PAD") ? atoi(getenv("TB_FAST_SMEM_PAD")) : 0;   // developer knob: fewer resident CTAs for A/B runs
            size_t smem = tb_fast_smem_bytes(g.hCell, a.scw, a.tpitch, a.list_cap) + smem_pad;
            if (ctx->pyr_pending && l >= 1 && l <= ctx->pyr_last) TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pyr[l], 0));
            int rc = tb_launch_fast(ctx, a, g.nColsRun, g.nRowsRun, n_frames, smem);
            if (rc) return rc;
        }
        OctArgs &o = oct[l];
        o.cand = fe->cand[l]; o.cand_cap = g.cand_cap; o.n_cand = fe->n_cand + l; o.n_stride = L;
        o.nid = fe->nid[l]; o.N = fe->budget[l];
        o.W = g.maxBX - 16; o.H = g.maxBY - 16; o.nIni = g.nIni; o.hX = g.hX; o.wCell = g.wCell; o.hCell = g.hCell;
        o.exit_xy = fe->exit_xy; o.n_exit = add_points ? n_exit : 0; o.enid = fe->enid ? fe->enid + (size_t)l * fe->exit_cap * fe->B : nullptr;
        o.kept = fe->kept[l]; o.kept_cap = fe->kept_cap; o.n_kept = fe->n_kept + l; o.overflow = fe->overflow;
        o.node_cap = fe->budget[l] + 4 * (g.nIni > 0 ? g.nIni : 1) + 8;
        if (fe->budget[l] > max_budget) max_budget = fe->budget[l];
        if (per_level) {
            TB_CUDA(ctx, cudaEventRecord(ctx->ev_lvl[l], ctx->stream));
            TB_CUDA(ctx, cudaStreamWaitEvent(os, ctx->ev_lvl[l], 0));
            int rc = tb_launch_octree(ctx, &oct[l], 1, n_frames, os);
            if (rc) return rc;
            if (desc_per_level) { rc = tb_launch_describe(ctx, d, fe->budget[l] + 4, n_frames, l, 1, os); if (rc) return rc; }
        }
    }
    if (ctx->pyr_pending) {   // every level has been waited for (levels without a FAST launch: wait for the last one)
        TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pyr[ctx->pyr_last], 0));
        ctx->pyr_pending = false;
    }
    if (!per_level) {
        int rc = tb_launch_octree(ctx, oct, L, n_frames, ctx->stream);
        if (rc) return rc;
    }
    if (overlap) {
        if (per_level) {
            TB_CUDA(ctx, cudaEventRecord(ctx->ev_join2, os));
            TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));
        }
        TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    } else {
        int rc = TB_OK;
        for (int l = 0; l < L && dense_blur; l++) {
            const LevelGeom &g = fe->g[l];
            rc = tb_launch_gauss(ctx, gs, fe->pyr[l], fe->blur[l], g.w, g.h, g.pitch, g.frame_stride, n_frames);
            if (rc) return rc;
        }
    }
    if (desc_per_level) return TB_OK;   // every level's describe launch already follows its octree on the side stream
    return tb_launch_describe(ctx, d, max_budget + 4, n_frames, 0, L, ctx->stream);
}

static int check_overflow(tb_frontend *fe)
{
    int ov[4] = {0, 0, 0, 0};
    TB_CUDA(fe->ctx, cudaMemcpyAsync(ov, fe->overflow, sizeof ov, cudaMemcpyDeviceToHost, fe->ctx->stream));
    TB_CUDA(fe->ctx, cudaStreamSynchronize(fe->ctx->stream));
    if (ov[0]) return tb_fail(fe->ctx, TB_E_CAPACITY, "internal buffer overflow during extract (candidates / octree nodes / keypoints)");
    return TB_OK;
}

extern "C" int tb_frontend_results_device(const tb_frontend *fe, void **kps, void **desc, void **counts)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe) return TB_E_INVALID;
    if (kps) *kps = fe->kps;
    if (desc) *desc = fe->desc;
    if (counts) *counts = fe->n_kp;
    return TB_OK;
}

extern "C" int tb_frontend_get_results(tb_frontend *fe, int frame, tb_keypoint *kps, uint8_t *desc, int capacity, int *n)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || frame < 0 || frame >= fe->B || !n || capacity < 0) return TB_E_INVALID;
    tb_ctx *ctx = fe->ctx;
    int rc = check_overflow(fe);
    if (rc) return rc;
    int cnt = 0;
    TB_CUDA(ctx, cudaMemcpy(&cnt, fe->n_kp + frame, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    int m = cnt < capacity ? cnt : capacity;
    if (m > 0 && kps) TB_CUDA(ctx, cudaMemcpy(kps, fe->kps + (size_t)frame * fe->kp_cap, sizeof(tb_keypoint) * m, cudaMemcpyDeviceToHost));
    if (m > 0 && desc) TB_CUDA(ctx, cudaMemcpy(desc, fe->desc + (size_t)frame * fe->kp_cap * 32, (size_t)32 * m, cudaMemcpyDeviceToHost));
    return cnt > capacity ? tb_fail(ctx, TB_E_CAPACITY, "keypoint output capacity too small") : TB_OK;
}

extern "C" int tb_frontend_get_results_batch(tb_frontend *fe, int n_frames, tb_keypoint *kps, uint8_t *desc, int capacity, int32_t *counts)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || n_frames < 0 || n_frames > fe->B || !counts || capacity < 0) return TB_E_INVALID;
    tb_ctx *ctx = fe->ctx;
    if (n_frames == 0) return TB_OK;
    const int m = capacity < fe->kp_cap ? capacity : fe->kp_cap;
    TB_CUDA(ctx, cudaMemcpyAsync(counts, fe->n_kp, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    if (kps) TB_CUDA(ctx, cudaMemcpy2DAsync(kps, sizeof(tb_keypoint) * capacity, fe->kps, sizeof(tb_keypoint) * fe->kp_cap,
                                            sizeof(tb_keypoint) * m, n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    if (desc) TB_CUDA(ctx, cudaMemcpy2DAsync(desc, (size_t)32 * capacity, fe->desc, (size_t)32 * fe->kp_cap, (size_t)32 * m, n_frames,
                                             cudaMemcpyDeviceToHost, ctx->stream));
    int rc = check_overflow(fe);
    if (rc) return rc;
    for (int f = 0; f < n_frames; f++) if (counts[f] > capacity) return tb_fail(ctx, TB_E_CAPACITY, "keypoint output capacity too small");
    return TB_OK;
}

extern "C" int tb_frontend_get_results_batch_async(tb_frontend *fe, int n_frames, tb_keypoint *kps, uint8_t *desc, int capacity, int32_t *counts)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || n_frames < 0 || n_frames > fe->B || !counts || capacity < 0) return TB_E_INVALID;
    tb_ctx *ctx = fe->ctx;
    if (n_frames == 0) return TB_OK;
    const int m = capacity < fe->kp_cap ? capacity : fe->kp_cap;
    TB_CUDA(ctx, cudaMemcpyAsync(counts, fe->n_kp, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    if (kps) TB_CUDA(ctx, cudaMemcpy2DAsync(kps, sizeof(tb_keypoint) * capacity, fe->kps, sizeof(tb_keypoint) * fe->kp_cap,
                                            sizeof(tb_keypoint) * m, n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    if (desc) TB_CUDA(ctx, cudaMemcpy2DAsync(desc, (size_t)32 * capacity, fe->desc, (size_t)32 * fe->kp_cap, (size_t)32 * m, n_frames,
                                             cudaMemcpyDeviceToHost, ctx->stream));
    return TB_OK;
}

extern "C" int tb_frontend_status(tb_frontend *fe)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe) return TB_E_INVALID;
    return check_overflow(fe);
}

extern "C" int tb_frontend_get_level_counts(tb_frontend *fe, int frame, int32_t *n_per_level)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || frame < 0 || frame >= fe->B || !n_per_level) return TB_E_INVALID;
    TB_CUDA(fe->ctx, cudaStreamSynchronize(fe->ctx->stream));
    TB_CUDA(fe->ctx, cudaMemcpy(n_per_level, fe->n_kept + (size_t)frame * fe->L, sizeof(int) * fe->L, cudaMemcpyDeviceToHost));
    return TB_OK;
}

// Candidates are stored unordered; the reference order is (cell row, cell col, y, x). Sort on the
// host for the stage test only (this accessor is test plumbing, not part of the data path).
extern "C" int tb_frontend_get_candidates(tb_frontend *fe, int frame, int level, tb_keypoint *out, int capacity, int *n)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || frame < 0 || frame >= fe->B || level < 0 || level >= fe->L || !n) return TB_E_INVALID;
    tb_ctx *ctx = fe->ctx;
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int cnt = 0;
    TB_CUDA(ctx, cudaMemcpy(&cnt, fe->n_cand + (size_t)frame * fe->L + level, sizeof(int), cudaMemcpyDeviceToHost));
    const LevelGeom &g = fe->g[level];
    if (cnt > g.cand_cap) cnt = g.cand_cap;
    std::vector<uint32_t> c(cnt);
    if (cnt) TB_CUDA(ctx, cudaMemcpy(c.data(), fe->cand[level] + (size_t)frame * g.cand_cap, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost));
    std::vector<uint64_t> key(cnt);
    for (int i = 0; i < cnt; i++) {
        uint64_t x = cand_x(c[i]), y = cand_y(c[i]);
        key[i] = (((y - 3) / g.hCell) << 52) | (((x - 3) / g.wCell) << 40) | (y << 28) | (x << 16) | (uint64_t)cand_m(c[i]);
    }
    std::sort(key.begin(), key.end());
    *n = cnt;
    for (int i = 0; i < cnt && i < capacity; i++) {
        tb_keypoint k = {(float)((key[i] >> 16) & 0xFFF), (float)((key[i] >> 28) & 0xFFF), 7.f, -1.f, (float)((int)(key[i] & 0xFF) - 1), 0, -1};
        out[i] = k;
    }
    return cnt > capacity ? TB_E_CAPACITY : TB_OK;
}

extern "C" int tb_orb_extract_host(tb_frontend *fe, const uint8_t *const *levels, const size_t *host_strides, const float *sf,
                                   int target, float initTh, float minTh, const tb_keypoint *exit_kps, int n_exit,
                                   tb_keypoint *kps, uint8_t *desc, int capacity, int *n)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !levels || !host_strides || !n) return TB_E_INVALID;
    for (int l = 0; l < fe->L; l++) {
        int rc = tb_frontend_upload_level(fe, 0, l, levels[l], host_strides[l]);
        if (rc) return rc;
    }
    int rc = tb_frontend_extract(fe, 1, sf, target, initTh, minTh, exit_kps, n_exit);
    if (rc) return rc;
    return tb_frontend_get_results(fe, 0, kps, desc, capacity, n);
}

extern "C" int tb_frontend_search_by_bf(tb_frontend *fe, int n_pairs, const int32_t *q_frames, const int32_t *t_frames,
                                        float ratio, float minTh, void *matches, int capacity, void *counts)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || n_pairs < 0 || !q_frames || !t_frames || !matches || !counts) return TB_E_INVALID;
    for (int p = 0; p < n_pairs; p++)
        if (q_frames[p] < 0 || q_frames[p] >= fe->B || t_frames[p] < 0 || t_frames[p] >= fe->B) return TB_E_INVALID;
    return tb_match_bf_batched(fe->ctx, n_pairs, q_frames, t_frames, fe->B, fe->desc, (size_t)fe->kp_cap * 32, fe->n_kp, fe->kp_cap,
                               ratio, minTh, (tb_match *)matches, capacity, (int *)counts);
}

extern "C" int tb_frontend_set_bow(tb_frontend *fe, const tb_vocab *voc, int n_frames, int levelsup, void *node, void *word, void *weight)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !voc || n_frames < 0 || n_frames > fe->B || !node) return TB_E_INVALID;
    return tb_window_set_bow_batched(fe->ctx, voc, n_frames, levelsup, fe->desc, fe->kp_cap, fe->n_kp, (int32_t *)node, (int32_t *)word, (double *)weight);
}

extern "C" int tb_frontend_search_by_bow(tb_frontend *fe, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames, const int32_t *f2_frames,
                                         const void *node, const void *mp_valid, int map_point_only, void *matches, int capacity, void *counts)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !p || n_pairs < 0 || !f1_frames || !f2_frames || !node || !matches || !counts) return TB_E_INVALID;
    for (int i = 0; i < n_pairs; i++)
        if (f1_frames[i] < 0 || f1_frames[i] >= fe->B || f2_frames[i] < 0 || f2_frames[i] >= fe->B) return TB_E_INVALID;
    return tb_window_bow_batched(fe->ctx, p, n_pairs, f1_frames, f2_frames, fe->B, fe->kps, fe->desc, fe->kp_cap, fe->n_kp, (const int32_t *)node,
                                 (const uint8_t *)mp_valid, map_point_only, (tb_match *)matches, capacity, (int *)counts);
}

extern "C" int tb_frontend_search_by_violence(tb_frontend *fe, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames,
                                              const int32_t *f2_frames, int min_level, int max_level, float search_r,
                                              void *matches, int capacity, void *counts)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !p || n_pairs < 0 || !f1_frames || !f2_frames || !matches || !counts) return TB_E_INVALID;
    for (int i = 0; i < n_pairs; i++)
        if (f1_frames[i] < 0 || f1_frames[i] >= fe->B || f2_frames[i] < 0 || f2_frames[i] >= fe->B) return TB_E_INVALID;
    return tb_window_violence_batched(fe->ctx, p, n_pairs, f1_frames, f2_frames, fe->B, fe->kps, fe->desc, fe->kp_cap, fe->n_kp,
                                      fe->g[0].w, fe->g[0].h, min_level, max_level, search_r, (tb_match *)matches, capacity, (int *)counts);
}

extern "C" int tb_frontend_undistort_points(tb_frontend *fe, const tb_pinhole *cam, int n_frames, void *out)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !cam || n_frames < 0 || n_frames > fe->B || !out) return TB_E_INVALID;
    return tb_undistort_points_launch(fe->ctx, cam, (const uint8_t *)fe->kps, sizeof(tb_keypoint), sizeof(tb_keypoint) * (size_t)fe->kp_cap, fe->n_kp,
                                      fe->kp_cap, n_frames, (float2 *)out);
}

extern "C" int tb_frontend_search_by_opflow(tb_frontend *fe, const tb_lk_params *p, int n_frames, int n_pairs, const int32_t *f1_frames,
                                            const int32_t *f2_frames, int equalized, void *next_pts, void *status, void *err)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !p || n_frames < 1 || n_frames > fe->B || n_pairs < 0 || !f1_frames || !f2_frames || !next_pts || !status) return TB_E_INVALID;
    for (int i = 0; i < n_pairs; i++)
        if (f1_frames[i] < 0 || f1_frames[i] >= n_frames || f2_frames[i] < 0 || f2_frames[i] >= n_frames) return TB_E_INVALID;
    const LevelGeom &g = fe->g[0];
    const uint8_t *next = fe->pyr[0];
    if (equalized) {   // img1 = F1->Equalize(): CLAHE(3.0, 8x8) of level 0 (matcher.cpp:734-737, Frame.cpp:453-458)
        uint8_t *eq = nullptr;
        int rc = tb_flow_eq_reserve(fe->ctx, g.frame_stride * (size_t)n_frames, &eq);
        if (rc) return rc;
        rc = tb_clahe_device(fe->ctx, fe->pyr[0], g.w, g.h, (size_t)g.pitch, g.frame_stride, eq, (size_t)g.pitch, g.frame_stride, n_frames, 3.0, 8, 8);
        if (rc) return rc;
        next = eq;
    }
    // calcOpticalFlowPyrLK(img2, img1, F2 keys, ...): previous image = F2 (with its keypoints), next image = F1 (matcher.cpp:745)
    return tb_flow_batched(fe->ctx, p, fe->pyr[0], (size_t)g.pitch, g.frame_stride, n_frames, next, (size_t)g.pitch, g.frame_stride, n_frames, g.w, g.h,
                           n_pairs, f2_frames, f1_frames, (const uint8_t *)fe->kps, sizeof(tb_keypoint), sizeof(tb_keypoint) * (size_t)fe->kp_cap, f2_frames,
                           fe->n_kp, fe->kp_cap, 1, (float2 *)next_pts, (uint8_t *)status, (float *)err);
}

extern "C" int tb_frontend_stereo_depth_by_opflow(tb_frontend *fe, const tb_lk_params *p, float bf, int n_frames, int n_pairs,
                                                  const int32_t *current_frames, const int32_t *stereo_frames, int equalized,
                                                  void *depth, void *next_pts, void *status)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!depth || !next_pts || !status) return TB_E_INVALID;
    // searchByOPFlow(stereo_frame, current_frame, pts, equalized, ...): F1 = stereo frame, F2 = current frame with its keys (LocalBA.cpp:54)
    int rc = tb_frontend_search_by_opflow(fe, p, n_frames, n_pairs, stereo_frames, current_frames, equalized, next_pts, status, nullptr);
    if (rc) return rc;
    return tb_flow_stereo_depth(fe->ctx, n_pairs, sizeof(tb_keypoint), fe->kp_cap, bf, (const float2 *)next_pts, (const uint8_t *)status, (float *)depth);
}

extern "C" int tb_frontend_stereo_match(tb_frontend *fe, const tb_stereo_params *p, int n_pairs, const int32_t *left_frames,
                                        const int32_t *right_frames, const float *inv_sf, void *matches, int capacity,
                                        void *counts, void *depth)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !p || n_pairs < 0 || !left_frames || !right_frames || !inv_sf || !matches || !counts || capacity < 0) return TB_E_INVALID;
    for (int i = 0; i < n_pairs; i++)
        if (left_frames[i] < 0 || left_frames[i] >= fe->B || right_frames[i] < 0 || right_frames[i] >= fe->B) return TB_E_INVALID;
    return tb_window_stereo_batched(fe->ctx, p, n_pairs, left_frames, right_frames, fe->kps, fe->desc, fe->kp_cap, fe->n_kp, inv_sf, fe->L,
                                    fe->g[0].h, (tb_match *)matches, capacity, (int *)counts, (float *)depth);
}

extern "C" int tb_frontend_search_by_projection(tb_frontend *fe, const tb_matcher_params *p, const tb_pinhole *cam, int n_pairs,
                                                const int32_t *f1_frames, const int32_t *f2_frames, const float *Tcw1,
                                                const void *f1_matched, const void *mp_valid, const void *mp_xyz,
                                                const float *sf1, void *matches, int capacity, void *counts)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !p || !cam || n_pairs < 0 || !f1_frames || !f2_frames || !Tcw1 || !mp_valid || !mp_xyz || !sf1 || !matches || !counts) return TB_E_INVALID;
    for (int i = 0; i < n_pairs; i++)
        if (f1_frames[i] < 0 || f1_frames[i] >= fe->B || f2_frames[i] < 0 || f2_frames[i] >= fe->B) return TB_E_INVALID;
    return tb_window_projection_batched(fe->ctx, p, cam, n_pairs, f1_frames, f2_frames, Tcw1, fe->B, fe->kps, fe->desc, fe->kp_cap, fe->n_kp,
                                        (const uint8_t *)f1_matched, (const uint8_t *)mp_valid, (const float *)mp_xyz, sf1, fe->L,
                                        fe->g[0].w, fe->g[0].h, (tb_match *)matches, capacity, (int *)counts);
}

// Matcher::rejectWithF behind the flow entries (matcher.cpp:751-755, :853-881): "last" points = keypoints of the frame the tracks
// started from (cv::KeyPoint: pt first, 28-byte stride)
extern "C" int tb_frontend_reject_with_f(tb_frontend *fe, const tb_ransac_params *p, int n_pairs, const int32_t *f2_frames,
                                         const void *next_pts, void *status, void *F_out, void *iters_out)
{
    TB_DEVICE_GUARD(fe ? fe->ctx : nullptr);
    if (!fe || !p || n_pairs < 0 || (n_pairs && (!f2_frames || !next_pts || !status))) return TB_E_INVALID;
    std::vector<RfPairHost> hp(n_pairs);
    for (int i = 0; i < n_pairs; i++) {
        const int f = f2_frames[i];
        if (f < 0 || f >= fe->B) return tb_fail(fe->ctx, TB_E_INVALID, "tb_frontend_reject_with_f: frame index out of range");
        hp[i].cur = (const float *)next_pts + 2 * (size_t)i * fe->kp_cap;
        hp[i].last = (const uint8_t *)(fe->kps + (size_t)f * fe->kp_cap);
        hp[i].last_stride = (int)sizeof(tb_keypoint);
        hp[i].count = fe->n_kp + f;
        hp[i].status = (uint8_t *)status + (size_t)i * fe->kp_cap;
    }
    return tb_reject_with_f_batched(fe->ctx, p, n_pairs, hp, fe->kp_cap, (double *)F_out, (int32_t *)iters_out);
}
