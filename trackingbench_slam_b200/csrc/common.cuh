// common.cuh — shared declarations for libtb_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/tb_capi.h"

#ifndef __CUDA_ARCH__
#define TB_HOST_ONLY 1
#endif

// Device scratch of the host-facing entry points (the facade's per-call work buffers): chunks are kept between calls and
// rewound when a call ends, so a steady caller performs no cudaMalloc/cudaFree; a call that outgrew the first chunk leaves
// several chunks behind, which the next rewind merges into one (after making sure the stream no longer uses them).
struct DevArena {
    struct Chunk { char *p; size_t cap, used; };
    std::vector<Chunk> chunks;
    void *alloc(size_t bytes)
    {
        bytes = (bytes + 255) & ~(size_t)255;
        if (!bytes) bytes = 256;
        for (Chunk &c : chunks)
            if (c.cap - c.used >= bytes) { void *r = c.p + c.used; c.used += bytes; return r; }
        size_t total = 0;
        for (Chunk &c : chunks) total += c.cap;
        Chunk n;
        n.cap = bytes > total ? bytes : total;            // geometric growth
        if (n.cap < ((size_t)1 << 20)) n.cap = (size_t)1 << 20;
        n.used = bytes;
        void *q = nullptr;
        if (cudaMalloc(&q, n.cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        n.p = (char *)q;
        chunks.push_back(n);
        return n.p;
    }
    void rewind(cudaStream_t stream)
    {
        if (chunks.size() > 1) {
            size_t total = 0;
            for (Chunk &c : chunks) total += c.cap;
            cudaStreamSynchronize(stream);
            for (Chunk &c : chunks) cudaFree(c.p);
            chunks.clear();
            void *q = nullptr;
            if (cudaMalloc(&q, total) == cudaSuccess) { Chunk n = {(char *)q, total, 0}; chunks.push_back(n); } else cudaGetLastError();
        }
        for (Chunk &c : chunks) c.used = 0;
    }
    void destroy() { for (Chunk &c : chunks) cudaFree(c.p); chunks.clear(); }
};

// Pinned staging ring for the small host tables an entry point hands to its kernels (frame-pair lists, poses): the table is
// copied into a pinned slot and uploaded asynchronously, so the entry point does not have to wait for the copy before its
// pageable source goes out of scope (no cudaStreamSynchronize between extract and read-back).
struct StageRing {
    enum { SLOTS = 8 };
    struct Slot { char *p = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool busy = false; };
    Slot slot[SLOTS];
    int next = 0;
    void destroy() { for (Slot &s : slot) { if (s.p) cudaFreeHost(s.p); if (s.ev) cudaEventDestroy(s.ev); s = Slot(); } }
};

struct tb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // side stream + events: the Gaussian levels run here, concurrently with the latency-bound octree kernel
    cudaStream_t side = nullptr, side2 = nullptr;   // side: Gaussian levels; side2: per-level octree launches behind FAST
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    cudaEvent_t ev_lvl[TB_MAX_LEVELS] = {};
    cudaEvent_t ev_pyr[TB_MAX_LEVELS] = {};   // level l of a pyramid built on the side stream is complete
    bool pyr_pending = false;                 // the main stream has not yet waited for the levels built on the side stream
    int pyr_last = 0;                         // index of the last ev_pyr recorded by that build
    int num_sms = 148;
    int64_t launches = 0;
    std::string err;
    struct MatchWs *mws = nullptr;   // lazily grown matching workspace (match.cu)
    struct TcWs *tcws = nullptr;     // tensor-core matcher workspace (match_tc.cu)
    struct WinWs *wws = nullptr;     // batched windowed-matcher workspace (window.cu)
    struct IngestWs *iws = nullptr;  // CLAHE tile LUTs (ingest.cu)
    struct FlowWs *fws = nullptr;    // LK pyramids + pair descriptors (flow.cu)
    struct RfWs *rfws = nullptr;     // rejectWithF: pair descriptors + compacted tracks (fundam.cu)
    int match_impl = 1;              // 1 = tcgen05 int8 (K8, the default), 0 = LOP3+POPC (K7)
    DevArena arena;                  // scratch of the host-facing entry points
    StageRing stage;                 // pinned staging of small host tables
    // optional per-kernel timing (bench.py's roofline leg): CUDA events around every launch
    bool profiling = false;
    struct ProfEvt { const char *name; cudaEvent_t a, b; };
    std::vector<ProfEvt> prof;
};
// RAII: records an event pair around the launches issued in its scope when profiling is on.
struct ProfScope {
    tb_ctx *c; int idx;
    ProfScope(tb_ctx *ctx, const char *name) : c(ctx), idx(-1)
    {
        if (!c->profiling) return;
        tb_ctx::ProfEvt e; e.name = name;
        cudaEventCreate(&e.a); cudaEventCreate(&e.b);
        cudaEventRecord(e.a, c->stream);
        idx = (int)c->prof.size();
        c->prof.push_back(e);
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof[idx].b, c->stream); }
};
void tb_match_ws_free(tb_ctx *ctx);
void tb_match_tc_free(tb_ctx *ctx);
void tb_window_ws_free(tb_ctx *ctx);
void tb_ingest_ws_free(tb_ctx *ctx);
void tb_flow_ws_free(tb_ctx *ctx);
void tb_fund_ws_free(tb_ctx *ctx);

static inline int tb_fail(tb_ctx *ctx, int code, const char *fmt, const char *a = "", const char *b = "")
{
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b);
    if (ctx) ctx->err = buf;
    return code;
}

// every extern "C" entry point makes its ctx's device current first: two ctxs on different GPUs may share a host thread, and the
// caller (e.g. torch) may have switched devices between calls
// ... and waits (on the ctx's main stream) for a pyramid that tb_frontend_build_pyramid is still building on the side stream:
// only tb_frontend_extract lets its FAST launches wait level by level instead (TB_DEVICE_GUARD_NOJOIN)
static inline void tb_device_guard(tb_ctx *c, bool join_pyramid)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (join_pyramid && c->pyr_pending) {
        cudaStreamWaitEvent(c->stream, c->ev_pyr[c->pyr_last], 0);
        c->pyr_pending = false;
    }
}
#define TB_DEVICE_GUARD(c) tb_device_guard(const_cast<tb_ctx *>(static_cast<const tb_ctx *>(c)), true)
#define TB_DEVICE_GUARD_NOJOIN(c) tb_device_guard(const_cast<tb_ctx *>(static_cast<const tb_ctx *>(c)), false)

#define TB_CUDA(ctx, expr)                                                                    \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) return tb_fail((ctx), TB_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

#define TB_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) return tb_fail((ctx), TB_E_CUDA, "kernel launch: %s", cudaGetErrorString(e__)); \
    } while (0)

// 3-D (x, y, frame) u8 tensor map for TMA tile loads of boxw x boxh x 1 (zero fill outside the image). The driver entry
// point is fetched through the runtime so libcuda is not a link dependency. base, pitch and frame stride must be
// multiples of 16 bytes; at run time the box origin must be a multiple of 16 in x (an unaligned origin faults).
static inline bool tb_encode_tiled_u8_3d(CUtensorMap *tm, const void *base, int w, int h, int n, size_t pitch, size_t fstride, int boxw, int boxh)
{
    typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static PFN encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
        encode = (PFN)fn;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (fstride & 15) || boxw > 256 || boxh > 256 || (boxw & 15)) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    const cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1}, estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// cudaMemcpyAsync(dst_dev <- src_host) of a small pageable table through the ctx's pinned staging ring (returns at once)
static inline int tb_stage_upload(tb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!bytes) return TB_OK;
    StageRing::Slot &s = ctx->stage.slot[ctx->stage.next];
    ctx->stage.next = (ctx->stage.next + 1) % StageRing::SLOTS;
    if (s.busy) { TB_CUDA(ctx, cudaEventSynchronize(s.ev)); s.busy = false; }
    if (bytes > s.cap) {
        if (s.p) cudaFreeHost(s.p);
        s.p = nullptr; s.cap = 0;
        const size_t cap = bytes < 65536 ? 65536 : bytes;
        TB_CUDA(ctx, cudaHostAlloc((void **)&s.p, cap, cudaHostAllocDefault));
        s.cap = cap;
    }
    if (!s.ev) TB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
    memcpy(s.p, src_host, bytes);
    TB_CUDA(ctx, cudaMemcpyAsync(dst_dev, s.p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(ctx, cudaEventRecord(s.ev, ctx->stream));
    s.busy = true;
    return TB_OK;
}
// Grow path of a persistent workspace buffer: the old buffer is released and the pointer cleared BEFORE the new allocation,
// so that a failed cudaMalloc leaves a null pointer (callers reset their capacity first and set it only after success).
template <typename T> static inline int tb_regrow(tb_ctx *ctx, T *&p, size_t n_elems)
{
    if (p) { cudaFree(p); p = nullptr; }
    TB_CUDA(ctx, cudaMalloc((void **)&p, sizeof(T) * (n_elems ? n_elems : 1)));
    return TB_OK;
}
// RAII view of the ctx arena for one host-facing call: typed bump allocations, rewound when the call returns
struct ArenaScope {
    tb_ctx *c;
    explicit ArenaScope(tb_ctx *ctx) : c(ctx) {}
    ~ArenaScope() { c->arena.rewind(c->stream); }
    template <typename T> T *alloc(size_t n) { return (T *)c->arena.alloc((n ? n : 1) * sizeof(T) + 64); }
    template <typename T> T *upload(const T *h, size_t n)
    {
        T *d = alloc<T>(n);
        if (d && n) cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream);
        return d;
    }
};

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Packed FAST candidate: y:12 | x:12 | m:8, coordinates relative to the (16,16) border.
// m = max(A,B) of the FAST-9 arcs; response = m - 1.
__host__ __device__ static inline uint32_t cand_pack(int x, int y, int m) { return ((uint32_t)y << 20) | ((uint32_t)x << 8) | (uint32_t)m; }
__host__ __device__ static inline int cand_x(uint32_t c) { return (c >> 8) & 0xFFF; }
__host__ __device__ static inline int cand_y(uint32_t c) { return c >> 20; }
__host__ __device__ static inline int cand_m(uint32_t c) { return c & 0xFF; }

// Per-level geometry of the ORB cell grid (ORBextractor.cpp:743-763), computed on the host.
struct LevelGeom {
    int w, h;           // level size
    int pitch;          // row pitch in bytes
    size_t frame_stride;
    int nCols, nRows, wCell, hCell;
    int nColsRun, nRowsRun; // cells not skipped by the iniX/iniY tests (:770,:779)
    int maxBX, maxBY;   // w-16, h-16
    int cand_cap;       // candidate slots per (frame, level)
    int nIni;           // initial octree nodes (:498)
    float hX;           // (:500)
};
