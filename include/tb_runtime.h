// tb_runtime.h — process-wide tb_ctx used by the header-only facade classes (one device, one stream,
// like the single-threaded reference).  TB_DEVICE in the environment selects the GPU.
#ifndef TB_RUNTIME_H
#define TB_RUNTIME_H
#include <cstdlib>
#include <stdexcept>
#include <string>
#include "tb_capi.h"
namespace TRACKING_BENCH {
inline tb_ctx *tb_default_ctx()
{
    static tb_ctx *ctx = nullptr;
    if (!ctx) {
        const char *d = std::getenv("TB_DEVICE");
        if (tb_ctx_create(d ? std::atoi(d) : 0, nullptr, &ctx) != TB_OK)
            throw std::runtime_error("trackingbench_slam_b200: no usable CUDA device (there is no CPU fallback)");
    }
    return ctx;
}
inline void tb_check(int rc, tb_ctx *ctx, const char *what)
{
    if (rc != TB_OK) throw std::runtime_error(std::string(what) + ": " + tb_last_error(ctx));
}
}  // namespace TRACKING_BENCH
#endif
