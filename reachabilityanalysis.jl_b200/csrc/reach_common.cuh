// reach_common.cuh -- context, error plumbing and small device-memory helpers shared by the
// LGG09 and GLGM06 translation units of libreach_cuda.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <map>
#include <string>
#include <vector>
#include <new>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, no-ops otherwise

#include "../../include/reach.h"

struct reach_ctx {
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    char name[256] = {0};
    std::string err;
    // Device-memory cache: cudaMalloc / cudaFree of the multi-GB plan buffers cost ~70 ms each per
    // one-shot call; freed blocks are kept (size-keyed) and handed to the next plan of this ctx.
    std::multimap<size_t, void*> free_blocks;
    size_t cached_bytes = 0;
    size_t cache_limit = size_t(48) << 30;
};

namespace reach {

extern thread_local std::string g_create_error;

inline int set_error(reach_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define REACH_CUDA(ctx, call)                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return reach::set_error((ctx), e__ == cudaErrorMemoryAllocation ? REACH_ENOMEM      \
                                                                             : REACH_ECUDA,     \
                                    "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                                    __FILE__, __LINE__);                                        \
    } while (0)

#define REACH_TRY(expr)            \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != REACH_OK) return rc__; \
    } while (0)

inline void ctx_cache_flush(reach_ctx* ctx) {
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    ctx->free_blocks.clear();
    ctx->cached_bytes = 0;
}

// RAII device buffer (returned to the ctx cache with the plan that owns it).  All work on a ctx is
// ordered on its single stream, so a cached block can be reused without further synchronisation.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;       // requested
    size_t cap = 0;         // allocated
    reach_ctx* owner = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            if (owner && owner->cached_bytes + cap <= owner->cache_limit) {
                owner->free_blocks.emplace(cap, p);
                owner->cached_bytes += cap;
            } else {
                cudaFree(p);
            }
        }
        p = nullptr;
        bytes = cap = 0;
    }
    int alloc(reach_ctx* ctx, size_t n, bool zero = true) {
        release();
        if (n == 0) n = 16;
        const size_t gran = size_t(1) << 20;
        const size_t want = (n + gran - 1) / gran * gran;
        owner = ctx;
        auto it = ctx->free_blocks.lower_bound(want);
        if (it != ctx->free_blocks.end() && it->first <= want + want / 4 + gran) {
            p = it->second;
            cap = it->first;
            ctx->cached_bytes -= cap;
            ctx->free_blocks.erase(it);
        } else {
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {             // make room: drop the cache and retry once
                cudaGetLastError();
                ctx_cache_flush(ctx);
                e = cudaMalloc(&p, want);
            }
            if (e != cudaSuccess) {
                p = nullptr;
                return set_error(ctx, REACH_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            }
            cap = want;
        }
        bytes = n;
        if (zero) {
            cudaError_t e = cudaMemsetAsync(p, 0, n, ctx->stream);
            if (e != cudaSuccess)
                return set_error(ctx, REACH_ECUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        }
        return REACH_OK;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// NVTX range for the scope (phases of plan_run: upload, stack build, main loop, chain, batched product, selection ...)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace reach
