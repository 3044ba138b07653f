// reach_common.cuh -- context, error plumbing and small device-memory helpers shared by the
// LGG09 and GLGM06 translation units of libreach_cuda.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include <new>

#include "../../include/reach.h"

struct reach_ctx {
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    char name[256] = {0};
    std::string err;
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

// RAII device buffer (freed with the plan that owns it)
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    int alloc(reach_ctx* ctx, size_t n, bool zero = true) {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            return set_error(ctx, REACH_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", n,
                             cudaGetErrorString(e));
        }
        bytes = n;
        if (zero) {
            e = cudaMemsetAsync(p, 0, n, ctx->stream);
            if (e != cudaSuccess)
                return set_error(ctx, REACH_ECUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        }
        return REACH_OK;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace reach
