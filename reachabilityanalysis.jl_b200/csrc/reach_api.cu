// reach_api.cu -- context management of libreach_cuda.so (see include/reach.h).
#include "reach_common.cuh"

namespace reach {
thread_local std::string g_create_error;
}
using namespace reach;

extern "C" int reach_version(void) { return 100; /* 0.1.0 */ }

extern "C" int reach_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error(nullptr, REACH_ECUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        return REACH_ECUDA;
    }
    return n;
}

extern "C" int reach_create(reach_ctx** out, int device) {
    if (!out) return set_error(nullptr, REACH_EINVAL, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(nullptr, REACH_ECUDA,
                         "no usable CUDA device (%s); libreach_cuda has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count)
        return set_error(nullptr, REACH_EINVAL, "device %d out of range [0, %d)", device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(nullptr, REACH_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_error(nullptr, REACH_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return set_error(nullptr, REACH_EUNSUPPORTED,
                         "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                         prop.major, prop.minor);
    reach_ctx* ctx = new (std::nothrow) reach_ctx;
    if (!ctx) return set_error(nullptr, REACH_ENOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    snprintf(ctx->name, sizeof(ctx->name), "%s", prop.name);
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return set_error(nullptr, REACH_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return REACH_OK;
}

extern "C" void reach_destroy(reach_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx_cache_flush(ctx);
    delete ctx;
}

extern "C" const char* reach_last_error(const reach_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int reach_device_name(const reach_ctx* ctx, char* buf, size_t len) {
    if (!ctx || !buf || len == 0) return REACH_EINVAL;
    snprintf(buf, len, "%s", ctx->name);
    return REACH_OK;
}

extern "C" int reach_sm_count(const reach_ctx* ctx) { return ctx ? ctx->sm_count : REACH_EINVAL; }

extern "C" void* reach_stream(const reach_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int reach_synchronize(reach_ctx* ctx) {
    if (!ctx) return REACH_EINVAL;
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REACH_OK;
}
