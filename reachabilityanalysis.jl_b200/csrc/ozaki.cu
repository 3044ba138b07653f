// ozaki.cu -- host side of the INT8-tcgen05 FP64-emulation GEMM (oz_gemm.cuh): the stand-alone probe used by
// tests and bench.py.  The LGG09 main loop launches the same kernels from lgg09.cu.
#include "reach_common.cuh"
#include "oz_gemm.cuh"
#include "crt_gemm.cuh"

#include <vector>
#include <stdlib.h>

using namespace reach;

// C[m x n] (row-major) = X[m x k] * Y[n x k]'  with X, Y row-major FP64 host matrices, through T digit planes.
// C_host may be NULL (timing only).  *tflops = 2 m n k iters / time of the GEMM launches (slicing excluded),
// *slice_ms = time of slicing both operands once.
extern "C" int reach_ozgemm_probe(reach_ctx* ctx, int m, int n, int k, int slices, const double* X_host,
                                  const double* Y_host, double* C_host, int iters, double* tflops, double* slice_ms) {
    if (!ctx) return REACH_EINVAL;
    if (m < 1 || n < 1 || k < 1 || iters < 1) return set_error(ctx, REACH_EINVAL, "probe sizes must be positive");
    if (slices < 4 || slices > OZ_MAX_T) return set_error(ctx, REACH_EINVAL, "slices must be in 4..%d", OZ_MAX_T);
    if (k > OZ_MAX_K) return set_error(ctx, REACH_EINVAL, "k must be <= %d (INT32 accumulator headroom)", OZ_MAX_K);
    if (!X_host || !Y_host) return set_error(ctx, REACH_EINVAL, "X and Y must not be NULL");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int T = slices;
    const int tiles_m = (m + OZ_TM - 1) / OZ_TM, tiles_n = (n + OZ_TN - 1) / OZ_TN, nkb = (k + OZ_KB - 1) / OZ_KB;
    DevBuf X, Y, C, Xq, Yq, eX, eY;
    REACH_TRY(X.alloc(ctx, size_t(m) * k * 8, false));
    REACH_TRY(Y.alloc(ctx, size_t(n) * k * 8, false));
    REACH_TRY(C.alloc(ctx, size_t(m) * n * 8));
    REACH_TRY(Xq.alloc(ctx, size_t(nkb) * tiles_m * T * OZ_TM * OZ_KB));
    REACH_TRY(Yq.alloc(ctx, size_t(nkb) * tiles_n * T * OZ_TN * OZ_KB));
    REACH_TRY(eX.alloc(ctx, size_t(tiles_m) * OZ_TM * 4));
    REACH_TRY(eY.alloc(ctx, size_t(tiles_n) * OZ_TN * 4));
    REACH_CUDA(ctx, cudaMemcpyAsync(X.p, X_host, size_t(m) * k * 8, cudaMemcpyHostToDevice, st));
    REACH_CUDA(ctx, cudaMemcpyAsync(Y.p, Y_host, size_t(n) * k * 8, cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1, e2;
    REACH_CUDA(ctx, cudaEventCreate(&e0));
    REACH_CUDA(ctx, cudaEventCreate(&e1));
    REACH_CUDA(ctx, cudaEventCreate(&e2));
    REACH_CUDA(ctx, cudaEventRecord(e0, st));
    oz_slice_kernel<OZ_TM><<<tiles_m * OZ_TM / 8, 256, 0, st>>>(X.as<double>(), k, m, k, T, nkb, tiles_m, Xq.as<int8_t>(), eX.as<int>());
    REACH_CUDA(ctx, cudaGetLastError());
    oz_slice_kernel<OZ_TN><<<tiles_n * OZ_TN / 8, 256, 0, st>>>(Y.as<double>(), k, n, k, T, nkb, tiles_n, Yq.as<int8_t>(), eY.as<int>());
    REACH_CUDA(ctx, cudaGetLastError());
    REACH_CUDA(ctx, cudaEventRecord(e1, st));
    OzOperands g{Xq.as<int8_t>(), Yq.as<int8_t>(), eX.as<int>(), eY.as<int>(), nkb, tiles_m, tiles_m, 0, tiles_n, tiles_n};
    OzStoreEpi epi{C.as<double>(), n, m, n};
    REACH_CUDA(ctx, (launch_oz_gemm(T, g, epi, st)));          // warm-up (also the result)
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    REACH_CUDA(ctx, cudaEventRecord(e1, st));
    for (int i = 0; i < iters; ++i) REACH_CUDA(ctx, (launch_oz_gemm(T, g, epi, st)));
    REACH_CUDA(ctx, cudaEventRecord(e2, st));
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    float ms_slice = 0.f, ms = 0.f;
    REACH_CUDA(ctx, cudaEventElapsedTime(&ms, e1, e2));
    (void)ms_slice;
    if (slice_ms) {   // time the slicing separately
        REACH_CUDA(ctx, cudaEventRecord(e0, st));
        oz_slice_kernel<OZ_TM><<<tiles_m * OZ_TM / 8, 256, 0, st>>>(X.as<double>(), k, m, k, T, nkb, tiles_m, Xq.as<int8_t>(), eX.as<int>());
        oz_slice_kernel<OZ_TN><<<tiles_n * OZ_TN / 8, 256, 0, st>>>(Y.as<double>(), k, n, k, T, nkb, tiles_n, Yq.as<int8_t>(), eY.as<int>());
        REACH_CUDA(ctx, cudaEventRecord(e1, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
        float s = 0.f;
        REACH_CUDA(ctx, cudaEventElapsedTime(&s, e0, e1));
        *slice_ms = s;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    if (tflops) *tflops = 2.0 * m * double(n) * k * iters / (ms * 1e-3) / 1e12;
    if (C_host) REACH_CUDA(ctx, cudaMemcpy(C_host, C.p, size_t(m) * n * 8, cudaMemcpyDeviceToHost));
    return REACH_OK;
}

// The same product through the residue (Chinese remainder) path of crt_gemm.cuh with `moduli` INT8 GEMMs per FP64 product.
// *tflops counts 2 m n k per launch (FP64-equivalent), *slice_ms the residue conversion.
extern "C" int reach_crtgemm_probe(reach_ctx* ctx, int m, int n, int k, int moduli, const double* X_host,
                                   const double* Y_host, double* C_host, int iters, double* tflops, double* slice_ms) {
    if (!ctx) return REACH_EINVAL;
    if (m < 1 || n < 1 || k < 1 || iters < 1) return set_error(ctx, REACH_EINVAL, "probe sizes must be positive");
    if (moduli < CRT_MIN_N || moduli > CRT_MAX_N) return set_error(ctx, REACH_EINVAL, "moduli must be in %d..%d", CRT_MIN_N, CRT_MAX_N);
    if (k > CRT_MAX_K) return set_error(ctx, REACH_EINVAL, "k must be <= %d (INT32 accumulator headroom)", CRT_MAX_K);
    if (!X_host || !Y_host) return set_error(ctx, REACH_EINVAL, "X and Y must not be NULL");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const CrtConsts cc = crt_make_consts(moduli);
    const int tiles_m = (m + 127) / 128, tiles_n = (n + 127) / 128, nkb = (k + CRT_KB - 1) / CRT_KB;
    DevBuf X, Y, C, Xr, Yr, eX, eY, scr;
    REACH_TRY(X.alloc(ctx, size_t(m) * k * 8, false));
    REACH_TRY(Y.alloc(ctx, size_t(n) * k * 8, false));
    REACH_TRY(C.alloc(ctx, size_t(m) * n * 8));
    REACH_TRY(Xr.alloc(ctx, size_t(moduli) * tiles_m * nkb * CRT_IMG));
    REACH_TRY(Yr.alloc(ctx, size_t(moduli) * tiles_n * nkb * CRT_IMG));
    REACH_TRY(eX.alloc(ctx, size_t(tiles_m) * 128 * 4));
    REACH_TRY(eY.alloc(ctx, size_t(tiles_n) * 128 * 4));
    REACH_TRY(scr.alloc(ctx, crt_scratch_bytes(moduli, ctx->sm_count), false));
    REACH_CUDA(ctx, cudaMemcpyAsync(X.p, X_host, size_t(m) * k * 8, cudaMemcpyHostToDevice, st));
    REACH_CUDA(ctx, cudaMemcpyAsync(Y.p, Y_host, size_t(n) * k * 8, cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1, e2;
    REACH_CUDA(ctx, cudaEventCreate(&e0));
    REACH_CUDA(ctx, cudaEventCreate(&e1));
    REACH_CUDA(ctx, cudaEventCreate(&e2));
    auto slice_both = [&]() {
        launch_crt_slice(X.as<double>(), k, m, tiles_m * 128, k, cc, false, nkb, tiles_m, Xr.as<int8_t>(), eX.as<int>(), 0, 1, st);
        launch_crt_slice(Y.as<double>(), k, n, tiles_n * 128, k, cc, true, nkb, tiles_n, Yr.as<int8_t>(), eY.as<int>(), 0, 1, st);
    };
    slice_both();
    REACH_CUDA(ctx, cudaGetLastError());
    CrtOperands g{Xr.as<int8_t>(), Yr.as<int8_t>(), eX.as<int>(), eY.as<int>(), nkb, tiles_m, tiles_m, 0, tiles_n, tiles_n,
                  scr.as<uint32_t>(), cc};
    OzStoreEpi epi{C.as<double>(), n, m, n};
    DevBuf dbg;
    const char* dbg_env = getenv("REACH_CRT_DEBUG");           // "<modulus index>": raw accumulators -> <C_host as int32 [2][128][256]>
    if (dbg_env && size_t(m) * n * 8 >= size_t(2) * 128 * 256 * 4) {
        REACH_TRY(dbg.alloc(ctx, size_t(2) * 128 * 256 * 4));
        g.dbg = dbg.as<int32_t>();
        g.dbg_mod = atoi(dbg_env);
    }
    if (getenv("REACH_CRT_FLAGS")) g.flags = atoi(getenv("REACH_CRT_FLAGS"));
    REACH_CUDA(ctx, (launch_crt_gemm(g, epi, st)));            // warm-up (also the result)
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    REACH_CUDA(ctx, cudaEventRecord(e1, st));
    for (int i = 0; i < iters; ++i) REACH_CUDA(ctx, (launch_crt_gemm(g, epi, st)));
    REACH_CUDA(ctx, cudaEventRecord(e2, st));
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    REACH_CUDA(ctx, cudaEventElapsedTime(&ms, e1, e2));
    if (slice_ms) {
        REACH_CUDA(ctx, cudaEventRecord(e0, st));
        slice_both();
        REACH_CUDA(ctx, cudaEventRecord(e1, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
        float s = 0.f;
        REACH_CUDA(ctx, cudaEventElapsedTime(&s, e0, e1));
        *slice_ms = s;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    if (tflops) *tflops = 2.0 * m * double(n) * k * iters / (ms * 1e-3) / 1e12;
    if (C_host) REACH_CUDA(ctx, cudaMemcpy(C_host, C.p, size_t(m) * n * 8, cudaMemcpyDeviceToHost));
    if (C_host && g.dbg) REACH_CUDA(ctx, cudaMemcpy(C_host, dbg.p, size_t(2) * 128 * 256 * 4, cudaMemcpyDeviceToHost));
    return REACH_OK;
}

// Host-only: the constants of the residue path for `moduli` moduli (crt_make_consts), so that a CPU test can pin them
// against exact integer arithmetic: the moduli p_i, the split fractions f_i = f1_i + f2_i of w_i / P, P rounded to FP64 and the
// norm budgets 2^alpha, 2^beta of the X and Y rows.  Arrays hold `moduli` entries.  No CUDA call is made.
extern "C" int reach_crt_constants(int moduli, int32_t* p, double* f1, double* f2, double* P, int32_t* alpha, int32_t* beta) {
    if (moduli < CRT_MIN_N || moduli > CRT_MAX_N) return REACH_EINVAL;
    const CrtConsts c = crt_make_consts(moduli);
    for (int i = 0; i < moduli; ++i) {
        if (p) p[i] = int32_t(c.p[i]);
        if (f1) f1[i] = c.f1[i];
        if (f2) f2[i] = c.f2[i];
    }
    if (P) *P = c.Pd;
    if (alpha) *alpha = c.alpha;
    if (beta) *beta = c.beta;
    return REACH_OK;
}
