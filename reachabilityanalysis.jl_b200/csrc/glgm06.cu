// glgm06.cu -- host side of the GLGM06 device path and its C ABI (include/reach.h).
//
// Replaces the inside of reach_homog_GLGM06! (src/Algorithms/GLGM06/reach_homog.jl:33-73) and
// reach_inhomog_GLGM06! (reach_inhomog.jl:6-43), including the initial reduce_order of
// GLGM06/post.jl:26, for a batch of initial sets that share Phi and U (the partitioned-initial-set
// caller, src/Continuous/solve.jl:84-117).
#include "reach_common.cuh"
#include "glgm06_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace reach;

namespace {

int next_pow2(int x) {
    int p = 32;
    while (p < x) p <<= 1;
    return p;
}

template <class Epi>
cudaError_t launch_tn64(const GemmOperands& g, const Epi& epi, cudaStream_t st) {
    // K <= 48 (Building): three k-tiles in all, so a 3-stage ring holds the whole operand panel and the smaller
    // footprint (61 KB instead of 82) lets a third CTA share the SM: prologue, L2 round trip and store epilogue of one
    // CTA overlap the MMA burst of the others (C4: 19.8 -> 17 ms; FP64-pipe floor incl. the 64-column block padding
    // 10.8 ms).  A resident-X "panel" kernel walking several Y panels per CTA was measured slower (28.5 vs 26.9 ms).
    if (g.K <= 3 * BK) return launch_nt_dgemm<2, 2, 4, 4, 3>(g, epi, st);
    return launch_nt_dgemm<2, 2, 4, 4, 4>(g, epi, st);     // 64 x 64 tiles
}
// 64 x TN tiles for the batched product at n <= 64: TN = n rounded up to 16, so that a power block has no padded
// columns to multiply (n = 48: 48 instead of 64, a quarter of the DMMA work of that launch)
template <class Epi>
cudaError_t launch_small_tn(int TN, const GemmOperands& g, const Epi& epi, cudaStream_t st) {
    const bool short_k = g.K <= 3 * BK;
    switch (TN) {
        case 16: return short_k ? launch_nt_dgemm<2, 2, 4, 1, 3>(g, epi, st) : launch_nt_dgemm<2, 2, 4, 1, 4>(g, epi, st);
        case 32: return short_k ? launch_nt_dgemm<2, 2, 4, 2, 3>(g, epi, st) : launch_nt_dgemm<2, 2, 4, 2, 4>(g, epi, st);
        case 48: return short_k ? launch_nt_dgemm<2, 2, 4, 3, 3, 4>(g, epi, st) : launch_nt_dgemm<2, 2, 4, 3, 4>(g, epi, st);
        case 64: return launch_tn64(g, epi, st);
    }
    return cudaErrorInvalidValue;
}
template <class Epi>
cudaError_t launch_tn128(const GemmOperands& g, const Epi& epi, cudaStream_t st) {
    return launch_nt_dgemm_ws<2, 4, 8, 4, 5>(g, epi, st);  // 128 x 128 tiles, warp-specialised
}

constexpr int kMaxSort = 4096;     // generators ranked by one CTA (shared-memory bitonic sort)

}  // namespace

struct reach_glgm06_plan {
    reach_ctx* ctx = nullptr;
    int n = 0, nsplits = 0, r = 0, p0 = 0, pU = -1;
    int64_t nsteps = 0;
    bool homog = true;

    int Kp = 0, nb = 0, TN = 64, TM = 64;
    int p0r = 0;            // generators per split after the initial reduce_order (>= 1 when homogeneous)
    int rows = 0;           // p0r + 1 (generators + centre)
    int M = 0, m_pad = 0;
    int64_t nblk = 0;       // steps after the first (nsteps - 1)
    int pWcap = 0, lda = 0, tpb = 1;
    std::vector<int> pW;    // shared generator count used at block b (size nblk + 1)
    bool uploaded = false, ran = false;
    bool force_stepwise = false;   // REACH_GLGM06_CHAIN=stepwise: use the per-step kernels even for small n (tests)

    DevBuf phi_cm, PhiT, PhiRows;       // column-major copy; rows = columns of Phi; rows = rows of Phi
    DevBuf Xown;                        // [m_pad][Kp]
    DevBuf XU, PU;                      // [(pU+1) padded to 64][Kp]
    DevBuf Pstack;                      // [nblk][nb][Kp]
    DevBuf GWall, cWall;                // [(nblk+1)][pWcap][n], [(nblk+1)][n]; GWall doubles as the generator pool
    DevBuf slotB;                       // [(nblk+1)][pWcap] pool row of column r of W_b (identity in the stepwise path)
    DevBuf wB, permB, bpref, pWdev;     // per block
    DevBuf A_store;                     // [nblk][m_pad][lda]
    DevBuf psum, pmax;                  // [nblk*tpb][m_pad]
    DevBuf pos_own, tb, diag, centre;   // per (block, split)
    DevBuf H;                           // homogeneous: [nsteps][m_pad][Kp]
    DevBuf scratch_pos;
    DevBuf progress;                    // powers published so far (chain kernel <- powers kernel)
    DevBuf dbg;                         // phase cycle counters of the chain kernel (REACH_GLGM06_DEBUG)
    int64_t bytes_stored = 0;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_powers = nullptr, ev_gemm = nullptr;
    cudaStream_t stream2 = nullptr;     // the batched GEMM runs here, beside the W chain on the ctx stream
    double last_ms = 0.0;
    int64_t launches = 0;

    ~reach_glgm06_plan() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_powers) cudaEventDestroy(ev_powers);
        if (ev_gemm) cudaEventDestroy(ev_gemm);
        if (stream2) cudaStreamDestroy(stream2);
    }
};

extern "C" int reach_glgm06_plan_create(reach_ctx* ctx, int n, int64_t nsteps, int nsplits, int max_order, int p0,
                                        int pU, reach_glgm06_plan** out) {
    if (!ctx) return set_error(nullptr, REACH_EINVAL, "ctx is NULL");
    if (!out) return set_error(ctx, REACH_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 1) return set_error(ctx, REACH_EINVAL, "state dimension must be >= 1, got %d", n);
    if (nsteps < 1) return set_error(ctx, REACH_EINVAL, "NSTEPS must be >= 1, got %lld", (long long)nsteps);
    if (nsplits < 1) return set_error(ctx, REACH_EINVAL, "need at least one initial set, got %d", nsplits);
    if (max_order < 1)
        return set_error(ctx, REACH_EINVAL, "the target order should be at least 1, but it is %d", max_order);
    if (p0 < 0) return set_error(ctx, REACH_EINVAL, "negative generator count");
    if (int64_t(max_order) * n + std::max(p0, std::max(pU, 0)) > kMaxSort || p0 > kMaxSort)
        return set_error(ctx, REACH_EUNSUPPORTED,
                         "order reduction of more than %d generators per zonotope is not supported "
                         "(max_order*n + p = %lld)", kMaxSort, (long long)max_order * n + std::max(p0, std::max(pU, 0)));
    reach_glgm06_plan* p = new (std::nothrow) reach_glgm06_plan;
    if (!p) return set_error(ctx, REACH_ENOMEM, "out of host memory");
    p->ctx = ctx;
    p->n = n;
    p->nsteps = nsteps;
    p->nsplits = nsplits;
    p->r = max_order;
    p->p0 = p0;
    p->pU = pU;
    p->homog = pU < 0;
    p->Kp = round_up(n, BK);
    p->TN = n <= 64 ? round_up(n, 16) : 128;
    p->TM = n <= 64 ? 64 : 128;
    p->nb = round_up(n, p->TN);
    p->tpb = p->nb / p->TN;
    p->p0r = p0 > max_order * n ? max_order * n : p0;
    if (p->homog && p->p0r == 0) p->p0r = 1;            // one zero generator (reach_homog.jl:52-54)
    p->rows = p->p0r + 1;
    p->M = nsplits * p->rows;
    p->m_pad = round_up(p->M, 128);
    p->nblk = nsteps - 1;
    p->lda = round_up(n, 2);
    {
        const char* env = getenv("REACH_GLGM06_CHAIN");
        p->force_stepwise = env && std::strcmp(env, "stepwise") == 0;
    }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_powers, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_gemm, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete p;
        return set_error(ctx, REACH_ECUDA, "plan_create: %s", cudaGetErrorString(e));
    }
    *out = p;
    return REACH_OK;
}

extern "C" void reach_glgm06_plan_destroy(reach_glgm06_plan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    if (p->stream2) cudaStreamSynchronize(p->stream2);
    cudaStreamSynchronize(p->ctx->stream);
    delete p;
}

namespace {

// reduce_order of a batch of zonotopes whose generators are rows [batch][p][ld_in]; writes
// [batch][r n] rows with stride out_stride (doubles) between zonotopes and ld_out per row.
int launch_reduce_batch(reach_ctx* ctx, const double* Gin, int ld_in, int p, int n, int r, int batch, double* Gout,
                        int ld_out, long long out_stride, int* sel, DevBuf& scratch) {
    // the kernel assumes contiguous output per zonotope with ld == ld_in; reduce into a temp when strides differ
    const int np2 = next_pow2(p);
    const size_t smem = size_t(np2) * (sizeof(WKey) + sizeof(double)) + size_t(n) * sizeof(double);
    REACH_TRY(scratch.alloc(ctx, size_t(batch) * p * sizeof(int), false));
    DevBuf tmp;
    const int rn = r * n;
    REACH_TRY(tmp.alloc(ctx, size_t(batch) * rn * ld_in * sizeof(double)));
    if (smem > 48 * 1024)
        REACH_CUDA(ctx, cudaFuncSetAttribute(glgm06_reduce_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    glgm06_reduce_batch_kernel<<<batch, 256, smem, ctx->stream>>>(Gin, ld_in, p, n, r, np2, tmp.as<double>(), sel,
                                                                  scratch.as<int>());
    REACH_CUDA(ctx, cudaGetLastError());
    for (int z = 0; z < batch; ++z)
        REACH_CUDA(ctx, cudaMemcpy2DAsync(Gout + size_t(z) * out_stride, size_t(ld_out) * 8,
                                          tmp.as<double>() + size_t(z) * rn * ld_in, size_t(ld_in) * 8,
                                          size_t(n) * 8, rn, cudaMemcpyDeviceToDevice, ctx->stream));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REACH_OK;
}

}  // namespace

extern "C" int reach_glgm06_plan_upload(reach_glgm06_plan* p, const double* Phi, const double* c0, const double* G0,
                                        const double* cU, const double* GU) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!Phi || !c0) return set_error(ctx, REACH_EINVAL, "Phi and c0 must not be NULL");
    if (p->p0 > 0 && !G0) return set_error(ctx, REACH_EINVAL, "G0 is NULL but p0 > 0");
    if (!p->homog && p->pU > 0 && !GU) return set_error(ctx, REACH_EINVAL, "GU is NULL but pU > 0");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = p->n, Kp = p->Kp, nb = p->nb, r = p->r, nsplits = p->nsplits;
    for (size_t i = 0; i < size_t(n) * n; ++i)
        if (!std::isfinite(Phi[i])) return set_error(ctx, REACH_EINVAL, "Phi contains a non-finite entry");

    // ---- sizes of the shared chain ----
    p->pW.assign(size_t(p->nblk) + 1, 0);
    p->pWcap = 0;
    if (!p->homog) {
        int cur = p->pU;
        for (int64_t b = 0; b <= p->nblk; ++b) {
            p->pW[b] = cur;
            p->pWcap = std::max(p->pWcap, cur);
            const int nxt = p->pU + cur;
            cur = nxt > r * n ? r * n : nxt;
        }
        p->pWcap = std::max(p->pWcap, 1);
    }

    // ---- memory budget ----
    const int64_t nblk = p->nblk;
    size_t need = 0;
    if (p->homog) {
        need = size_t(p->nsteps) * p->m_pad * Kp * 8;
    } else {
        need = size_t(nblk) * p->m_pad * p->lda * 8                      // A_store
               + size_t(nblk) * p->tpb * p->m_pad * 16                   // psum, pmax
               + size_t(nblk) * nsplits * (size_t(p->p0r) * 4 + 4 + size_t(n) * 16)
               + size_t(nblk + 1) * p->pWcap * n * 8                     // GWall
               + size_t(nblk) * (size_t(p->pWcap) * 12 + size_t(p->pWcap + 1) * n * 8)
               + size_t(nblk) * nb * Kp * 8;
    }
    size_t free_b = 0, total_b = 0;
    REACH_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    free_b += ctx->cached_bytes;                      // blocks parked in the ctx cache are reusable
    if (need > size_t(0.9 * double(free_b)))
        return set_error(ctx, REACH_ENOMEM,
                         "the device-resident flowpipe needs %.1f GB but only %.1f GB of HBM are free; reduce NSTEPS "
                         "or the number of splits", need / 1e9, free_b / 1e9);
    p->bytes_stored = int64_t(need);

    // ---- Phi in the three layouts the contractions read ----
    REACH_TRY(p->phi_cm.alloc(ctx, size_t(n) * n * 8, false));
    REACH_TRY(p->PhiT.alloc(ctx, size_t(round_up(nb, 64)) * Kp * 8));          // 64-row tiles of gemm_store
    REACH_TRY(p->PhiRows.alloc(ctx, size_t(round_up(nb, 64)) * Kp * 8));
    REACH_CUDA(ctx, cudaMemcpyAsync(p->phi_cm.p, Phi, size_t(n) * n * 8, cudaMemcpyHostToDevice, st));
    // rows of PhiT = columns of Phi (contiguous in column-major storage)
    REACH_CUDA(ctx, cudaMemcpy2DAsync(p->PhiT.p, size_t(Kp) * 8, p->phi_cm.p, size_t(n) * 8, size_t(n) * 8, n,
                                      cudaMemcpyDeviceToDevice, st));
    {
        dim3 grid((n + 31) / 32, (n + 31) / 32), blk(32, 8);
        transpose_kernel<<<grid, blk, 0, st>>>(p->phi_cm.as<double>(), n, p->PhiRows.as<double>(), Kp, n);
        REACH_CUDA(ctx, cudaGetLastError());
    }

    // ---- initial sets: rows = generators then centre, per split ----
    REACH_TRY(p->Xown.alloc(ctx, size_t(p->m_pad) * Kp * 8));
    const int rows = p->rows;
    if (p->p0 > r * n) {
        // post.jl:26: Omega0 = reduce_order(Omega0, max_order, GIR05) -- on the device, per split
        DevBuf raw;
        REACH_TRY(raw.alloc(ctx, size_t(nsplits) * p->p0 * n * 8, false));
        REACH_CUDA(ctx, cudaMemcpyAsync(raw.p, G0, size_t(nsplits) * p->p0 * n * 8, cudaMemcpyHostToDevice, st));
        REACH_TRY(launch_reduce_batch(ctx, raw.as<double>(), n, p->p0, n, r, nsplits, p->Xown.as<double>(), Kp,
                                      (long long)rows * Kp, nullptr, p->scratch_pos));
    } else if (p->p0 > 0) {
        for (int s = 0; s < nsplits; ++s)
            REACH_CUDA(ctx, cudaMemcpy2DAsync(p->Xown.as<double>() + size_t(s) * rows * Kp, size_t(Kp) * 8,
                                              G0 + size_t(s) * p->p0 * n, size_t(n) * 8, size_t(n) * 8, p->p0,
                                              cudaMemcpyHostToDevice, st));
    }
    // centres: row p0r of every split
    REACH_CUDA(ctx, cudaMemcpy2DAsync(p->Xown.as<double>() + size_t(p->p0r) * Kp, size_t(rows) * Kp * 8, c0,
                                      size_t(n) * 8, size_t(n) * 8, nsplits, cudaMemcpyHostToDevice, st));
    REACH_CUDA(ctx, cudaStreamSynchronize(st));

    if (p->homog) {
        REACH_TRY(p->H.alloc(ctx, size_t(p->nsteps) * p->m_pad * Kp * 8));
    } else {
        const int urows = round_up(p->pU + 1, 64);
        REACH_TRY(p->XU.alloc(ctx, size_t(urows) * Kp * 8));
        REACH_TRY(p->PU.alloc(ctx, size_t(urows) * Kp * 8));
        if (p->pU > 0)
            REACH_CUDA(ctx, cudaMemcpy2DAsync(p->XU.p, size_t(Kp) * 8, GU, size_t(n) * 8, size_t(n) * 8, p->pU,
                                              cudaMemcpyHostToDevice, st));
        if (cU)
            REACH_CUDA(ctx, cudaMemcpyAsync(p->XU.as<double>() + size_t(p->pU) * Kp, cU, size_t(n) * 8,
                                            cudaMemcpyHostToDevice, st));
        REACH_TRY(p->Pstack.alloc(ctx, (size_t(std::max<int64_t>(1, nblk)) * nb + 64) * Kp * 8));
        REACH_TRY(p->GWall.alloc(ctx, size_t(nblk + 1) * p->pWcap * n * 8));
        REACH_TRY(p->slotB.alloc(ctx, size_t(nblk + 1) * p->pWcap * 4, false));
        {
            const size_t cnt = size_t(nblk + 1) * p->pWcap;
            glgm06_iota_kernel<<<unsigned((cnt + 255) / 256), 256, 0, st>>>(p->slotB.as<int>(), cnt);
            REACH_CUDA(ctx, cudaGetLastError());
        }
        REACH_TRY(p->cWall.alloc(ctx, size_t(nblk + 1) * n * 8));
        REACH_TRY(p->wB.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * p->pWcap * 8));
        REACH_TRY(p->permB.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * p->pWcap * 4));
        REACH_TRY(p->bpref.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * (p->pWcap + 1) * n * 8));
        REACH_TRY(p->pWdev.alloc(ctx, size_t(nblk + 1) * 4, false));
        REACH_CUDA(ctx, cudaMemcpyAsync(p->pWdev.p, p->pW.data(), size_t(nblk + 1) * 4, cudaMemcpyHostToDevice, st));
        REACH_TRY(p->A_store.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * p->m_pad * p->lda * 8, false));
        REACH_TRY(p->psum.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * p->tpb * p->m_pad * 8, false));
        REACH_TRY(p->pmax.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * p->tpb * p->m_pad * 8, false));
        REACH_TRY(p->pos_own.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * nsplits * std::max(1, p->p0r) * 4, false));
        REACH_TRY(p->tb.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * nsplits * 4, false));
        REACH_TRY(p->diag.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * nsplits * n * 8, false));
        REACH_TRY(p->centre.alloc(ctx, size_t(std::max<int64_t>(1, nblk)) * nsplits * n * 8, false));
        REACH_TRY(p->scratch_pos.alloc(ctx, size_t(std::max(p->pU, 1)) * 4, false));
        // W_0 = U: generators and centre
        if (p->pU > 0)
            REACH_CUDA(ctx, cudaMemcpyAsync(p->GWall.p, GU, size_t(p->pU) * n * 8, cudaMemcpyHostToDevice, st));
        if (cU) REACH_CUDA(ctx, cudaMemcpyAsync(p->cWall.p, cU, size_t(n) * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }
    p->uploaded = true;
    p->ran = false;
    return REACH_OK;
}

extern "C" int reach_glgm06_plan_run(reach_glgm06_plan* p) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->uploaded) return set_error(ctx, REACH_ESTATE, "plan_run before plan_upload");
    NvtxRange nvtx_fn("reach_glgm06_plan_run (powers, shared chain, batched product, selection)");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = p->n, Kp = p->Kp, nb = p->nb, r = p->r, nsplits = p->nsplits;
    const int64_t nblk = p->nblk;
    p->launches = 0;
    REACH_CUDA(ctx, cudaEventRecord(p->ev0, st));

    auto gemm_store = [&](const double* X, int xrows, const double* Y, int yrows, double* out, int ldc, int m_valid,
                          int n_valid) -> int {
        GemmOperands g{X, Y, Kp, Kp, Kp, round_up(xrows, 64) / 64, round_up(yrows, 64) / 64};
        StoreEpi epi{out, ldc, m_valid, n_valid};
        REACH_CUDA(ctx, (launch_tn64(g, epi, st)));
        ++p->launches;
        return REACH_OK;
    };

    if (p->homog) {
        // Z_k = Phi Z_{k-1}: generators and centres of every split are rows of X (reach_homog.jl:65-71)
        double* H = p->H.as<double>();
        const size_t slab = size_t(p->m_pad) * Kp;
        REACH_CUDA(ctx, cudaMemcpyAsync(H, p->Xown.p, slab * 8, cudaMemcpyDeviceToDevice, st));
        for (int64_t k = 1; k < p->nsteps; ++k)
            REACH_TRY(gemm_store(H + (k - 1) * slab, p->m_pad, p->PhiRows.as<double>(), n, H + k * slab, Kp, p->M, n));
        REACH_CUDA(ctx, cudaEventRecord(p->ev1, st));
        p->ran = true;
        return REACH_OK;
    }

    // ---- (1) shared chain ----
    double* Pst = p->Pstack.as<double>();
    const size_t pblk = size_t(nb) * Kp;
    if (nblk > 0)   // P_0 = Phi (rows of Phi)
        REACH_CUDA(ctx, cudaMemcpyAsync(Pst, p->PhiRows.p, pblk * 8, cudaMemcpyDeviceToDevice, st));
    bool overlap = false;
    // small n: the whole chain runs in one persistent CTA with P_k and Phi in shared memory
    bool fused = n <= 64 && p->pU + 1 <= 256 && p->pWcap <= 2048 && nblk > 0 && !p->force_stepwise;
    size_t chain_smem = 0;
    if (fused) {
        int np2 = 32;
        while (np2 < p->pU) np2 <<= 1;
        const size_t ldp = n + 1;
        chain_smem = 8 * (3 * n * ldp + 2 * size_t(p->pU + 1) * ldp + 3 * size_t(p->pWcap) + np2 + n + GLGM06_NSEG * size_t(n)) +
                     16 * size_t(np2) + 4 * (5 * size_t(p->pWcap) + 2 * size_t(p->pU) + 2) + 64;
        if (chain_smem > ctx->smem_optin) fused = false;
    }
    if (fused) {
        ChainParams c{};
        c.n = n; c.ldk = Kp; c.nb = nb; c.pU = p->pU; c.r = r; c.pWcap = p->pWcap; c.nblk = nblk;
        c.pa_max = std::max(p->p0r, p->pU);
        c.XU = p->XU.as<double>();
        c.Pstack = Pst;
        c.GWall = p->GWall.as<double>();
        c.cWall = p->cWall.as<double>();
        c.pW = p->pWdev.as<int>();
        c.wB = p->wB.as<double>();
        c.permB = p->permB.as<int>();
        c.bpref = p->bpref.as<double>();
        c.slotB = p->slotB.as<int>();
        if (getenv("REACH_GLGM06_DEBUG")) {
            if (!p->dbg.p) REACH_TRY(p->dbg.alloc(ctx, 64));
            c.dbg = p->dbg.as<long long>();
        }
        // Launch order: the powers kernel (1 CTA, waits on nothing, ~8 us per step) FIRST on the second stream, then the
        // chain CTA on the ctx stream, then the batched GEMM behind the powers in stream order.  The chain consumes the
        // powers as they are published (progress counter).  Because its producer is already in flight (or, in a
        // serialising environment -- ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING=1 -- already finished) when the chain
        // starts, the spin can never wait on a kernel that has not been launched.  The GEMM's CTAs only arrive once
        // the powers kernel is done, long after the chain CTA (200 KB of shared memory) has taken its SM.
        // Every kernel that is launched while the chain CTA may be spinning on the progress counter must be loaded
        // beforehand: with CUDA's lazy module loading the first launch of a kernel can wait for running kernels to
        // finish, which would deadlock against the spinning chain.
        {
            cudaFuncAttributes fa;
            REACH_CUDA(ctx, cudaFuncGetAttributes(&fa, glgm06_powers_kernel));
            REACH_CUDA(ctx, cudaFuncGetAttributes(&fa, glgm06_select_kernel));
            REACH_CUDA(ctx, cudaFuncGetAttributes(&fa, glgm06_select_warp_kernel<1>));
            REACH_CUDA(ctx, cudaFuncGetAttributes(&fa, glgm06_select_warp_kernel<2>));
            REACH_CUDA(ctx, cudaFuncGetAttributes(&fa, glgm06_select_warp_kernel<4>));
            REACH_CUDA(ctx, cudaFuncGetAttributes(&fa, glgm06_select_warp_kernel<8>));
            const GemmOperands none{nullptr, nullptr, Kp, Kp, Kp, 0, 0};
            if (p->TN <= 64) REACH_CUDA(ctx, (launch_small_tn(p->TN, none, GenEpi{}, p->stream2)));
            else REACH_CUDA(ctx, (launch_tn128(none, GenEpi{}, p->stream2)));
        }
        if (!p->progress.p) REACH_TRY(p->progress.alloc(ctx, sizeof(int)));
        REACH_CUDA(ctx, cudaMemsetAsync(p->progress.p, 0, sizeof(int), st));
        c.progress = p->progress.as<int>();
        REACH_CUDA(ctx, cudaEventRecord(p->ev_powers, st));          // "state reset, previous run finished"
        const size_t psmem = 8 * 3 * size_t(n) * (n + 1);
        REACH_CUDA(ctx, cudaFuncSetAttribute(glgm06_powers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(psmem)));
        REACH_CUDA(ctx, cudaFuncSetAttribute(glgm06_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(chain_smem)));
        REACH_CUDA(ctx, cudaStreamWaitEvent(p->stream2, p->ev_powers, 0));
        glgm06_powers_kernel<<<1, 1024, psmem, p->stream2>>>(p->PhiRows.as<double>(), p->PhiT.as<double>(), n, Kp, nb, nblk, Pst,
                                                             p->progress.as<int>());
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
        glgm06_chain_kernel<<<1, 1024, chain_smem, st>>>(c);
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
        overlap = true;
    } else {
    const int urows = round_up(p->pU + 1, 64);
    const size_t gw_stride = size_t(p->pWcap) * n;
    for (int64_t b = 0; b < nblk; ++b) {
        const int pWb = p->pW[b];
        double* GWb = p->GWall.as<double>() + b * gw_stride;
        double* wBb = p->wB.as<double>() + b * p->pWcap;
        int* permBb = p->permB.as<int>() + b * p->pWcap;
        double* bprefb = p->bpref.as<double>() + b * size_t(p->pWcap + 1) * n;
        {
            const int np2 = next_pow2(std::max(1, pWb));
            const size_t smem = size_t(np2) * sizeof(WKey);
            if (smem > 48 * 1024)
                REACH_CUDA(ctx, cudaFuncSetAttribute(glgm06_shared_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            glgm06_shared_prep_kernel<<<1, 512, smem, st>>>(GWb, n, pWb, n, np2, wBb, permBb, bprefb);
            REACH_CUDA(ctx, cudaGetLastError());
            ++p->launches;
        }
        if (b + 1 < nblk) {
            // P_b [G_U | c_U]
            REACH_TRY(gemm_store(p->XU.as<double>(), urows, Pst + b * pblk, n, p->PU.as<double>(), Kp, p->pU + 1, n));
            const int np2 = next_pow2(std::max(1, p->pU));
            const size_t smem = size_t(np2) * (sizeof(WKey) + sizeof(double)) + size_t(n) * sizeof(double);
            if (smem > 48 * 1024)
                REACH_CUDA(ctx, cudaFuncSetAttribute(glgm06_w_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            glgm06_w_reduce_kernel<<<1, 512, smem, st>>>(p->PU.as<double>(), Kp, p->pU, GWb, n, pWb,
                                                         p->cWall.as<double>() + b * n, wBb, permBb, bprefb, n, r, np2,
                                                         GWb + gw_stride, p->cWall.as<double>() + (b + 1) * n,
                                                         p->scratch_pos.as<int>());
            REACH_CUDA(ctx, cudaGetLastError());
            ++p->launches;
        }
        if (b + 1 < nblk)   // P_{b+1} = P_b Phi : Y[m][l] = Phi(l, m) = rows of PhiT (mul!(cache, P, Phi), reach_inhomog.jl:38)
            REACH_TRY(gemm_store(Pst + b * pblk, n, p->PhiT.as<double>(), n, Pst + (b + 1) * pblk, Kp, n, n));
    }

    }

    // ---- (2) batched part: all splits x all steps ----
    cudaStream_t gst = st;                              // stream of the batched GEMM
    if (overlap) gst = p->stream2;                      // after the powers kernel, in stream order
    if (nblk > 0) {
        GenEpi epi{};
        epi.out = p->A_store.as<double>();
        epi.blk_stride = (long long)p->m_pad * p->lda;
        epi.ldc = p->lda;
        epi.n = n;
        epi.nb = nb;
        epi.m_valid = p->M;
        epi.psum = p->psum.as<double>();
        epi.pmax = p->pmax.as<double>();
        epi.m_pad = p->m_pad;
        // grid.x is limited to 2^31-1 tiles; chunk the step axis to stay far below it
        const int64_t tiles_m = p->m_pad / p->TM;
        const int64_t max_tiles = 1 << 30;
        const int64_t blk_per_launch = std::max<int64_t>(1, max_tiles / (tiles_m * p->tpb));
        for (int64_t b0 = 0; b0 < nblk; b0 += blk_per_launch) {
            const int64_t nb_here = std::min(blk_per_launch, nblk - b0);
            GemmOperands g{p->Xown.as<double>(), Pst + b0 * pblk, Kp, Kp, Kp, int(tiles_m), int(nb_here * p->tpb)};
            GenEpi e2 = epi;
            e2.out = epi.out + b0 * epi.blk_stride;
            e2.psum = epi.psum + b0 * p->tpb * p->m_pad;
            e2.pmax = epi.pmax + b0 * p->tpb * p->m_pad;
            if (p->TN <= 64) REACH_CUDA(ctx, (launch_small_tn(p->TN, g, e2, gst)));
            else REACH_CUDA(ctx, (launch_tn128(g, e2, gst)));
            ++p->launches;
        }
        if (overlap) {                                  // the selection needs both the GEMM and the chain
            REACH_CUDA(ctx, cudaEventRecord(p->ev_gemm, gst));
            REACH_CUDA(ctx, cudaStreamWaitEvent(st, p->ev_gemm, 0));
        }
        SelectParams s{};
        s.A = p->A_store.as<double>();
        s.a_blk_stride = epi.blk_stride;
        s.lda = p->lda;
        s.psum = p->psum.as<double>();
        s.pmax = p->pmax.as<double>();
        s.tpb = p->tpb;
        s.m_pad = p->m_pad;
        s.n = n;
        s.p0 = p->p0r;
        s.nsplits = nsplits;
        s.nblk = int(nblk);
        s.r = r;
        s.pW = p->pWdev.as<int>();
        s.wB = p->wB.as<double>();
        s.bpref = p->bpref.as<double>();
        s.cW = p->cWall.as<double>();
        s.pWcap = p->pWcap;
        s.pos_own = p->pos_own.as<int>();
        s.tb = p->tb.as<int>();
        s.diag = p->diag.as<double>();
        s.centre = p->centre.as<double>();
        const int np2 = next_pow2(std::max(1, p->p0r));
        const size_t smem = size_t(np2) * (sizeof(WKey) + sizeof(double));
        if (smem > 48 * 1024)
            REACH_CUDA(ctx, cudaFuncSetAttribute(glgm06_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        for (int64_t b0 = 0; b0 < nblk; b0 += 32768) {
            const int nbh = int(std::min<int64_t>(32768, nblk - b0));
            SelectParams sb = s;
            sb.A += b0 * s.a_blk_stride;
            sb.psum += b0 * s.tpb * s.m_pad;
            sb.pmax += b0 * s.tpb * s.m_pad;
            sb.pW += b0;
            sb.wB += b0 * s.pWcap;
            sb.bpref += b0 * size_t(s.pWcap + 1) * n;
            sb.cW += b0 * n;
            sb.pos_own += b0 * size_t(nsplits) * s.p0;
            sb.tb += b0 * nsplits;
            sb.diag += b0 * size_t(nsplits) * n;
            sb.centre += b0 * size_t(nsplits) * n;
            // up to 256 own generators: one warp per (split, step), keys in registers; beyond: one CTA, keys in shared memory
            const unsigned wgrid = unsigned((size_t(nsplits) * nbh + 3) / 4);
            if (np2 <= 32) glgm06_select_warp_kernel<1><<<wgrid, 128, 0, st>>>(sb, nbh);
            else if (np2 == 64) glgm06_select_warp_kernel<2><<<wgrid, 128, 0, st>>>(sb, nbh);
            else if (np2 == 128) glgm06_select_warp_kernel<4><<<wgrid, 128, 0, st>>>(sb, nbh);
            else if (np2 == 256) glgm06_select_warp_kernel<8><<<wgrid, 128, 0, st>>>(sb, nbh);
            else glgm06_select_kernel<<<dim3(nsplits, nbh), 128, smem, st>>>(sb, np2);
            REACH_CUDA(ctx, cudaGetLastError());
            ++p->launches;
        }
    }
    REACH_CUDA(ctx, cudaEventRecord(p->ev1, st));
    p->ran = true;
    return REACH_OK;
}

extern "C" int reach_glgm06_plan_sync(reach_glgm06_plan* p) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->ran) return set_error(ctx, REACH_ESTATE, "plan_sync before plan_run");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    REACH_CUDA(ctx, cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    p->last_ms = ms;
    if (p->dbg.p && getenv("REACH_GLGM06_DEBUG")) {
        long long h[6];
        REACH_CUDA(ctx, cudaMemcpy(h, p->dbg.p, sizeof(h), cudaMemcpyDeviceToHost));
        const double k = double(std::max<int64_t>(1, p->nsteps));
        fprintf(stderr, "[glgm06 chain dbg] cycles per step: publish+prefix %.0f, P[G_U|c_U] %.0f, wait+load P %.0f, weights+sort %.0f, "
                        "cut+copy W %.0f, diag+order %.0f\n", h[0] / k, h[1] / k, h[2] / k, h[3] / k, h[4] / k, h[5] / k);
    }
    return REACH_OK;
}

namespace {
int check_index(reach_glgm06_plan* p, int split, int64_t step) {
    if (!p->ran) return set_error(p->ctx, REACH_ESTATE, "the plan has not been run");
    if (split < 0 || split >= p->nsplits)
        return set_error(p->ctx, REACH_EINVAL, "split %d outside [0, %d)", split, p->nsplits);
    if (step < 0 || step >= p->nsteps)
        return set_error(p->ctx, REACH_EINVAL, "step %lld outside [0, %lld)", (long long)step, (long long)p->nsteps);
    return REACH_OK;
}
int ngens_of(const reach_glgm06_plan* p, int64_t step) {
    if (p->homog || step == 0) return p->p0r;
    const int tot = p->p0r + p->pW[step - 1];
    return tot > p->r * p->n ? p->r * p->n : tot;
}
}  // namespace

extern "C" int reach_glgm06_plan_ngens(reach_glgm06_plan* p, int split, int64_t step, int* out) {
    if (!p || !out) return REACH_EINVAL;
    REACH_TRY(check_index(p, split, step));
    *out = ngens_of(p, step);
    return REACH_OK;
}

extern "C" int reach_glgm06_plan_fetch(reach_glgm06_plan* p, int split, int64_t step, double* c_host, double* G_host,
                                       int32_t* sel) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    REACH_TRY(check_index(p, split, step));
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = p->n, Kp = p->Kp;
    const int ng = ngens_of(p, step);
    if (p->homog || step == 0) {
        const double* Z = (p->homog ? p->H.as<double>() + size_t(step) * p->m_pad * Kp : p->Xown.as<double>()) +
                          size_t(split) * p->rows * Kp;
        if (G_host && ng > 0)
            REACH_CUDA(ctx, cudaMemcpy2DAsync(G_host, size_t(n) * 8, Z, size_t(Kp) * 8, size_t(n) * 8, ng,
                                              cudaMemcpyDeviceToHost, st));
        if (c_host)
            REACH_CUDA(ctx, cudaMemcpyAsync(c_host, Z + size_t(p->p0r) * Kp, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
        if (sel) for (int j = 0; j < ng; ++j) sel[j] = j;
        return REACH_OK;
    }
    const int64_t b = step - 1;
    const size_t o = size_t(b) * p->nsplits + split;
    if (c_host)
        REACH_CUDA(ctx, cudaMemcpyAsync(c_host, p->centre.as<double>() + o * n, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
    int tb = 0;
    REACH_CUDA(ctx, cudaMemcpyAsync(&tb, p->tb.as<int>() + o, sizeof(int), cudaMemcpyDeviceToHost, st));
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    if (G_host || sel) {
        DevBuf Gd, seld, occ;
        REACH_TRY(Gd.alloc(ctx, size_t(std::max(1, ng)) * n * 8));
        REACH_TRY(seld.alloc(ctx, size_t(std::max(1, ng)) * 4));
        REACH_TRY(occ.alloc(ctx, size_t(std::max(1, ng)) * 4));
        glgm06_materialize_kernel<<<1, 256, 0, st>>>(
            p->A_store.as<double>() + size_t(b) * p->m_pad * p->lda + size_t(split) * p->rows * p->lda, p->lda, p->p0r,
            p->pos_own.as<int>() + o * p->p0r, tb, p->GWall.as<double>(), p->slotB.as<int>() + size_t(b) * p->pWcap, n, p->pW[b],
            p->permB.as<int>() + size_t(b) * p->pWcap, p->diag.as<double>() + o * n, n, p->r, Gd.as<double>(),
            seld.as<int>(), occ.as<int>());
        REACH_CUDA(ctx, cudaGetLastError());
        if (G_host) REACH_CUDA(ctx, cudaMemcpyAsync(G_host, Gd.p, size_t(ng) * n * 8, cudaMemcpyDeviceToHost, st));
        if (sel) REACH_CUDA(ctx, cudaMemcpyAsync(sel, seld.p, size_t(ng) * 4, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return REACH_OK;
}

extern "C" int reach_glgm06_plan_support(reach_glgm06_plan* p, const double* d, double* out_host) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->ran) return set_error(ctx, REACH_ESTATE, "the plan has not been run");
    if (!d || !out_host) return set_error(ctx, REACH_EINVAL, "d and out_host must not be NULL");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = p->n, Kp = p->Kp, nsplits = p->nsplits;
    DevBuf dd, out;
    REACH_TRY(dd.alloc(ctx, size_t(n) * 8, false));
    REACH_TRY(out.alloc(ctx, size_t(p->nsteps) * nsplits * 8, false));
    REACH_CUDA(ctx, cudaMemcpyAsync(dd.p, d, size_t(n) * 8, cudaMemcpyHostToDevice, st));
    const int th = 256;
    if (p->homog) {
        // zonotope (step k, split s) = rows of H at (k*m_pad + s*rows)
        for (int64_t k = 0; k < p->nsteps; ++k) {
            zono_support_dense_kernel<<<(nsplits * 32 + th - 1) / th, th, 0, st>>>(
                p->H.as<double>() + size_t(k) * p->m_pad * Kp, (long long)p->rows * Kp, Kp, p->p0r, n, dd.as<double>(),
                nsplits, out.as<double>() + size_t(k) * nsplits);
            REACH_CUDA(ctx, cudaGetLastError());
        }
    } else {
        zono_support_dense_kernel<<<(nsplits * 32 + th - 1) / th, th, 0, st>>>(
            p->Xown.as<double>(), (long long)p->rows * Kp, Kp, p->p0r, n, dd.as<double>(), nsplits, out.as<double>());
        REACH_CUDA(ctx, cudaGetLastError());
        if (p->nblk > 0) {
            DevBuf sfx;
            REACH_TRY(sfx.alloc(ctx, size_t(p->nblk) * (p->pWcap + 1) * 8, false));
            const size_t smem = size_t(p->pWcap) * 8;
            if (smem > 48 * 1024)
                REACH_CUDA(ctx, cudaFuncSetAttribute(glgm06_shared_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            glgm06_shared_dots_kernel<<<unsigned(p->nblk), 256, smem, st>>>(
                p->GWall.as<double>(), p->slotB.as<int>(), n, p->pWdev.as<int>(), p->permB.as<int>(), p->pWcap,
                dd.as<double>(), n, sfx.as<double>());
            REACH_CUDA(ctx, cudaGetLastError());
            SupportParams s{};
            s.A = p->A_store.as<double>();
            s.a_blk_stride = (long long)p->m_pad * p->lda;
            s.lda = p->lda;
            s.n = n;
            s.p0 = p->p0r;
            s.nsplits = nsplits;
            s.nblk = int(p->nblk);
            s.pos_own = p->pos_own.as<int>();
            s.tb = p->tb.as<int>();
            s.diag = p->diag.as<double>();
            s.centre = p->centre.as<double>();
            s.sfx = sfx.as<double>();
            s.pWcap = p->pWcap;
            s.d = dd.as<double>();
            s.out = out.as<double>();
            const int64_t warps = int64_t(nsplits) * p->nblk;
            glgm06_support_kernel<<<unsigned((warps * 32 + th - 1) / th), th, 0, st>>>(s);
            REACH_CUDA(ctx, cudaGetLastError());
            REACH_CUDA(ctx, cudaStreamSynchronize(st));
        }
    }
    REACH_CUDA(ctx, cudaMemcpyAsync(out_host, out.p, size_t(p->nsteps) * nsplits * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    return REACH_OK;
}

extern "C" int reach_glgm06_plan_stats(reach_glgm06_plan* p, double* ms, int64_t* launches, int64_t* bytes_stored) {
    if (!p) return REACH_EINVAL;
    if (ms) *ms = p->last_ms;
    if (launches) *launches = p->launches;
    if (bytes_stored) *bytes_stored = p->bytes_stored;
    return REACH_OK;
}

extern "C" int reach_glgm06(reach_ctx* ctx, int n, int64_t nsteps, int nsplits, int max_order, const double* Phi,
                            const double* c0, const double* G0, int p0, const double* cU, const double* GU, int pU,
                            double* c_host, double* G_host, int gcap, int32_t* ngens_host) {
    reach_glgm06_plan* p = nullptr;
    REACH_TRY(reach_glgm06_plan_create(ctx, n, nsteps, nsplits, max_order, p0, pU, &p));
    int rc = reach_glgm06_plan_upload(p, Phi, c0, G0, cU, GU);
    if (rc == REACH_OK) rc = reach_glgm06_plan_run(p);
    if (rc == REACH_OK) rc = reach_glgm06_plan_sync(p);
    for (int64_t k = 0; rc == REACH_OK && k < nsteps; ++k) {
        const int ng = ngens_of(p, k);
        if (G_host && ng > gcap) {
            rc = set_error(ctx, REACH_EINVAL, "gcap = %d is smaller than the %d generators of step %lld", gcap, ng,
                           (long long)k);
            break;
        }
        for (int s = 0; rc == REACH_OK && s < nsplits; ++s) {
            const size_t z = size_t(k) * nsplits + s;
            if (ngens_host) ngens_host[z] = ng;
            if (c_host || G_host)
                rc = reach_glgm06_plan_fetch(p, s, k, c_host ? c_host + z * n : nullptr,
                                             G_host ? G_host + z * size_t(gcap) * n : nullptr, nullptr);
        }
    }
    reach_glgm06_plan_destroy(p);
    return rc;
}

extern "C" int reach_reduce_order_gir05(reach_ctx* ctx, int n, int p, int max_order, const double* G, double* Gred,
                                        int* pred, int32_t* sel) {
    if (!ctx) return REACH_EINVAL;
    if (n < 1 || p < 0 || !pred) return set_error(ctx, REACH_EINVAL, "bad arguments");
    if (max_order < 1)
        return set_error(ctx, REACH_EINVAL, "the target order should be at least 1, but it is %d", max_order);
    if (p > kMaxSort) return set_error(ctx, REACH_EUNSUPPORTED, "more than %d generators", kMaxSort);
    if (max_order * n >= p) {                           // returned unchanged
        if (Gred && G && Gred != G) std::memcpy(Gred, G, size_t(n) * p * 8);
        if (sel) for (int j = 0; j < p; ++j) sel[j] = j;
        *pred = p;
        return REACH_OK;
    }
    if (!G || !Gred) return set_error(ctx, REACH_EINVAL, "G and Gred must not be NULL");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rn = max_order * n;
    DevBuf Gd, Go, seld, scratch;
    REACH_TRY(Gd.alloc(ctx, size_t(p) * n * 8, false));
    REACH_TRY(Go.alloc(ctx, size_t(rn) * n * 8));
    REACH_TRY(seld.alloc(ctx, size_t(rn) * 4));
    REACH_CUDA(ctx, cudaMemcpyAsync(Gd.p, G, size_t(p) * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    REACH_TRY(launch_reduce_batch(ctx, Gd.as<double>(), n, p, n, max_order, 1, Go.as<double>(), n, (long long)rn * n,
                                  seld.as<int>(), scratch));
    REACH_CUDA(ctx, cudaMemcpyAsync(Gred, Go.p, size_t(rn) * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (sel) REACH_CUDA(ctx, cudaMemcpyAsync(sel, seld.p, size_t(rn) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *pred = rn;
    return REACH_OK;
}
