// expm.cu -- device-side discretisation matrices: Phi = exp(A d), Phi1(A, d) = sum d^(i+1)/(i+1)! A^i and
// Phi2(A, d) = sum d^(i+2)/(i+2)! A^i, the three matrix functions src/Discretization/Exponentiation.jl computes
// (`exp` :168-171; Phi1 by the 2n x 2n block exponential :263-291; Phi2 by the 3n x 3n block exponential :397-427)
// and the Forward / NoBloating models consume (ForwardModule.jl:86-164, NoBloatingModule.jl:72-105).
//
// The reference exponentiates the block matrix [[A d, d I, 0], [0, 0, d I], [0, 0, 0]] with the dense `exp`
// (scaling and squaring on 27 n^3-blocks).  The blocks of every power of that matrix are functions of A alone, so the
// same scaling and squaring runs here on the three n x n blocks directly, with this library's FP64 tensor-core GEMM:
//   h = d / 2^s with ||A||_inf h <= 1/2;
//   Phi2(h) by Horner on its Taylor series (14 terms: truncation < 1e-21 relative), Phi1(h) = h I + A Phi2(h),
//   Phi(h) = I + A Phi1(h);
//   s doublings:  Phi2(2h) = Phi2 + h Phi1 + Phi2 Phi,   Phi1(2h) = Phi1 + Phi1 Phi,   Phi(2h) = Phi Phi
// (exactly the block rows of the squared block matrix).  Everything commutes (functions of A), so every product is
// written X * Y' with Y' = A' (the caller's column-major A read as row-major) or Phi' (kept beside Phi).
#include "reach_common.cuh"
#include "nt_dgemm.cuh"

#include <cmath>
#include <vector>

using namespace reach;

namespace {

// out = alpha * acc + beta1 * B1 + beta2 * B2 + gamma * I   (row-major, ld = ldc)
struct AffineEpi {
    double* out;
    const double* B1;
    const double* B2;
    int ldc, n_valid;
    double alpha, beta1, beta2, gamma;

    template <int MT, int NT, int WN, class TC>
    __device__ __forceinline__ void run(double (&acc)[MT][NT][2], const TC& tc, double*) const {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int m = tc.tm * TC::TM + tc.wm0 + i * 8 + tc.g4;
            if (m >= n_valid) continue;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = tc.tn * TC::TN + tc.wn0 + j * 8 + 2 * tc.t4 + e;
                    if (n >= n_valid) continue;
                    const size_t at = size_t(m) * ldc + n;
                    double v = alpha * acc[i][j][e];
                    if (B1) v = fma(beta1, B1[at], v);
                    if (B2) v = fma(beta2, B2[at], v);
                    if (m == n) v += gamma;
                    out[at] = v;
                }
            }
        }
    }
};

}  // namespace

extern "C" int reach_discretize_phi(reach_ctx* ctx, int n, const double* A, double delta, double* Phi, double* Phi1,
                                    double* Phi2, int32_t* squarings) {
    if (!ctx) return REACH_EINVAL;
    if (n < 1 || !A) return set_error(ctx, REACH_EINVAL, "n must be >= 1 and A must not be NULL");
    if (!(delta > 0) || !std::isfinite(delta)) return set_error(ctx, REACH_EINVAL, "the step size must be positive, got %g", delta);
    double norm = 0.0;                                      // opnorm(A, Inf): max row sum of the column-major input
    {
        std::vector<double> rows(n, 0.0);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                const double v = A[size_t(j) * n + i];
                if (!std::isfinite(v)) return set_error(ctx, REACH_EINVAL, "A contains a non-finite entry");
                rows[i] += std::fabs(v);
            }
        for (double r : rows) norm = std::max(norm, r);
    }
    int s = 0;
    while (norm * delta / std::ldexp(1.0, s) > 0.5 && s < 60) ++s;
    if (squarings) *squarings = s;
    const double h = delta / std::ldexp(1.0, s);

    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool big = n >= 512;
    const int T = big ? 128 : 64;
    const int Kp = round_up(n, BK), R = round_up(n, T);
    const size_t bytes = size_t(R) * Kp * sizeof(double);
    DevBuf At, P, Pt, F1, F2, W[2];
    REACH_TRY(At.alloc(ctx, bytes));
    REACH_TRY(P.alloc(ctx, bytes));
    REACH_TRY(Pt.alloc(ctx, bytes));
    REACH_TRY(F1.alloc(ctx, bytes));
    REACH_TRY(F2.alloc(ctx, bytes));
    REACH_TRY(W[0].alloc(ctx, bytes));
    REACH_TRY(W[1].alloc(ctx, bytes));
    // A column-major == A' row-major
    REACH_CUDA(ctx, cudaMemcpy2DAsync(At.p, size_t(Kp) * 8, A, size_t(n) * 8, size_t(n) * 8, n, cudaMemcpyHostToDevice, st));

    auto gemm = [&](const double* X, const double* Yt, double* out, double alpha, const double* B1, double beta1,
                    const double* B2, double beta2, double gamma) -> int {
        GemmOperands g{X, Yt, Kp, Kp, Kp, R / T, R / T};
        AffineEpi epi{out, B1, B2, Kp, n, alpha, beta1, beta2, gamma};
        if (big) REACH_CUDA(ctx, (launch_nt_dgemm_ws<2, 4, 8, 4, 5>(g, epi, st)));
        else REACH_CUDA(ctx, (launch_nt_dgemm<2, 2, 4, 4, 4>(g, epi, st)));
        return REACH_OK;
    };
    // Horner: S_m = I,  S_i = I + (h / (i + 3)) S_{i+1} A   =>  Phi2(h) = (h^2 / 2) S_0
    constexpr int kTerms = 14;
    int cur = 0;
    {   // S_m = I
        std::vector<double> eye(size_t(R) * Kp, 0.0);
        for (int i = 0; i < n; ++i) eye[size_t(i) * Kp + i] = 1.0;
        REACH_CUDA(ctx, cudaMemcpyAsync(W[cur].p, eye.data(), bytes, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }
    for (int i = kTerms - 1; i >= 0; --i) {
        const bool last = i == 0;
        // last step folds the h^2/2 factor: Phi2 = (h^2/2) I + (h^2/2)(h/3) S_1 A
        const double c = h / double(i + 3);
        if (!last) REACH_TRY(gemm(W[cur].as<double>(), At.as<double>(), W[1 - cur].as<double>(), c, nullptr, 0, nullptr, 0, 1.0));
        else REACH_TRY(gemm(W[cur].as<double>(), At.as<double>(), F2.as<double>(), 0.5 * h * h * c, nullptr, 0, nullptr, 0, 0.5 * h * h));
        cur ^= 1;
    }
    REACH_TRY(gemm(F2.as<double>(), At.as<double>(), F1.as<double>(), 1.0, nullptr, 0, nullptr, 0, h));     // Phi1 = h I + Phi2 A
    REACH_TRY(gemm(F1.as<double>(), At.as<double>(), P.as<double>(), 1.0, nullptr, 0, nullptr, 0, 1.0));    // Phi  = I + Phi1 A
    auto transpose = [&](const double* in, double* out) -> int {
        dim3 grid((n + 31) / 32, (n + 31) / 32), blk(32, 8);
        transpose_kernel<<<grid, blk, 0, st>>>(in, Kp, out, Kp, n);
        REACH_CUDA(ctx, cudaGetLastError());
        return REACH_OK;
    };
    double hh = h;
    for (int q = 0; q < s; ++q) {
        REACH_TRY(transpose(P.as<double>(), Pt.as<double>()));
        // Phi2(2h) = Phi2 + h Phi1 + Phi2 Phi ; Phi1(2h) = Phi1 + Phi1 Phi ; Phi(2h) = Phi Phi
        REACH_TRY(gemm(F2.as<double>(), Pt.as<double>(), W[0].as<double>(), 1.0, F2.as<double>(), 1.0, F1.as<double>(), hh, 0.0));
        REACH_TRY(gemm(F1.as<double>(), Pt.as<double>(), W[1].as<double>(), 1.0, F1.as<double>(), 1.0, nullptr, 0, 0.0));
        REACH_CUDA(ctx, cudaMemcpyAsync(F2.p, W[0].p, bytes, cudaMemcpyDeviceToDevice, st));
        REACH_CUDA(ctx, cudaMemcpyAsync(F1.p, W[1].p, bytes, cudaMemcpyDeviceToDevice, st));
        REACH_TRY(gemm(P.as<double>(), Pt.as<double>(), W[0].as<double>(), 1.0, nullptr, 0, nullptr, 0, 0.0));
        REACH_CUDA(ctx, cudaMemcpyAsync(P.p, W[0].p, bytes, cudaMemcpyDeviceToDevice, st));
        hh *= 2.0;
    }
    // results back in the caller's column-major layout: transpose on the device, then a pitched copy
    auto fetch = [&](const double* src, double* host) -> int {
        if (!host) return REACH_OK;
        REACH_TRY(transpose(src, W[1].as<double>()));
        REACH_CUDA(ctx, cudaMemcpy2DAsync(host, size_t(n) * 8, W[1].p, size_t(Kp) * 8, size_t(n) * 8, n, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
        return REACH_OK;
    };
    REACH_TRY(fetch(P.as<double>(), Phi));
    REACH_TRY(fetch(F1.as<double>(), Phi1));
    REACH_TRY(fetch(F2.as<double>(), Phi2));
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    return REACH_OK;
}
