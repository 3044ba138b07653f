// nt_dgemm.cuh -- FP64 tensor-core ("DMMA") GEMM for sm_100a with a pluggable fused epilogue.
//
//     C[m][n] = sum_k X[m][k] * Y[n][k]            (both operands K-contiguous: "NT" form)
//
// This is the one contraction shape the whole hot path reduces to:
//   LGG09   X = direction block L_k (one row per direction chain), Y = stacked rows
//           [(Phi')^s ; E (Phi')^s], s = 1..S  -> C = L_{k+s} rows and the mapped support rows
//           (mul!(r_{i+1}, Phi', r_i) at reach_homog.jl:58 / reach_inhomog.jl:60, all
//           directions and S time steps at once)
//   GLGM06  X = generators (one row per generator, all splits), Y = rows of the power P_k
//           -> C = P_k g  (mul!(view(G,:,1:n2), M, G1) at Continuous/setops.jl:21)
//   powers  P_{k+1} = P_k Phi (GLGM06/reach_inhomog.jl:38) and the (Phi')^s stack.
//
// FP64 has no tcgen05 kind on Blackwell; the FP64 tensor path is mma.sync.m8n8k4.f64 (SASS
// DMMA.8x8x4).  Tiles are staged global->shared with a multi-stage cp.async ring; shared rows
// are padded to BK+4 doubles so that the 8-byte fragment loads of a half-warp (4 rows x 4 k)
// hit 16 distinct 8-byte banks.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace reach {

constexpr int BK = 16;         // k-tile, doubles (128 B per row)
constexpr int SK = BK + 4;     // padded shared-memory row, doubles

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4],
// C[l/4][2*(l%4)] and C[l/4][2*(l%4)+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct GemmOperands {
    const double* X;   // [tiles_m*TM][ldx], rows beyond the logical M are zero
    const double* Y;   // [tiles_n*TN][ldy], rows beyond the logical N are zero
    int ldx, ldy;
    int K;             // multiple of BK (operands are zero-padded in k)
    int tiles_m, tiles_n;
};

template <int TM_, int TN_>
struct TileCoord {
    static constexpr int TM = TM_, TN = TN_;
    int tm, tn;        // tile indices
    int wm0, wn0;      // warp origin inside the tile
    int g4, t4;        // lane/4, lane%4
    int tid, nthr;
};

template <int WM, int WN, int MT, int NT, int STAGES>
struct GemmShape {
    static constexpr int TM = WM * MT * 8, TN = WN * NT * 8, NTHR = WM * WN * 32;
    static constexpr size_t SMEM = size_t(STAGES) * (TM + TN) * SK * sizeof(double);
};

// The accumulator of lane (g4,t4): acc[mt][nt][e] is C[wm0 + 8*mt + g4][wn0 + 8*nt + 2*t4 + e].
template <int WM, int WN, int MT, int NT, int STAGES, class Epi>
__global__ void __launch_bounds__(WM* WN * 32, 1) nt_dgemm_kernel(const GemmOperands g, const Epi epi) {
    using Shape = GemmShape<WM, WN, MT, NT, STAGES>;
    constexpr int TM = Shape::TM, TN = Shape::TN, NTHR = Shape::NTHR;
    extern __shared__ __align__(16) double smem[];
    double* sX = smem;                                  // [STAGES][TM][SK]
    double* sY = smem + size_t(STAGES) * TM * SK;       // [STAGES][TN][SK]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp % WM, wn = warp / WM;
    const int g4 = lane >> 2, t4 = lane & 3;
    const int tm = blockIdx.x % g.tiles_m;              // direction tiles vary fastest: CTAs in
    const int tn = blockIdx.x / g.tiles_m;              // flight share the same Y row block in L2

    const double* gX = g.X + size_t(tm) * TM * g.ldx;
    const double* gY = g.Y + size_t(tn) * TN * g.ldy;
    const int KT = g.K / BK;

    auto load_tile = [&](int stage, int kt) {
        double* dx = sX + size_t(stage) * TM * SK;
        double* dy = sY + size_t(stage) * TN * SK;
        const double* sx = gX + size_t(kt) * BK;
        const double* sy = gY + size_t(kt) * BK;
#pragma unroll
        for (int c = tid; c < TM * 8; c += NTHR) {      // 8 x 16-byte chunks per row
            int r = c >> 3, ch = c & 7;
            cp_async16(dx + r * SK + ch * 2, sx + size_t(r) * g.ldx + ch * 2);
        }
#pragma unroll
        for (int c = tid; c < TN * 8; c += NTHR) {
            int r = c >> 3, ch = c & 7;
            cp_async16(dy + r * SK + ch * 2, sy + size_t(r) * g.ldy + ch * 2);
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();                                 // tile kt landed; stage (kt-1)%STAGES is free
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_tile(nk % STAGES, nk);
            cp_async_commit();
        }
        const int stage = kt % STAGES;
        const double* x = sX + (size_t(stage) * TM + wm * MT * 8 + g4) * SK + t4;
        const double* y = sY + (size_t(stage) * TN + wn * NT * 8 + g4) * SK + t4;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = x[i * 8 * SK + k4 * 4];
#pragma unroll
            for (int j = 0; j < NT; ++j) b[j] = y[j * 8 * SK + k4 * 4];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                     // shared memory is reusable by the epilogue

    TileCoord<TM, TN> tc;
    tc.tm = tm; tc.tn = tn;
    tc.wm0 = wm * MT * 8; tc.wn0 = wn * NT * 8;
    tc.g4 = g4; tc.t4 = t4; tc.tid = tid; tc.nthr = NTHR;
    epi.template run<MT, NT, WN>(acc, tc, smem);
}

// ---------------------------------------------------------------------------------------------
// Epilogue: plain store.  C[m][n] -> out[m*ldc + n] for m < m_valid, n < n_valid.
// ---------------------------------------------------------------------------------------------
struct StoreEpi {
    double* out;
    int ldc;
    int m_valid, n_valid;

    template <int MT, int NT, int WN, class TC>
    __device__ __forceinline__ void run(double (&acc)[MT][NT][2], const TC& tc, double*) const {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            int m = tc.tm * TC::TM + tc.wm0 + i * 8 + tc.g4;
            if (m >= m_valid) continue;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                int n = tc.tn * TC::TN + tc.wn0 + j * 8 + 2 * tc.t4;
                double* o = out + size_t(m) * ldc + n;
                if (n + 1 < n_valid && (ldc & 1) == 0) {
                    *reinterpret_cast<double2*>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
                } else {
                    if (n < n_valid) o[0] = acc[i][j][0];
                    if (n + 1 < n_valid) o[1] = acc[i][j][1];
                }
            }
        }
    }
};

template <int WM, int WN, int MT, int NT, int STAGES, class Epi>
inline cudaError_t launch_nt_dgemm(const GemmOperands& g, const Epi& epi, cudaStream_t stream) {
    using Shape = GemmShape<WM, WN, MT, NT, STAGES>;
    auto kern = nt_dgemm_kernel<WM, WN, MT, NT, STAGES, Epi>;
    // the opt-in shared-memory size is a per-device function attribute
    static thread_local int configured_dev = -1;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Shape::SMEM));
        if (e != cudaSuccess) return e;
        configured_dev = dev;
    }
    if (g.tiles_m <= 0 || g.tiles_n <= 0) return cudaSuccess;
    kern<<<g.tiles_m * g.tiles_n, Shape::NTHR, Shape::SMEM, stream>>>(g, epi);
    return cudaGetLastError();
}

// out[r][c] = in[c][r] for an n x n block (in: ld_in, out: ld_out); used once per run to get
// the row-major copy of Phi that seeds the power stack.
static __global__ void transpose_kernel(const double* __restrict__ in, int ld_in, double* __restrict__ out,
                                 int ld_out, int n) {
    __shared__ double tile[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    for (int dy = 0; dy < 32; dy += 8)
        if (x < n && y + dy < n) tile[threadIdx.y + dy][threadIdx.x] = in[size_t(y + dy) * ld_in + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y = blockIdx.x * 32 + threadIdx.y;
    for (int dy = 0; dy < 32; dy += 8)
        if (x < n && y + dy < n) out[size_t(y + dy) * ld_out + x] = tile[threadIdx.x][threadIdx.y + dy];
}

}  // namespace reach
