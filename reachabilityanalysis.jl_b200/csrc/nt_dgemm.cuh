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

// NAMED_BAR == 0: the whole CTA computes (sync = __syncthreads); otherwise only `nthr` consumer
// threads reach the epilogue and synchronise on that named barrier (the producer warp is gone).
template <int TM_, int TN_, int NAMED_BAR = 0>
struct TileCoord {
    static constexpr int TM = TM_, TN = TN_;
    int tm, tn;        // tile indices
    int wm0, wn0;      // warp origin inside the tile
    int g4, t4;        // lane/4, lane%4
    int tid, nthr;     // thread id among / number of the threads that run the epilogue
    __device__ __forceinline__ void sync() const {
        if constexpr (NAMED_BAR == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"n"(NAMED_BAR), "r"(nthr) : "memory");
    }
};

template <int WM, int WN, int MT, int NT, int STAGES>
struct GemmShape {
    static constexpr int TM = WM * MT * 8, TN = WN * NT * 8, NTHR = WM * WN * 32;
    static constexpr size_t SMEM = size_t(STAGES) * (TM + TN) * SK * sizeof(double);
};

// The accumulator of lane (g4,t4): acc[mt][nt][e] is C[wm0 + 8*mt + g4][wn0 + 8*nt + 2*t4 + e].
// MINB: CTAs per SM the register allocation must allow (short-K launches are latency-bound per CTA and want 4).
template <int WM, int WN, int MT, int NT, int STAGES, int MINB = 1, class Epi>
__global__ void __launch_bounds__(WM* WN * 32, MINB) nt_dgemm_kernel(const GemmOperands g, const Epi epi) {
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
// v2 main loop: warp-specialised.  A producer warpgroup streams k-tiles into a STAGES-deep ring
// with cp.async and signals each stage's "full" mbarrier when its copies have landed; WM*WN consumer
// warps wait on "full", issue DMMAs and arrive on the stage's "empty" mbarrier; setmaxnreg moves
// the producers' registers to the consumers.  The consumer warps never issue a copy and never meet at
// a CTA-wide barrier, so the FP64 tensor pipe of a sub-partition only idles when its own data is late.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 27)) __trap();     // a protocol error must fail, not hang the GPU
    } while (!ok);
}
// The same wait for warps that idle through a whole main loop (epilogue warps waiting for the accumulators): sleep
// between polls so that hundreds of spinning threads do not burn issue slots and power next to the tensor pipe.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns = 512) {
    uint32_t ok = 0, spins = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok) {
            __nanosleep(ns);
            if (++spins > (1u << 24)) __trap();
        }
    } while (!ok);
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int WM, int WN, int MT, int NT, int STAGES>
struct GemmShapeWS {
    static constexpr int TM = WM * MT * 8, TN = WN * NT * 8;
    // consumers + one producer warpgroup (register reallocation with setmaxnreg works per 4 warps)
    static constexpr int NCONS = WM * WN * 32, NTHR = NCONS + 128;
    static constexpr size_t TILE_BYTES = size_t(STAGES) * (TM + TN) * SK * sizeof(double);
    static constexpr size_t SMEM = TILE_BYTES + 2 * STAGES * sizeof(uint64_t);
};

template <int WM, int WN, int MT, int NT, int STAGES, class Epi>
__global__ void __launch_bounds__(WM* WN * 32 + 128, 1) nt_dgemm_ws_kernel(const GemmOperands g, const Epi epi) {
    using Shape = GemmShapeWS<WM, WN, MT, NT, STAGES>;
    constexpr int TM = Shape::TM, TN = Shape::TN, NCONS = Shape::NCONS;
    extern __shared__ __align__(128) double smem[];
    double* sX = smem;                                  // [STAGES][TM][SK]
    double* sY = smem + size_t(STAGES) * TM * SK;       // [STAGES][TN][SK]
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(smem) + Shape::TILE_BYTES);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tm = blockIdx.x % g.tiles_m;
    const int tn = blockIdx.x / g.tiles_m;
    const int KT = g.K / BK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 128);                   // one deferred arrive per producer thread
            mbar_init(&empty[s], WM * WN);              // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp >= WM * WN) {
        // ===== producer warpgroup: hands its registers to the DMMA warps and streams the k-tiles =====
        // 16-byte cp.async (LDGSTS) into the padded rows; each producer thread then lets the stage's
        // "full" mbarrier count the completion of its own copies (cp.async.mbarrier.arrive.noinc).
        // (One cp.async.bulk per 128-byte row was measured 3x slower: the copy engine issues one
        // small bulk copy per ~57 cycles.)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        const int ptid = tid - NCONS;                   // 0..127
        const double* gX = g.X + size_t(tm) * TM * g.ldx;
        const double* gY = g.Y + size_t(tn) * TN * g.ldy;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);               // slot free (passes immediately on the first lap)
            double* dx = sX + size_t(s) * TM * SK;
            double* dy = sY + size_t(s) * TN * SK;
            const double* sx = gX + size_t(kt) * BK;
            const double* sy = gY + size_t(kt) * BK;
#pragma unroll
            for (int c = ptid; c < TM * 8; c += 128) {
                const int r = c >> 3, ch = c & 7;
                cp_async16(dx + r * SK + ch * 2, sx + size_t(r) * g.ldx + ch * 2);
            }
#pragma unroll
            for (int c = ptid; c < TN * 8; c += 128) {
                const int r = c >> 3, ch = c & 7;
                cp_async16(dy + r * SK + ch * 2, sy + size_t(r) * g.ldy + ch * 2);
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
        }
        return;
    }

    // ===== consumer warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int wm = warp % WM, wn = warp / WM;
    const int g4 = lane >> 2, t4 = lane & 3;
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const double* x = sX + (size_t(s) * TM + wm * MT * 8 + g4) * SK + t4;
        const double* y = sY + (size_t(s) * TN + wn * NT * 8 + g4) * SK + t4;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);          // this warp is done reading stage s
    }

    TileCoord<TM, TN, 1> tc;
    tc.tm = tm; tc.tn = tn;
    tc.wm0 = wm * MT * 8; tc.wn0 = wn * NT * 8;
    tc.g4 = g4; tc.t4 = t4; tc.tid = tid; tc.nthr = NCONS;
    tc.sync();                                          // every consumer is past the main loop: smem is free
    epi.template run<MT, NT, WN>(acc, tc, smem);
}

template <int WM, int WN, int MT, int NT, int STAGES, class Epi>
inline cudaError_t launch_nt_dgemm_ws(const GemmOperands& g, const Epi& epi, cudaStream_t stream) {
    using Shape = GemmShapeWS<WM, WN, MT, NT, STAGES>;
    auto kern = nt_dgemm_ws_kernel<WM, WN, MT, NT, STAGES, Epi>;
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

template <int WM, int WN, int MT, int NT, int STAGES, int MINB = 1, class Epi>
inline cudaError_t launch_nt_dgemm(const GemmOperands& g, const Epi& epi, cudaStream_t stream) {
    using Shape = GemmShape<WM, WN, MT, NT, STAGES>;
    auto kern = nt_dgemm_kernel<WM, WN, MT, NT, STAGES, MINB, Epi>;
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
