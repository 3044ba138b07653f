// oz_gemm.cuh -- FP64-accurate GEMM on the INT8 tcgen05 tensor pipe (error-free slicing, "Ozaki scheme I").
//
//     C[m][n] = sum_k X[m][k] * Y[n][k]            (the same NT contraction as nt_dgemm.cuh)
//
// B200's FP64 tensor path (DMMA) tops out at ~36 TFLOP/s; its INT8 tcgen05 path is >100x faster.  Each FP64
// operand row is scaled by a power of two and cut into T signed 7-bit digits,
//     x = 2^e * sum_s q_s 2^(-6-7s),   q_s in [-64, 64]  (exact: every step is a scaling and an integer subtraction)
// so that  x_i . y_j = 2^(e_i+e_j-12) * sum_d 2^(-7d) * ( sum_{s+u=d} q_s . p_u ).  The inner integer dot products
// are exact in the INT32 accumulators (|.| <= (d+1) K 2^12 < 2^31), all (s,u) pairs of one level d accumulate
// into the SAME tensor-memory accumulator, and only levels d < T are kept: T(T+1)/2 INT8 MMAs per k-step and a
// truncation error below (T+1) K 2^(-7T) relative to rowmax(X) * rowmax(Y) (T = 8: ~2^-42 worst case, ~2^-50 typical:
// on par with DGEMM's own rounding).  The result is independent of the summation order, hence deterministic.
//
// Kernel shape (one CTA per 128 x 64 output tile, 6 warps):
//   warp 4  producer: per 32-byte k-block two cp.async.bulk copies (all T digit planes of the X tile, of the Y
//           tile) into a 4-stage shared-memory ring; the digit planes are stored in global memory already in the
//           canonical no-swizzle K-major UMMA layout (8-row x 16-byte core matrices), so a plain bulk copy lands
//           the exact shared-memory image the MMA descriptors expect.
//   warp 5  one thread issues tcgen05.mma.kind::i8 (M=128, N=64, K=32): T accumulators of 64 columns = the whole
//           tensor memory (512 columns) for T = 8; tcgen05.commit frees the ring slot / signals the epilogue.
//   warps 0-3 epilogue: tcgen05.ld 32x32b (thread = output row), Horner over the levels in FP64, scale, then the
//           fused support-function epilogue (or a plain store).
#pragma once

#include "nt_dgemm.cuh"

namespace reach {

constexpr int OZ_TM = 128, OZ_TN = 64, OZ_KB = 32, OZ_STAGES = 4, OZ_THREADS = 192;
constexpr int OZ_MAX_T = 8;

struct OzOperands {
    const int8_t* Xq;     // [nkb][tiles_m_total][T][16][2][8][16]
    const int8_t* Yq;     // [nkb][tiles_n_total][T][ 8][2][8][16]
    const int* eX;        // [tiles_m_total*128] row exponents
    const int* eY;        // [tiles_n_total*64]
    int nkb;              // k-blocks of 32
    int tiles_m, tiles_m_total;
    int tile_n0, tiles_n, tiles_n_total;   // this launch covers Y tiles [tile_n0, tile_n0 + tiles_n)
};

__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory matrix descriptor, SWIZZLE_NONE, K-major: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
    // version 1 at [46,48)
    return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) |
           (uint64_t(1) << 46);
}

__device__ __forceinline__ double oz_pow2(int e) {   // 2^e for e in the normal range
    return __hiloint2double((1023 + e) << 20, 0);
}
__device__ __forceinline__ double oz_scale(double v, int e) {   // v * 2^e, any e
    if (e >= -1000 && e <= 1000) return v * oz_pow2(e);
    return scalbn(v, e);
}

// ---- plain store epilogue (probe / tests): C -> out[m*ldc + n] -------------------------------------------
struct OzStoreEpi {
    double* out;
    int ldc, m_valid, n_valid;
    struct State {};
    __device__ __forceinline__ void prepare(unsigned char*, int, int) const {}
    __device__ __forceinline__ void begin(State&) const {}
    __device__ __forceinline__ void accept(State&, const unsigned char*, int m, int tn, int j, double v) const {
        const int n = tn * OZ_TN + j;
        if (m < m_valid && n < n_valid) out[size_t(m) * ldc + n] = v;
    }
    __device__ __forceinline__ void finish(State&, int, int, int) const {}
    static constexpr size_t SMEM = 0;
};

template <int T, class Epi>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_gemm_kernel(const OzOperands g, const Epi epi) {
    constexpr uint32_t XBYTES = T * OZ_TM * OZ_KB, YBYTES = T * OZ_TN * OZ_KB, STAGE = XBYTES + YBYTES;
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    unsigned char* ring = oz_smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(oz_smem + size_t(OZ_STAGES) * STAGE);
    uint64_t* empty = full + OZ_STAGES;
    uint64_t* tmem_full = empty + OZ_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    double* sYscale = reinterpret_cast<double*>(tmem_full + 2);       // [64] 2^eY of the tile's columns
    unsigned char* epi_smem = reinterpret_cast<unsigned char*>(sYscale + OZ_TN);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tm = blockIdx.x % g.tiles_m;
    const int tn = g.tile_n0 + blockIdx.x / g.tiles_m;

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {   // the MMA warp owns the tensor-memory allocation (all 512 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== producer =====
        if (lane == 0) {
            const int8_t* gx = g.Xq + size_t(tm) * XBYTES;
            const int8_t* gy = g.Yq + size_t(tn) * YBYTES;
            const size_t xstep = size_t(g.tiles_m_total) * XBYTES, ystep = size_t(g.tiles_n_total) * YBYTES;
            for (int kb = 0; kb < g.nkb; ++kb) {
                const int s = kb % OZ_STAGES;
                const uint32_t ph = (kb / OZ_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], STAGE);
                unsigned char* dst = ring + size_t(s) * STAGE;
                bulk_copy_g2s(dst, gx + size_t(kb) * xstep, XBYTES, &full[s]);
                bulk_copy_g2s(dst + XBYTES, gy + size_t(kb) * ystep, YBYTES, &full[s]);
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D = S32 (2<<4), A = B = signed 8-bit (1<<7, 1<<10), K-major both, N>>3 at 17, M>>4 at 24
            constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(OZ_TN >> 3) << 17) | (uint32_t(OZ_TM >> 4) << 24);
            for (int kb = 0; kb < g.nkb; ++kb) {
                const int s = kb % OZ_STAGES;
                const uint32_t ph = (kb / OZ_STAGES) & 1;
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t xs = smem_u32(ring + size_t(s) * STAGE), ys = xs + XBYTES;
#pragma unroll
                for (int d = 0; d < T; ++d) {
#pragma unroll
                    for (int sx = 0; sx <= d; ++sx) {
                        const int sy = d - sx;
                        // one digit plane of a tile: [rows/8][2 k-chunks][8 rows][16 B]: LBO = 128 B (next k-chunk), SBO = 256 B (next 8 rows)
                        const uint64_t da = oz_smem_desc(xs + sx * (OZ_TM * OZ_KB), 128, 256);
                        const uint64_t db = oz_smem_desc(ys + sy * (OZ_TN * OZ_KB), 128, 256);
                        const uint32_t acc = (kb > 0 || sx > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                            " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base + uint32_t(d * OZ_TN)),
                            "l"(da), "l"(db), "r"(idesc), "r"(acc)
                            : "memory");
                    }
                }
                // frees the ring slot once the MMAs above have read it (implies fence::before_thread_sync)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(tmem_full)) : "memory");
        }
    } else {
        // ===== epilogue warps 0..3: thread = output row =====
        const int ml = warp * 32 + lane;
        const int m = tm * OZ_TM + ml;
        if (tid < OZ_TN) sYscale[tid] = oz_pow2(max(-1000, min(1000, g.eY[size_t(tn) * OZ_TN + tid])));
        epi.prepare(epi_smem, tn, tid);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int ex = g.eX[size_t(g.tiles_m_total > 0 ? tm : 0) * OZ_TM + ml] - 12;
        typename Epi::State st;
        epi.begin(st);
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16);
#pragma unroll 1
        for (int cc = 0; cc < OZ_TN / 8; ++cc) {
            uint32_t a[T][8];
#pragma unroll
            for (int d = 0; d < T; ++d)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(a[d][0]), "=r"(a[d][1]), "=r"(a[d][2]), "=r"(a[d][3]), "=r"(a[d][4]), "=r"(a[d][5]),
                               "=r"(a[d][6]), "=r"(a[d][7])
                             : "r"(taddr + uint32_t(d * OZ_TN + cc * 8)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                double v = double(int(a[T - 1][jj]));
#pragma unroll
                for (int d = T - 2; d >= 0; --d) v = fma(v, 0.0078125, double(int(a[d][jj])));   // * 2^-7
                const int j = cc * 8 + jj;
                v = oz_scale(v * sYscale[j], ex);
                epi.accept(st, epi_smem, m, tn, j, v);
            }
        }
        epi.finish(st, tm, tn, ml);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int T, class Epi>
struct OzShape {
    static constexpr size_t STAGE = size_t(T) * (OZ_TM + OZ_TN) * OZ_KB;
    static constexpr size_t SMEM = OZ_STAGES * STAGE + (2 * OZ_STAGES + 2) * sizeof(uint64_t) + OZ_TN * sizeof(double) + Epi::SMEM;
};

template <int T, class Epi>
inline cudaError_t launch_oz_gemm_t(const OzOperands& g, const Epi& epi, cudaStream_t stream) {
    auto kern = oz_gemm_kernel<T, Epi>;
    constexpr size_t smem = OzShape<T, Epi>::SMEM;
    static thread_local int configured_dev = -1;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        configured_dev = dev;
    }
    if (g.tiles_m <= 0 || g.tiles_n <= 0) return cudaSuccess;
    kern<<<g.tiles_m * g.tiles_n, OZ_THREADS, smem, stream>>>(g, epi);
    return cudaGetLastError();
}

template <class Epi>
inline cudaError_t launch_oz_gemm(int T, const OzOperands& g, const Epi& epi, cudaStream_t stream) {
    switch (T) {
        case 4: return launch_oz_gemm_t<4>(g, epi, stream);
        case 5: return launch_oz_gemm_t<5>(g, epi, stream);
        case 6: return launch_oz_gemm_t<6>(g, epi, stream);
        case 7: return launch_oz_gemm_t<7>(g, epi, stream);
        case 8: return launch_oz_gemm_t<8>(g, epi, stream);
    }
    return cudaErrorInvalidValue;
}

// ---- slicing: FP64 rows -> T digit planes in the UMMA core-matrix layout --------------------------------
// One CTA per 8-row group.  Phase 1: row maxima -> exponents (2^e > max|x|).  Phase 2: thread = (row, 16-wide k chunk):
// 16 values -> T x 16 digits, written as T 16-byte stores (8 consecutive rows = one 128-byte core matrix).
//   Q[((kb * tiles_total + tile) * T + s) * (TR*32) + ((g * 2 + c) * 8 + r) * 16 + b],   tile = row / TR, g = (row % TR) / 8
template <int TR>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ X, int ldx, int rows_valid, int k_valid,
                                                       int T, int nkb, int tiles_total, int8_t* __restrict__ Q,
                                                       int* __restrict__ eout) {
    __shared__ int s_e[8];
    const int row0 = blockIdx.x * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {   // phase 1: warp w scans row row0 + w
        const int row = row0 + warp;
        double mx = 0.0;
        if (row < rows_valid) {
            const double* x = X + size_t(row) * ldx;
            for (int k = lane; k < k_valid; k += 32) mx = fmax(mx, fabs(x[k]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) {
            int e = 0;
            if (mx > 0.0 && mx < 1.0e308) e = ilogb(mx) + 1;           // 2^(e-1) <= mx < 2^e
            s_e[warp] = e;
            eout[row] = e;
        }
    }
    __syncthreads();
    const int tile = row0 / TR, gidx = (row0 % TR) / 8;
    const int nchunks = nkb * 2;
    for (int item = threadIdx.x; item < 8 * nchunks; item += blockDim.x) {
        const int r = item & 7, chunk = item >> 3;
        const int row = row0 + r;
        const int kb = chunk >> 1, c = chunk & 1;
        const int k0 = chunk * 16;
        double v[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int k = k0 + b;
            v[b] = (row < rows_valid && k < k_valid) ? X[size_t(row) * ldx + k] : 0.0;
        }
        const int e = s_e[r];
#pragma unroll
        for (int b = 0; b < 16; ++b) v[b] = oz_scale(v[b], 6 - e);          // |y| < 64
        int8_t* q = Q + ((size_t(kb) * tiles_total + tile) * T) * (TR * 32) + ((gidx * 2 + c) * 8 + r) * 16;
        for (int s = 0; s < T; ++s) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const double d = rint(v[b]);
                v[b] = (v[b] - d) * 128.0;                                   // exact
                w[b >> 2] |= (uint32_t(int(d)) & 0xFFu) << (8 * (b & 3));
            }
            *reinterpret_cast<uint4*>(q + size_t(s) * (TR * 32)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

}  // namespace reach
