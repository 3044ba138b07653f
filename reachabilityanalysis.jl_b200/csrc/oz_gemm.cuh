// oz_gemm.cuh -- FP64-accurate GEMM on the INT8 tcgen05 tensor pipe (error-free slicing, "Ozaki scheme I").
//
//     C[m][n] = sum_k X[m][k] * Y[n][k]            (the same NT contraction as nt_dgemm.cuh)
//
// B200's FP64 tensor path (DMMA) tops out at ~36 TFLOP/s; its INT8 tcgen05 path is >100x faster.  Each FP64
// operand row is scaled by a power of two and cut into T signed 8-bit digits (balanced radix 256),
//     x = 2^(e-7) * sum_s q_s 2^(-8s),   q_s in [-128, 127]  (exact: scalings, floors and integer subtractions only)
// so that  x_i . y_j = 2^(e_i+e_j-14) * sum_d 2^(-8d) * ( sum_{s+u=d} q_s . p_u ).  The inner integer dot products
// are exact in the INT32 accumulators (|.| <= (d+1) K 2^14 < 2^31 for K <= 16384), all (s,u) pairs of one level d
// accumulate into the SAME tensor-memory accumulator, and only levels d < T are kept: T(T+1)/2 INT8 MMAs per k-step and a
// truncation error below T K 2^(-8T) relative to rowmax(X) * rowmax(Y) (T = 7: 56 bits, ~2^-42 worst case, ~2^-50
// typical: on par with DGEMM's own rounding).  The result is independent of the summation order, hence deterministic.
// (Rounds 1-2a used 7-bit digits in [-64, 64] from round-to-nearest: 8 planes = 36 plane pairs for the same 56 bits that
// 7 planes = 28 pairs give with full 8-bit digits.  The kernel runs at the chip's power-limited INT8 rate, so the
// plane-pair count is what sets its time.)
//
// Kernel shape (one CTA per 128 x 128 output tile, 18 warps):
//   warp 16 producer: per 32-byte k-block two cp.async.bulk copies (the digit planes of the X tile, of the Y
//           tile) into a 3-stage shared-memory ring; the digit planes are stored in global memory already in the
//           canonical no-swizzle K-major UMMA layout (8-row x 16-byte core matrices), so a plain bulk copy lands
//           the exact shared-memory image the MMA descriptors expect.
//   warp 17 one thread issues tcgen05.mma.kind::i8 (M=128, N=128, K=32): 4 level accumulators of 128 columns = the
//           whole tensor memory (512 columns); T > 4 takes two passes over K (levels 0..3, then 4..T-1).
//           tcgen05.commit frees the ring slot / signals the epilogue.  (A first version with N=64 and all 8 levels
//           resident was bound by the shared-memory operand feed of the tensor pipe: 73 % of its peak at 49 % MMA
//           utilisation, ncu profiles/; N=128 halves the operand bytes per MAC.)
//   warps 0-15 epilogue: tcgen05.ld 32x32b (thread = output row x 32 columns), Horner over the levels in FP64 (exact),
//           scale, then the fused support-function epilogue (or a plain store).
#pragma once

#include "nt_dgemm.cuh"

namespace reach {

constexpr int OZ_TM = 128, OZ_TN = 128, OZ_KB = 32, OZ_STAGES = 3;
constexpr int OZ_EPI_WARPS = 16, OZ_THREADS = (OZ_EPI_WARPS + 2) * 32;
constexpr int OZ_CPT = OZ_TN / (OZ_EPI_WARPS / 4);    // output columns per epilogue thread (32)
constexpr int OZ_MAX_T = 8, OZ_G = 4;                  // OZ_G levels live in tensor memory at a time (4 x 128 columns)
constexpr int OZ_DEFAULT_T = 7;                        // 7 x 8 = 56 bits below the row maximum
constexpr int OZ_BITS = 8;                             // bits per digit; x-hat = x 2^(OZ_BITS-1-e) lies in [-127, 127]
constexpr int OZ_MAX_K = (1 << (31 - 2 * (OZ_BITS - 1))) / OZ_MAX_T * 1;   // INT32 level sums stay exact up to this K (16384)
constexpr int OZ_PLANE = OZ_TM * OZ_KB;                // bytes of one digit plane of a 128-row tile per k-block

struct OzOperands {
    const int8_t* Xq;     // [nkb][tiles_m_total][T][16][2][8][16]
    const int8_t* Yq;     // [nkb][tiles_n_total][T][16][2][8][16]
    const int* eX;        // [tiles_m_total*128] row exponents
    const int* eY;        // [tiles_n_total*128]
    int nkb;              // k-blocks of 32
    int tiles_m, tiles_m_total;
    int tile_n0, tiles_n, tiles_n_total;   // this launch covers Y tiles [tile_n0, tile_n0 + tiles_n)
    int no_pair = 0;      // 1: one level per MMA (N = 128) everywhere; A/B switch for the two-level N = 256 MMAs
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
// exact int32 -> double without the conversion pipe: 2^52 + 2^31 + x has x in its low mantissa bits
__device__ __forceinline__ double oz_i2d(uint32_t x) {
    return __hiloint2double(0x43300000, int(x ^ 0x80000000u)) - 4503601774854144.0;
}

// ---- plain store epilogue (probe / tests): C -> out[m*ldc + n] -------------------------------------------
struct OzStoreEpi {
    double* out;
    int ldc, m_valid, n_valid;
    struct State {};
    static constexpr size_t SMEM = 0;
    __device__ __forceinline__ void prepare(unsigned char*, int, int) const {}
    __device__ __forceinline__ void begin(State&, int) const {}
    // value of output (row m, column j of tile tn)
    __device__ __forceinline__ void accept(State&, const unsigned char*, int m, int tn, int j, double v) const {
        const int n = tn * OZ_TN + j;
        if (m < m_valid && n < n_valid) out[size_t(m) * ldc + n] = v;
    }
    __device__ __forceinline__ void accept4(State& st, const unsigned char* sm, int m, int tn, int j0, const double (&v)[4]) const {
#pragma unroll
        for (int e = 0; e < 4; ++e) accept(st, sm, m, tn, j0 + e, v[e]);
    }
    __device__ __forceinline__ void finish(State&, unsigned char*, int, int, int, int) const {}
    template <int NG, int BAR>
    __device__ __forceinline__ void finish_g(State&, unsigned char*, int, int, int, int) const {}
};

// ---- carry epilogue (LGG09 pass groups): Y = [(Phi')^S; (Phi')^2S; ...], blocks of tiles_per_block tiles ----
// value (chain m, row r of block j) -> out[(j * rows_per_slot + m) * ldl + r]: the direction rows of pass t + j + 1.
struct OzCarryEpi {
    double* out;
    int ldl, n, tiles_per_block, rows_per_slot;
    struct State {};
    static constexpr size_t SMEM = 0;
    __device__ __forceinline__ void prepare(unsigned char*, int, int) const {}
    __device__ __forceinline__ void begin(State&, int) const {}
    __device__ __forceinline__ void accept(State&, const unsigned char*, int m, int tn, int j, double v) const {
        const int blk = tn / tiles_per_block, col = (tn - blk * tiles_per_block) * OZ_TN + j;
        if (col < n) out[(size_t(blk) * rows_per_slot + m) * ldl + col] = v;
    }
    __device__ __forceinline__ void accept4(State& st, const unsigned char* sm, int m, int tn, int j0, const double (&v)[4]) const {
#pragma unroll
        for (int e = 0; e < 4; ++e) accept(st, sm, m, tn, j0 + e, v[e]);
    }
    __device__ __forceinline__ void finish(State&, unsigned char*, int, int, int, int) const {}
    template <int NG, int BAR>
    __device__ __forceinline__ void finish_g(State&, unsigned char*, int, int, int, int) const {}
};

// ---- fused LGG09 epilogue on the digit-plane GEMM (packed stack layout, see Lgg09PackedEpi) ---------------
// Stack rows are addressed absolutely (tile tn covers rows [128 tn, 128 tn + 128) of the stack; block b = row / nbe).
// A tile straddles at most one block boundary (nbe >= 128): segment 0 = rows of the tile's first block, segment 1 =
// rows of the next.  Thread (row m, 32 columns) accumulates lin_q / abs_q per segment; the four column groups of a
// row are summed in fixed order through shared memory (the operand ring, idle by then).
template <int NQ>
struct OzLgg09Epi {
    double* Lnext;              // [mu_pad][ldl]
    int ldl, n, nbe;
    int store_block;            // absolute block whose rows are L_{k0+S}; -1: none
    const double* Wl;           // [nq][nbe]
    const double* Wa;
    int nq;
    double* featpart;           // [tiles_n_total][2 segments][nq][2][mu_pad]
    int mu_pad;
    struct State { double lin[NQ][2], ab[NQ][2]; int bnd; };
    static constexpr size_t SMEM = size_t(OZ_TN) * (2 * NQ) * sizeof(double) + size_t(OZ_TN) * sizeof(int) +
                                   size_t(NQ) * 4 * OZ_TM * sizeof(double);      // staged weights, row info, reduction buffer

    __device__ __forceinline__ void prepare(unsigned char* sm, int tn, int tid) const {
        double* sw = reinterpret_cast<double*>(sm);                      // [128 rows][2*NQ]: wl_0, wa_0, wl_1, wa_1, ... of the row
        int* srib = reinterpret_cast<int*>(sw + size_t(2 * NQ) * OZ_TN); // row in block | store flag << 30
        if (tid < OZ_TN) {
            const int row = tn * OZ_TN + tid;
            const int blk = row / nbe, rib = row - blk * nbe;
            srib[tid] = rib | (blk == store_block ? (1 << 30) : 0);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                sw[tid * (2 * NQ) + 2 * q] = q < nq ? Wl[size_t(q) * nbe + rib] : 0.0;
                sw[tid * (2 * NQ) + 2 * q + 1] = q < nq ? Wa[size_t(q) * nbe + rib] : 0.0;
            }
        }
    }
    __device__ __forceinline__ void begin(State& st, int tn) const {
#pragma unroll
        for (int q = 0; q < NQ; ++q) st.lin[q][0] = st.lin[q][1] = st.ab[q][0] = st.ab[q][1] = 0.0;
        const int row0 = tn * OZ_TN;
        st.bnd = (row0 / nbe + 1) * nbe - row0;                          // tile-local row where the next block starts
    }
    __device__ __forceinline__ void accept(State& st, const unsigned char* sm, int m, int, int j, double v) const {
        const double2* sw = reinterpret_cast<const double2*>(sm) + j * NQ;           // (wl_q, wa_q) of row j: NQ 16-byte loads
        const int* srib = reinterpret_cast<const int*>(sm + size_t(2 * NQ) * OZ_TN * sizeof(double));
        const int info = srib[j], rib = info & 0x3FFFFFFF;
        if ((info >> 30) && rib < n) Lnext[size_t(m) * ldl + rib] = v;
        const double av = fabs(v);
        if (j < st.bnd) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double2 w = sw[q];
                st.lin[q][0] = fma(w.x, v, st.lin[q][0]);
                st.ab[q][0] = fma(w.y, av, st.ab[q][0]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double2 w = sw[q];
                st.lin[q][1] = fma(w.x, v, st.lin[q][1]);
                st.ab[q][1] = fma(w.y, av, st.ab[q][1]);
            }
        }
    }
    // four consecutive columns j0 .. j0 + 3 (j0 a multiple of 4): when they lie on one side of the block boundary -- always,
    // except in the one slice that straddles it -- all shared-memory loads are issued before the first FMA (the tensor pipe's
    // operand reads keep the shared memory busy: a load per dependent FMA pair costs hundreds of cycles each)
    __device__ __forceinline__ void accept4(State& st, const unsigned char* sm, int m, int tn, int j0, const double (&v)[4]) const {
        if (j0 < st.bnd && j0 + 3 >= st.bnd) {
#pragma unroll
            for (int e = 0; e < 4; ++e) accept(st, sm, m, tn, j0 + e, v[e]);
            return;
        }
        const double2* sw = reinterpret_cast<const double2*>(sm) + j0 * NQ;
        const int4 info = *reinterpret_cast<const int4*>(sm + size_t(2 * NQ) * OZ_TN * sizeof(double) + size_t(j0) * sizeof(int));
        const int inf[4] = {info.x, info.y, info.z, info.w};
        if ((inf[0] | inf[3]) >> 30) {                                   // rows of the block that is carried to the next pass
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int rib = inf[e] & 0x3FFFFFFF;
                if ((inf[e] >> 30) && rib < n) Lnext[size_t(m) * ldl + rib] = v[e];
            }
        }
        const bool seg0 = j0 < st.bnd;
#pragma unroll
        for (int e2 = 0; e2 < 4; e2 += 2) {                              // two columns at a time: 8 independent 16-byte loads
            double2 w[2][NQ];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int q = 0; q < NQ; ++q) w[e][q] = sw[(e2 + e) * NQ + q];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double x = v[e2 + e], av = fabs(x);
                if (seg0) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        st.lin[q][0] = fma(w[e][q].x, x, st.lin[q][0]);
                        st.ab[q][0] = fma(w[e][q].y, av, st.ab[q][0]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        st.lin[q][1] = fma(w[e][q].x, x, st.lin[q][1]);
                        st.ab[q][1] = fma(w[e][q].y, av, st.ab[q][1]);
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void finish(State& st, unsigned char* sm, int tm, int tn, int ml, int cgrp) const {
        finish_g<OZ_EPI_WARPS / 4, 1>(st, sm, tm, tn, ml, cgrp);
    }
    // NG threads per output row (column groups 0..NG-1, 128 NG threads in all) meet on named barrier BAR
    template <int NG, int BAR>
    __device__ __forceinline__ void finish_g(State& st, unsigned char* sm, int tm, int tn, int ml, int cgrp) const {
        // red[q][seg][part][128 rows] after the staged weights: the column groups of a row add their partials in
        // fixed order (group 0 stores, 1.. add), one named barrier between the phases -> deterministic
        double* red = reinterpret_cast<double*>(sm + size_t(OZ_TN) * (2 * NQ) * sizeof(double) + size_t(OZ_TN) * sizeof(int));
#pragma unroll
        for (int c = 0; c < NG; ++c) {
            if (cgrp == c) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int sg = 0; sg < 2; ++sg) {
                        double* rl = &red[((q * 2 + sg) * 2 + 0) * OZ_TM + ml];
                        double* ra = &red[((q * 2 + sg) * 2 + 1) * OZ_TM + ml];
                        if (c == 0) { *rl = st.lin[q][sg]; *ra = st.ab[q][sg]; }
                        else { *rl += st.lin[q][sg]; *ra += st.ab[q][sg]; }
                    }
            }
            asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NG * OZ_TM) : "memory");
        }
        const int t = cgrp * OZ_TM + ml;
        for (int idx = t; idx < nq * 4 * OZ_TM; idx += NG * OZ_TM) {
            const int r = idx % OZ_TM, part = (idx / OZ_TM) & 1, sg = (idx / (2 * OZ_TM)) & 1, q = idx / (4 * OZ_TM);
            featpart[((((size_t(tn) * 2 + sg) * nq + q) * 2) + part) * mu_pad + tm * OZ_TM + r] = red[((q * 2 + sg) * 2 + part) * OZ_TM + r];
        }
    }
};

// MMAs of one k-block of sweep SW (levels 4 SW .. min(4 SW + 4, TE) - 1) with everything but the ring addresses folded
// at compile time (one thread issues them: a run-time loop nest costs more issue cycles than the MMAs take).
// X plane sx meets the Y planes sy = d - sx of the sweep's levels d.  Two consecutive Y planes are 256 consecutive rows
// of the shared-memory image (a plane is 16 row groups of 256 B = the SBO), and the accumulators of levels d, d + 1 are
// 256 consecutive tensor-memory columns: ONE M = 128, N = 256 MMA does the plane pairs (sx, sy) and (sx, sy + 1) and reads
// the X plane once instead of twice.  sx = 0 touches every level of the sweep first: at the first k-block those MMAs
// overwrite, everything else accumulates.
template <int TE, int SW, bool PAIR>
__device__ __forceinline__ void oz_issue_kblock(uint32_t xs, uint32_t ys, uint32_t tmem_base, bool first_kb) {
    // instruction descriptor: D = S32 (2<<4), A = B = signed 8-bit (1<<7, 1<<10), K-major both, N>>3 at 17, M>>4 at 24
    constexpr uint32_t idesc1 = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(OZ_TN >> 3) << 17) | (uint32_t(OZ_TM >> 4) << 24);
    constexpr uint32_t idesc2 = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t((2 * OZ_TN) >> 3) << 17) | (uint32_t(OZ_TM >> 4) << 24);
    constexpr int d0 = SW * OZ_G, d1 = (d0 + OZ_G < TE ? d0 + OZ_G : TE) - 1;
#pragma unroll
    for (int sx = 0; sx <= d1; ++sx) {
        const int lo = d0 - sx > 0 ? d0 - sx : 0, hi = d1 - sx;
        // one digit plane of a tile: [rows/8][2 k-chunks][8 rows][16 B]: LBO = 128 B (next k-chunk), SBO = 256 B (next 8 rows)
        const uint64_t da = oz_smem_desc(xs + sx * OZ_PLANE, 128, 256);
        const uint32_t acc = (!first_kb || sx > 0) ? 1u : 0u;
#pragma unroll
        for (int j = 0; j < OZ_G / 2; ++j) {
            const int sy = lo + 2 * j;
            if (sy <= hi) {
                const uint64_t db = oz_smem_desc(ys + sy * OZ_PLANE, 128, 256);
                const bool two = PAIR && sy + 1 <= hi;
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                    " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base + uint32_t((sx + sy - d0) * OZ_TN)),
                    "l"(da), "l"(db), "r"(two ? idesc2 : idesc1), "r"(acc)
                    : "memory");
                if (!two && sy + 1 <= hi) {
                    const uint64_t db1 = oz_smem_desc(ys + (sy + 1) * OZ_PLANE, 128, 256);
                    asm volatile(
                        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                        " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base + uint32_t((sx + sy + 1 - d0) * OZ_TN)),
                        "l"(da), "l"(db1), "r"(idesc1), "r"(acc)
                        : "memory");
                }
            }
        }
    }
}
// Two passes over K ("sweeps") when T > 4: levels 0..3 first (digit planes 0..3 of both operands, 10 MMAs per k-block),
// then levels 4..T-1 (all planes, up to 26 MMAs).  Each sweep's level sum  sum_d 2^(-7 (d - d0)) acc_d  is EXACT in
// FP64 (<= 53 bits for K <= 8192; beyond, one rounding at 2^-53 of the sum), so the epilogue threads keep sweep 1 as one
// double per output while sweep 2 runs.
//
// Persistent: CTA c walks tiles c, c + gridDim.x, ...  The epilogue warps hand the accumulators back to the MMA
// warp as soon as they have pulled a sweep into registers (tmem_empty), so the arithmetic part of a tile's epilogue
// (scaling, support features, stores) runs while the tensor pipe is already in sweep 1 of the CTA's next tile, and
// the producer keeps the ring full across tile boundaries.
template <int T, class Epi>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_gemm_kernel(const OzOperands g, const Epi epi) {
    constexpr uint32_t SLOT = 2 * T * OZ_PLANE;                       // X planes then Y planes
    constexpr int G1 = T < OZ_G ? T : OZ_G;
    constexpr int NSWEEP = T > OZ_G ? 2 : 1;
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    unsigned char* ring = oz_smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(oz_smem + size_t(OZ_STAGES) * SLOT);
    uint64_t* empty = full + OZ_STAGES;
    uint64_t* tmem_full = empty + OZ_STAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    double* sYscale = reinterpret_cast<double*>(tmem_empty + 2);      // [128] 2^eY of the tile's columns
    unsigned char* epi_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sYscale + OZ_TN) + 15) & ~uintptr_t(15));   // 16-byte loads

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = g.tiles_m * g.tiles_n;

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, OZ_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == OZ_EPI_WARPS + 1) {   // the MMA warp owns the tensor-memory allocation (all 512 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == OZ_EPI_WARPS) {
        // ===== producer =====
        if (lane == 0) {
            const size_t xstep = size_t(g.tiles_m_total) * (T * OZ_PLANE), ystep = size_t(g.tiles_n_total) * (T * OZ_PLANE);
            int jb = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int8_t* gx = g.Xq + size_t(t % g.tiles_m) * (T * OZ_PLANE);
                const int8_t* gy = g.Yq + size_t(g.tile_n0 + t / g.tiles_m) * (T * OZ_PLANE);
                for (int sw = 0; sw < NSWEEP; ++sw) {
                    const uint32_t bytes = uint32_t(sw == 0 ? G1 : T) * OZ_PLANE;
                    for (int kb = 0; kb < g.nkb; ++kb, ++jb) {
                        const int s = jb % OZ_STAGES;
                        const uint32_t ph = (jb / OZ_STAGES) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], 2 * bytes);
                        unsigned char* dst = ring + size_t(s) * SLOT;
                        bulk_copy_g2s(dst, gx + size_t(kb) * xstep, bytes, &full[s]);
                        bulk_copy_g2s(dst + T * OZ_PLANE, gy + size_t(kb) * ystep, bytes, &full[s]);
                    }
                }
            }
        }
    } else if (warp == OZ_EPI_WARPS + 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int jb = 0;
            uint32_t nsweeps_done = 0;                            // sweeps issued so far (all tiles)
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                for (int sw = 0; sw < NSWEEP; ++sw, ++nsweeps_done) {
                    if (nsweeps_done > 0) {                       // the epilogue has pulled the previous sweep out of TMEM
                        mbar_wait(tmem_empty, (nsweeps_done - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int kb = 0; kb < g.nkb; ++kb, ++jb) {
                        const int s = jb % OZ_STAGES;
                        const uint32_t ph = (jb / OZ_STAGES) & 1;
                        mbar_wait(&full[s], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t xs = smem_u32(ring + size_t(s) * SLOT), ys = xs + T * OZ_PLANE;
                        if (g.no_pair) {
                            if (sw == 0) oz_issue_kblock<T, 0, false>(xs, ys, tmem_base, kb == 0);
                            else oz_issue_kblock<T, NSWEEP - 1, false>(xs, ys, tmem_base, kb == 0);
                        } else {
                            if (sw == 0) oz_issue_kblock<T, 0, true>(xs, ys, tmem_base, kb == 0);
                            else oz_issue_kblock<T, NSWEEP - 1, true>(xs, ys, tmem_base, kb == 0);
                        }
                        // frees the ring slot once the MMAs above have read it (implies fence::before_thread_sync)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(tmem_full)) : "memory");
                }
            }
        }
    } else {
        // ===== epilogue warps 0..15: thread = (output row, 32-column group) =====
        const int quarter = warp & 3, cgrp = warp >> 2;
        const int ml = quarter * 32 + lane;
        const int col0 = cgrp * OZ_CPT;
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(col0);
        uint32_t nfull = 0;                                       // accumulator hand-overs consumed so far
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int tm = t % g.tiles_m, tn = g.tile_n0 + t / g.tiles_m;
            const int m = tm * OZ_TM + ml;
            // every epilogue thread is past the previous tile's reads of the staged column data
            asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_WARPS * 32) : "memory");
            if (tid < OZ_TN) sYscale[tid] = oz_pow2(max(-1000, min(1000, g.eY[size_t(tn) * OZ_TN + tid])));
            epi.prepare(epi_smem, tn, tid);
            asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_WARPS * 32) : "memory");
            const int ex = g.eX[size_t(tm) * OZ_TM + ml] - 2 * (OZ_BITS - 1);
            typename Epi::State st;
            epi.begin(st, tn);
            double part[OZ_CPT];
#pragma unroll
            for (int sw = 0; sw < NSWEEP; ++sw, ++nfull) {
                const int nlev = sw == 0 ? G1 : T - OZ_G;
                mbar_wait_sleep(tmem_full, nfull & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int cc = 0; cc < OZ_CPT / 8; ++cc) {
                    uint32_t a[OZ_G][8];
#pragma unroll
                    for (int dd = 0; dd < OZ_G; ++dd) {
                        if (dd < nlev)
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                         : "=r"(a[dd][0]), "=r"(a[dd][1]), "=r"(a[dd][2]), "=r"(a[dd][3]), "=r"(a[dd][4]),
                                           "=r"(a[dd][5]), "=r"(a[dd][6]), "=r"(a[dd][7])
                                         : "r"(taddr + uint32_t(dd * OZ_TN + cc * 8)));
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        double v = oz_i2d(a[nlev - 1][jj]);
#pragma unroll
                        for (int dd = OZ_G - 2; dd >= 0; --dd)
                            if (dd < nlev - 1) v = fma(v, 1.0 / (1 << OZ_BITS), oz_i2d(a[dd][jj]));   // * 2^-8; exact (<= 53 bits) for K <= 8192
                        if (sw == 0) part[cc * 8 + jj] = v;
                        else part[cc * 8 + jj] = fma(v, 1.0 / double(1ull << (OZ_G * OZ_BITS)), part[cc * 8 + jj]);   // 2^-32
                    }
                }
                // the accumulators may be overwritten: the next sweep (of this tile or of the CTA's next tile) can start
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty);
            }
#pragma unroll
            for (int jl = 0; jl < OZ_CPT; ++jl) {
                const int j = col0 + jl;
                epi.accept(st, epi_smem, m, tn, j, oz_scale(part[jl] * sYscale[j], ex));
            }
            epi.finish(st, epi_smem, tm, tn, ml, cgrp);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == OZ_EPI_WARPS + 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int T, class Epi>
struct OzShape {
    static constexpr size_t SLOT = size_t(2) * T * OZ_PLANE;
    static constexpr size_t SMEM = OZ_STAGES * SLOT + (2 * OZ_STAGES + 4) * sizeof(uint64_t) + OZ_TN * sizeof(double) + 16 + Epi::SMEM;
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
    static thread_local int sms = 0;
    if (sms == 0) {
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    const int ntiles = g.tiles_m * g.tiles_n;
    static const int env_no_pair = getenv("REACH_OZ_NOPAIR") ? 1 : 0;
    OzOperands gg = g;
    if (env_no_pair) gg.no_pair = 1;
    kern<<<ntiles < sms ? ntiles : sms, OZ_THREADS, smem, stream>>>(gg, epi);      // persistent: one CTA per SM
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
// One CTA per 8-row group and k range (gridDim.y splits the 16-wide chunks of a row: a direction block of a few hundred
// rows would otherwise occupy a fifth of the SMs; every CTA re-derives the row maxima, 8 rows from L2).
// Phase 1: row maxima -> exponents (max|x| <= 127/128 2^e).  Phase 2: thread = (row, 16-wide k chunk):
// 16 values -> T x 16 digits, written as T 16-byte stores (8 consecutive rows = one 128-byte core matrix).
//   Q[((kb * tiles_total + tile) * T + s) * (TR*32) + ((g * 2 + c) * 8 + r) * 16 + b],   tile = row / TR, g = (row % TR) / 8
// Digits of x-hat = x 2^(7-e) in [-127, 127]: q_0 = floor, then radix-256 digits of the remainder in [0, 255] (the last one
// rounded to nearest, so it may be 256), then balanced from the last digit up: a digit >= 128 becomes digit - 256 and
// carries 1 into the next higher one.  Every digit ends in [-128, 127] (q_0 + carry <= 127 because x-hat <= 127 + 1/2 is
// excluded by the exponent), signs are as random as the data (no one-sided truncation bias in the dropped levels), and
// sum_s q_s 2^(-8s) = x-hat up to the final rounding at 2^(-8(T-1)-1).
template <int TR>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ X, int ldx, int rows_valid, int k_valid,
                                                       int T, int nkb, int tiles_total, int8_t* __restrict__ Q,
                                                       int* __restrict__ eout, int tile0 = 0) {
    // tile0: the rows of X are tiles tile0, tile0 + 1, ... of the plane buffer (and exponents eout[tile0 * TR ...])
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
            if (mx > 0.0 && mx < 1.0e308) {
                e = ilogb(mx) + 1;                                     // 2^(e-1) <= mx < 2^e
                if (oz_scale(mx, OZ_BITS - 1 - e) > double((1 << (OZ_BITS - 1)) - 1)) ++e;   // keep |x-hat| <= 127
            }
            s_e[warp] = e;
            if (blockIdx.y == 0) eout[tile0 * TR + row] = e;
        }
    }
    __syncthreads();
    const int tile = tile0 + row0 / TR, gidx = (row0 % TR) / 8;
    const int nchunks = nkb * 2;
    const int per = (nchunks + gridDim.y - 1) / gridDim.y;
    const int item_lo = 8 * per * blockIdx.y, item_hi = min(8 * nchunks, item_lo + 8 * per);
    constexpr double RADIX = double(1 << OZ_BITS);
    constexpr int HALF = 1 << (OZ_BITS - 1);
    for (int item = item_lo + threadIdx.x; item < item_hi; item += blockDim.x) {
        const int r = item & 7, chunk = item >> 3;
        const int row = row0 + r;
        const int kb = chunk >> 1, c = chunk & 1;
        const int k0 = chunk * 16;
        const int e = s_e[r];
        uint32_t w[OZ_MAX_T][4];
#pragma unroll
        for (int s = 0; s < OZ_MAX_T; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int k = k0 + b;
            double v = (row < rows_valid && k < k_valid) ? X[size_t(row) * ldx + k] : 0.0;
            v = oz_scale(v, OZ_BITS - 1 - e);                                  // |v| <= 127, exact
            int q[OZ_MAX_T];
            double f = floor(v);
            q[0] = int(f);
            v -= f;                                                            // [0, 1), exact
#pragma unroll
            for (int s = 1; s < OZ_MAX_T; ++s) {
                q[s] = 0;
                if (s < T) {
                    v *= RADIX;
                    f = (s == T - 1) ? rint(v) : floor(v);
                    q[s] = int(f);
                    v -= f;
                }
            }
#pragma unroll
            for (int s = OZ_MAX_T - 1; s >= 1; --s)
                if (s < T && q[s] >= HALF) { q[s] -= 2 * HALF; q[s - 1] += 1; }
#pragma unroll
            for (int s = 0; s < OZ_MAX_T; ++s) w[s][b >> 2] |= (uint32_t(q[s]) & 0xFFu) << (8 * (b & 3));
        }
        int8_t* q = Q + ((size_t(kb) * tiles_total + tile) * T) * (TR * 32) + ((gidx * 2 + c) * 8 + r) * 16;
#pragma unroll
        for (int s = 0; s < OZ_MAX_T; ++s)
            if (s < T) *reinterpret_cast<uint4*>(q + size_t(s) * (TR * 32)) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

}  // namespace reach
