// crt_gemm.cuh -- FP64-accurate GEMM on the INT8 tcgen05 tensor pipe through residue arithmetic ("Ozaki scheme II":
// Ozaki, Uchino, Imamura 2025, integer modular technique), with 2-CTA (cta_group::2) MMAs.
//
//     C[m][n] = sum_k X[m][k] * Y[n][k]            (the same NT contraction as nt_dgemm.cuh / oz_gemm.cuh)
//
// oz_gemm.cuh cuts every operand into T = 7 digit planes and needs T(T+1)/2 = 28 INT8 products per FP64 product.  The
// kernel runs at the chip's power / L2-to-SM limit, so the number of INT8 products is what sets its time.  Here each
// operand row is scaled by a power of two and rounded to an integer vector of 2-norm <= 2^alpha (resp. 2^beta),
//     x' = rint(x 2^sx),  y' = rint(y 2^sy),      |sum_k x'_k y'_k| <= ||x'|| ||y'|| <= 2^(alpha+beta) < P/4
// and the integer dot product C' is computed EXACTLY from its residues modulo N pairwise coprime moduli p_i <= 256
// (P = prod p_i): the residues of x', y' fit signed bytes, one INT8 GEMM per modulus gives C'_i = X_i Y_i' with INT32
// accumulators (exact for K <= 65472), and the Chinese remainder theorem rebuilds C' from u_i = C'_i mod p_i:
//     C'/P = frac( sum_i u_i w_i / P ),   w_i = (P/p_i) ((P/p_i)^-1 mod p_i)
// The fractions f_i = w_i/P are split f_i = f1_i + f2_i with f1_i a multiple of 2^-40: sum_i u_i f1_i is exact in FP64,
// its integer part drops out, and sum_i u_i f2_i (< 2^-28) is accurate to 2^-80.  N = 14 moduli (P ~ 2^110, alpha = beta = 54)
// round each operand element at 2^-54 of its row NORM -- below FP64's own rounding of the products -- with 14 INT8
// products instead of 28; N = 13 matches the digit-plane path's accuracy.  The result does not depend on summation order.
//
// Kernel shape: CTA pairs (cluster of 2, cta_group::2).  A pair owns a 256 x 256 output tile: CTA r holds rows
// [128 r, 128 r + 128) of X and of Y in its shared memory (16 KB + 16 KB per 128-byte k-block, SWIZZLE_128B images: half
// the L2 -> SM bytes per MAC of a 128 x 256 one-CTA tile; an unswizzled core-matrix layout ran the 2-CTA MMAs at a quarter
// of their rate), the leader issues M = 256, N = 256, K = 32 MMAs, and each CTA's tensor memory holds its 128 rows x 256
// columns, double buffered (2 x 256 columns), so modulus i+1 is multiplied while modulus i is drained.  Per tile the pair
// loops over the moduli.
//   warp 16   producer: per k-block two cp.async.bulk copies (own X rows, own Y rows) into a 4-stage ring
//   warp 17   leader: one thread issues the MMAs and commits (multicast) to both CTAs' barriers;
//             peer: one thread relays "my stage is full" to the leader's barrier (a bulk copy can only signal an
//             mbarrier of the CTA it writes to)
//   warps 0-7 drain: tcgen05.ld, every INT32 accumulator reduced mod p_i to a byte, parked in a global scratch (two
//             halves per CTA, thread = output row x 2 x 32 columns of each 128-column half)
//   warps 8-15 reconstruction of the pair's PREVIOUS tile from the other scratch half + the same epilogue functors as
//             oz_gemm.cuh (plain store, carry rows, fused LGG09 support features): off the accumulator hand-over path (with
//             all 16 warps doing both, the MMAs waited for the FP64 work: 2.6 ms per launch of C5 against 1.86 ms without
//             reconstruction)
//   warp 18   one thread streams the scratch back into a 4-stage shared-memory ring, one bulk copy per slice (the N
//             residue words of 4 outputs of each thread), so the reconstruction warps never wait for an L2 / DRAM round trip
#pragma once

#include "oz_gemm.cuh"

namespace reach {

constexpr int CRT_TM = 128, CRT_TN = 256, CRT_KB = 128, CRT_STAGES = 4;
constexpr int CRT_MAX_N = 16, CRT_MIN_N = 10, CRT_DEFAULT_N = 14;
constexpr int CRT_IMG = CRT_TM * CRT_KB;                  // bytes of one 128-row residue image per k-block (16 KB)
constexpr int CRT_SLOT = 2 * CRT_IMG;                     // X rows then Y rows
constexpr int CRT_MAX_K = 65472;                          // K 2^14 < 2^30
constexpr int CRT_EPI_HALF = OZ_EPI_WARPS * 32 / 2;       // threads that drain, threads that reconstruct (256 each)
constexpr int CRT_THREADS = OZ_THREADS + 32;              // + the scratch prefetch warp
constexpr int CRT_SCR_WORDS = 4 * 8 * CRT_EPI_HALF;       // u32 words of scratch per modulus and CTA (4 x 32 columns x 128 rows)
constexpr int CRT_SLICES = 32;                            // slices of a tile: (half, column group of the pair, word) = 2 x 2 x 8
constexpr int CRT_SCR_STAGES = 4;                         // shared-memory ring of slices (5 operand + 2 slice stages: same time)
constexpr int CRT_SLICE_BYTES = CRT_MAX_N * CRT_EPI_HALF * 4;   // ring stage (N KB are used)

// pairwise coprime, largest first (256 = 2^8, 255 = 3 5 17, 253 = 11 23, 247 = 13 19, 217 = 7 31, the rest prime)
__host__ __device__ constexpr int crt_modulus(int i) {
    constexpr int m[CRT_MAX_N] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193};
    return m[i];
}

struct CrtConsts {
    int N = 0;
    int alpha = 0, beta = 0;       // 2-norm budgets of the X and Y rows (bits)
    double Pd = 0.0;               // P rounded to FP64
    uint32_t p[CRT_MAX_N];
    uint32_t magic[CRT_MAX_N];     // floor(2^32 / p)
    uint32_t off[CRT_MAX_N];       // multiple of p, >= 2^30: makes the INT32 accumulators non-negative
    double f1[CRT_MAX_N], f2[CRT_MAX_N];
};

inline CrtConsts crt_make_consts(int N) {
    CrtConsts c{};
    c.N = N;
    unsigned __int128 P = 1;
    for (int i = 0; i < N; ++i) P *= (unsigned)crt_modulus(i);
    int lp = 0;
    for (unsigned __int128 t = P; t > 1; t >>= 1) ++lp;               // floor(log2 P)
    const int tot = lp - 2;                                           // ||x'|| ||y'|| <= 2^tot <= P/4
    c.alpha = (tot + 1) / 2;
    c.beta = tot - c.alpha;
    c.Pd = (double)P;
    for (int i = 0; i < N; ++i) {
        const unsigned p = (unsigned)crt_modulus(i);
        const unsigned __int128 Mi = P / p;
        unsigned inv = 0;
        const unsigned mi = (unsigned)(Mi % p);
        for (unsigned t = 1; t < p; ++t)
            if ((mi * t) % p == 1) { inv = t; break; }
        const unsigned __int128 w = Mi * inv;                         // < P
        // binary expansion of w / P: bits 1..40 -> f1, bits 41..120 -> f2 (rounded to FP64)
        unsigned __int128 r = w;
        double f1 = 0.0, f2 = 0.0;
        for (int b = 1; b <= 120; ++b) {
            r <<= 1;
            int bit = 0;
            if (r >= P) { r -= P; bit = 1; }
            if (bit) {
                if (b <= 40) f1 += ldexp(1.0, -b);
                else f2 += ldexp(1.0, -b);                            // exact while the partial sum spans <= 53 bits, rounded after
            }
        }
        c.p[i] = p;
        c.magic[i] = (uint32_t)((((uint64_t)1) << 32) / p);
        c.off[i] = (uint32_t)(((((uint64_t)1) << 30) / p + 1) * p);
        c.f1[i] = f1;
        c.f2[i] = f2;
    }
    return c;
}

struct CrtOperands {
    const int8_t* Xr;     // [N][tiles_m_total][nkb][16 row groups][8 rows][128 B, 16-byte chunks swizzled]
    const int8_t* Yr;     // [N][tiles_n_total][nkb][...]            (128-row tiles)
    const int* eX;        // [tiles_m_total*128]: value = C' 2^(eX[m] + eY[n])
    const int* eY;
    int nkb;              // k-blocks of 128
    int tiles_m, tiles_m_total;
    int tile_n0, tiles_n, tiles_n_total;   // this launch covers 128-row Y tiles [tile_n0, tile_n0 + tiles_n)
    uint32_t* scratch;    // [gridDim.x][2 halves][32 slices][N][256 threads]
    CrtConsts c;
    int32_t* dbg = nullptr;   // debug: raw accumulators of modulus dbg_mod of the first tile of pair 0, [2][128][256]
    int dbg_mod = 0;
    int flags = 0;            // debug / timing experiments: 1 = no reconstruction, 2 = no residue reduction (drain only)
    unsigned long long* tim = nullptr;   // debug: [gridDim.x][16] cycle counters of the warp roles (REACH_CRT_TIMING)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// bulk copy global -> the same shared-memory offset of every CTA in ctaMask; each destination's mbarrier (same offset) gets the bytes
__device__ __forceinline__ void bulk_copy_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n"
        " mbarrier.arrive.shared::cluster.b64 _, [ra];\n}" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// Every wait of this kernel: try_wait with a suspend-time hint (SASS: TRYWAIT, NANOSLEEP.SYNCS, re-check), so a waiting warp
// sleeps until the barrier's phase completes instead of spinning.  (Spinning waits of the 8 drain warps, which idle ~70 % of
// the time, took 40 % of all issued instructions and starved the reconstruction warps on the same schedulers.)
// Default (CTA-scope) semantics as in CUTLASS' ClusterBarrier also where the peer CTA arrives: the operands travel through
// the async proxy only (bulk copy -> shared memory -> MMA), and a cluster-scope acquire / release per k-block costs a
// memory fence each (measured: 1.1 us per k-block instead of 0.15).
__device__ __forceinline__ void crt_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(20000u) : "memory");
        if (!ok && ++spins > (1u << 20)) __trap();     // a protocol error must fail, not hang the GPU
    } while (!ok);
}

// c mod p in [0, p), for |c| < 2^30 (negp = -p)
__device__ __forceinline__ uint32_t crt_mod(uint32_t c, uint32_t negp, uint32_t magic, uint32_t off) {
    const uint32_t cu = c + off;                      // in (0, 2^31 + 2^30): residue unchanged (off is a multiple of p)
    const uint32_t q = __umulhi(cu, magic);           // floor(cu / p) or one less
    const uint32_t r = cu + q * negp;                 // [0, 2p)
    return min(r, r + negp);                          // unsigned: r - p wraps unless r >= p
}
// scratch stores: L2 only (read back through the async proxy by the prefetch thread)
__device__ __forceinline__ void crt_st_scratch(uint32_t* p, uint32_t v) {
    asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// SWIZZLE_128B K-major shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart, 16-byte chunk c
// of row r stored at chunk c ^ (r % 8).  start >> 4 at [0,14), LBO (unused) = 1 at [16,30), SBO >> 4 at [32,46), version 1 at
// [46,48), layout type 2 = SWIZZLE_128B at [61,64).  The k-th 32-byte slice of the rows starts 32 k bytes further.
__device__ __forceinline__ uint64_t crt_smem_desc(uint32_t saddr) {
    return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}

__device__ __forceinline__ double crt_byte_i2d(uint32_t word, int e) {
    // unsigned byte e of word -> double without the conversion pipe: 2^52 + u has u in its low mantissa bits
    const uint32_t u = __byte_perm(word, 0u, 0x4440u | uint32_t(e));
    return __hiloint2double(0x43300000, int(u)) - 4503599627370496.0;
}

// s1[e] = sum_k u_k f1_k (exact), s2[e] = sum_k u_k f2_k for the 4 outputs whose residue bytes are in word k of the slice
template <int NN>
__device__ __forceinline__ void crt_slice_sums(const uint32_t* sw, const CrtConsts& c, double (&s1)[4], double (&s2)[4]) {
    uint32_t wq[NN];
#pragma unroll
    for (int k = 0; k < NN; ++k) wq[k] = sw[k * CRT_EPI_HALF];
#pragma unroll
    for (int e = 0; e < 4; ++e) s1[e] = s2[e] = 0.0;
#pragma unroll
    for (int k = 0; k < NN; ++k) {
        const double f1 = c.f1[k], f2 = c.f2[k];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double ud = crt_byte_i2d(wq[k], e);
            s1[e] = fma(ud, f1, s1[e]);                       // exact: multiples of 2^-40 below 2^12
            s2[e] = fma(ud, f2, s2[e]);
        }
    }
}

#define CRT_T0() (timing ? clock64() : 0ll)
#define CRT_ACC(slot, t0) do { if (timing) tacc[slot] += clock64() - (t0); } while (0)

template <class Epi>
__global__ void __launch_bounds__(CRT_THREADS, 1) crt_gemm_kernel(const CrtOperands g, const Epi epi) {
    extern __shared__ __align__(1024) unsigned char crt_smem[];
    unsigned char* ring = crt_smem;
    uint32_t* sring = reinterpret_cast<uint32_t*>(crt_smem + size_t(CRT_STAGES) * CRT_SLOT);          // scratch slices
    uint64_t* full = reinterpret_cast<uint64_t*>(crt_smem + size_t(CRT_STAGES) * CRT_SLOT + size_t(CRT_SCR_STAGES) * CRT_SLICE_BYTES);
    uint64_t* pfull = full + CRT_STAGES;              // leader: the peer's stage is full
    uint64_t* empty = pfull + CRT_STAGES;
    uint64_t* tmem_full = empty + CRT_STAGES;         // [2]
    uint64_t* tmem_empty = tmem_full + 2;             // [2] (leader's are waited on)
    uint64_t* scr_ready = tmem_empty + 2;             // [2] drain warps -> prefetch thread: the tile's residues are in the scratch half
    uint64_t* scr_free = scr_ready + 2;               // [2] reconstruction -> drain warps: the half may be overwritten
    uint64_t* sfull = scr_free + 2;                   // [CRT_SCR_STAGES] slice ring
    uint64_t* sempty = sfull + CRT_SCR_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + CRT_SCR_STAGES);
    int* sYexp = reinterpret_cast<int*>(sempty + CRT_SCR_STAGES + 1);    // [128] (+ 128 unused: keeps the epilogue data 8-byte aligned)
    unsigned char* epi_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sYexp + 2 * OZ_TN) + 15) & ~uintptr_t(15));   // 16-byte loads

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Cluster = NQ CTA pairs (NQ = 1 or 2) that share the pair's 256 Y rows and own consecutive 256-row blocks of X: CTA
    // (pair q, rank r in the pair) fetches 1/NQ of the Y rows r needs and multicasts it to the CTAs with the same r.
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    const uint32_t rank = crank & 1, cq = crank >> 1, NQ = csize >> 1;
    const uint32_t lead = crank & ~1u;                // cluster rank of this pair's leader
    const int pair = blockIdx.x / int(csize), npairs = gridDim.x / int(csize);     // cluster index / number of clusters
    const int N = g.c.N;
    const int n2_lo = g.tile_n0 >> 1, n2_hi = (g.tile_n0 + g.tiles_n + 1) >> 1;
    const int mstep = 2 * int(NQ);                    // 128-row X tiles per cluster
    const int mgroups = (g.tiles_m + mstep - 1) / mstep;
    const int ngroups = mgroups * (n2_hi - n2_lo);
    const bool timing = g.tim != nullptr;
    long long tacc[4] = {0, 0, 0, 0};
    const long long t_begin = CRT_T0();

    if (tid == 0) {
        for (int s = 0; s < CRT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&pfull[s], 1); mbar_init(&empty[s], NQ); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], OZ_EPI_WARPS);            // 8 drain warps of each CTA
            mbar_init(&scr_ready[b], CRT_EPI_HALF);
            mbar_init(&scr_free[b], CRT_EPI_HALF + 1);          // the reconstruction threads and the prefetch thread
        }
        for (int s = 0; s < CRT_SCR_STAGES; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], OZ_EPI_WARPS / 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == OZ_EPI_WARPS + 1) {   // warp 17 of both CTAs allocates the pair's tensor memory (all 512 columns)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();               // both CTAs' barriers are initialised and both allocations are done
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    uint32_t* const scr_cta = g.scratch + size_t(blockIdx.x) * 2 * N * CRT_SCR_WORDS;

    if (warp == OZ_EPI_WARPS) {
        // ===== producer (both CTAs): own 128 rows of X and own 128 rows of the pair's 256 Y rows =====
        if (lane == 0) {
            const size_t img_stride = size_t(CRT_IMG);
            uint16_t ymask = 0;
            for (uint32_t q = 0; q < NQ; ++q) ymask |= uint16_t(1u << (2 * q + rank));
            uint32_t jb = 0;
            for (int grp = pair; grp < ngroups; grp += npairs) {
                int tm = (grp % mgroups) * mstep + int(cq) * 2 + int(rank);
                if (tm >= g.tiles_m) tm = 0;                             // odd tile count: the peer multiplies a dummy tile
                int tn = (n2_lo + grp / mgroups) * 2 + int(rank);
                if (tn >= g.tiles_n_total) tn = g.tiles_n_total - 1;     // odd tile count at the end of the stack
                for (int i = 0; i < N; ++i) {
                    const int8_t* gx = g.Xr + (size_t(i) * g.tiles_m_total + tm) * g.nkb * img_stride;
                    const int8_t* gy = g.Yr + (size_t(i) * g.tiles_n_total + tn) * g.nkb * img_stride;
                    for (int kb = 0; kb < g.nkb; ++kb, ++jb) {
                        const int s = jb % CRT_STAGES;
                        const uint32_t ph = (jb / CRT_STAGES) & 1;
                        crt_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], CRT_SLOT);
                        unsigned char* dst = ring + size_t(s) * CRT_SLOT;
                        // (L2 eviction hints -- evict_last for X, which every tile re-reads, evict_first for the Y stream -- measured: no
                        // gain resp. 6 % slower)
                        bulk_copy_g2s(dst, gx + size_t(kb) * img_stride, CRT_IMG, &full[s]);
                        if (NQ == 1) {
                            bulk_copy_g2s(dst + CRT_IMG, gy + size_t(kb) * img_stride, CRT_IMG, &full[s]);
                        } else {                      // my part of the Y rows, to every CTA of the cluster with my rank in its pair
                            const uint32_t part = CRT_IMG / NQ;
                            bulk_copy_g2s_mc(dst + CRT_IMG + cq * part, gy + size_t(kb) * img_stride + cq * part, part, &full[s], ymask);
                        }
                    }
                }
            }
        }
    } else if (warp == OZ_EPI_WARPS + 1) {
        if (lane == 0 && rank == 0) {
            // ===== MMA issuer (leader) =====
            // instruction descriptor: D = S32, A = B = signed 8-bit, K-major both, N = 256, M = 256 (128 rows per CTA)
            constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(CRT_TN >> 3) << 17) | (uint32_t(256 >> 4) << 24);
            const uint16_t allmask = uint16_t((1u << csize) - 1), pairmask = uint16_t(3u << lead);
            uint32_t jb = 0, u = 0;
            for (int grp = pair; grp < ngroups; grp += npairs) {
                for (int i = 0; i < N; ++i, ++u) {
                    const uint32_t buf = u & 1, use = u >> 1;
                    if (use > 0) {                                     // both CTAs have pulled the buffer's previous contents
                        const long long t0 = CRT_T0();
                        crt_wait(&tmem_empty[buf], (use - 1) & 1);
                        CRT_ACC(0, t0);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t dcol = tmem_base + buf * uint32_t(CRT_TN);
                    for (int kb = 0; kb < g.nkb; ++kb, ++jb) {
                        const int s = jb % CRT_STAGES;
                        const uint32_t ph = (jb / CRT_STAGES) & 1;
                        const long long t1 = CRT_T0();
                        crt_wait(&full[s], ph);
                        CRT_ACC(1, t1);
                        const long long t2 = CRT_T0();
                        crt_wait(&pfull[s], ph);
                        CRT_ACC(2, t2);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t xs = smem_u32(ring + size_t(s) * CRT_SLOT), ys = xs + CRT_IMG;
#pragma unroll
                        for (int h = 0; h < CRT_KB / 32; ++h) {
                            const uint64_t da = crt_smem_desc(xs + h * 32);
                            const uint64_t db = crt_smem_desc(ys + h * 32);
                            const uint32_t acc = (kb > 0 || h > 0) ? 1u : 0u;
                            if (g.flags & 16) continue;                  // timing experiment: no tensor work
                            asm volatile(
                                "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                                " tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                                : "memory");
                        }
                        // frees the ring slot in both CTAs once the MMAs above have read it
                        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                     ::"r"(smem_u32(&empty[s])), "h"(allmask) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32(&tmem_full[buf])), "h"(pairmask) : "memory");
                }
            }
            if (timing) { g.tim[blockIdx.x * 16 + 0] = tacc[0]; g.tim[blockIdx.x * 16 + 1] = tacc[1]; g.tim[blockIdx.x * 16 + 2] = tacc[2];
                          g.tim[blockIdx.x * 16 + 3] = clock64() - t_begin; }
        } else if (lane == 0) {
            // ===== relay (peer): stage s of my ring is full -> the leader's pfull[s] =====
            uint32_t jb = 0;
            for (int grp = pair; grp < ngroups; grp += npairs)
                for (int i = 0; i < N; ++i)
                    for (int kb = 0; kb < g.nkb; ++kb, ++jb) {
                        const int s = jb % CRT_STAGES;
                        crt_wait(&full[s], (jb / CRT_STAGES) & 1);
                        mbar_arrive_remote(&pfull[s], lead);
                    }
        }
    } else if (warp == OZ_EPI_WARPS + 2) {
        // ===== scratch prefetch: slice by slice from the global scratch half into the shared-memory ring =====
        if (lane == 0) {
            uint32_t js = 0, nt = 0;
            for (int grp = pair; grp < ngroups; grp += npairs, ++nt) {
                const uint32_t par = nt & 1;
                const int tm = (grp % mgroups) * mstep + int(cq) * 2 + int(rank);
                const int tn2 = n2_lo + grp / mgroups;
                crt_wait(&scr_ready[par], (nt >> 1) & 1);     // all 256 drain threads have written the tile's residues
                asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy stores -> async-proxy (bulk copy) reads
                if (tm >= g.tiles_m || (g.flags & 1)) { mbar_arrive(&scr_free[par]); continue; }
                const uint32_t* half = scr_cta + size_t(par) * N * CRT_SCR_WORDS;
                for (int sl = 0; sl < CRT_SLICES; ++sl) {
                    const int tn = tn2 * 2 + (sl >> 4);
                    if (tn < g.tile_n0 || tn >= g.tile_n0 + g.tiles_n) continue;
                    const int s = js % CRT_SCR_STAGES;
                    crt_wait(&sempty[s], ((js / CRT_SCR_STAGES) & 1) ^ 1);
                    const uint32_t bytes = uint32_t(N) * CRT_EPI_HALF * 4;
                    mbar_expect_tx(&sfull[s], bytes);
                    bulk_copy_g2s(sring + size_t(s) * (CRT_SLICE_BYTES / 4), half + size_t(sl) * N * CRT_EPI_HALF, bytes, &sfull[s]);
                    ++js;
                }
                mbar_arrive(&scr_free[par]);      // (the half is free once the reconstruction threads have arrived as well)
            }
        }
    } else {
        // ===== epilogue: warps 0..7 drain the accumulators, warps 8..15 rebuild the previous tile's outputs =====
        // Both groups map a thread to (output row = tensor-memory lane, two of the four 32-column groups of each 128-column
        // half): drain thread d and reconstruction thread 256 + d own the same outputs.
        // Scratch half: [slice = (half h, column group c2 of the pair, word w)][modulus][256 threads].
        const bool is_drain = warp < OZ_EPI_WARPS / 2;
        const int etid = tid & (CRT_EPI_HALF - 1);                       // 0..255 inside the group
        const int quarter = warp & 3, grp2 = (warp >> 2) & 1;            // lane quarter, pair of column groups
        const int ml = quarter * 32 + lane;
        if (is_drain) {
            const uint32_t tlane = tmem_base + (uint32_t(quarter * 32) << 16);
            uint32_t u = 0, nt = 0;
            for (int grp = pair; grp < ngroups; grp += npairs, ++nt) {
                const uint32_t par = nt & 1;
                { const long long t0 = CRT_T0();
                if (nt >= 2) crt_wait(&scr_free[par], ((nt >> 1) - 1) & 1);     // tile nt - 2 has been rebuilt from this half
                CRT_ACC(0, t0); }
                uint32_t* scr = scr_cta + size_t(par) * N * CRT_SCR_WORDS + etid;
                for (int i = 0; i < N; ++i, ++u) {
                    const uint32_t buf = u & 1, use = u >> 1;
                    const uint32_t p = g.c.p[i], negp = 0u - p, magic = g.c.magic[i], off = g.c.off[i];
                    const long long t1 = CRT_T0();
                    crt_wait(&tmem_full[buf], use & 1);
                    CRT_ACC(1, t1);
                    const long long t2 = CRT_T0();
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int hc = 0; hc < 4; ++hc) {                 // (half, column group of the pair)
                        if (g.flags & 2) break;
                        const int h = hc >> 1, cg = grp2 * 2 + (hc & 1);
                        uint32_t a[32];
                        const uint32_t taddr = tlane + buf * uint32_t(CRT_TN) + uint32_t(h * OZ_TN + cg * OZ_CPT);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                            : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
                              "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]),
                              "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]),
                              "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (g.dbg && grp == 0 && i == g.dbg_mod) {
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj)
                                g.dbg[(size_t(rank) * 128 + ml) * 256 + h * OZ_TN + cg * OZ_CPT + jj] = int32_t(a[jj]);
                        }
#pragma unroll
                        for (int w = 0; w < 8; ++w) {
                            const uint32_t r0 = crt_mod(a[4 * w], negp, magic, off), r1 = crt_mod(a[4 * w + 1], negp, magic, off);
                            const uint32_t r2 = crt_mod(a[4 * w + 2], negp, magic, off), r3 = crt_mod(a[4 * w + 3], negp, magic, off);
                            const uint32_t word = __byte_perm(__byte_perm(r0, r1, 0x0040), __byte_perm(r2, r3, 0x0040), 0x5410);
                            crt_st_scratch(scr + (size_t(hc * 8 + w) * N + i) * CRT_EPI_HALF, word);
                        }
                    }
                    // the accumulator may be overwritten: modulus i + 2 (or the pair's next tile) can start
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&tmem_empty[buf], lead);
                    CRT_ACC(2, t2);
                }
                asm volatile("fence.proxy.async;" ::: "memory");        // the residues are read back by bulk copies (async proxy)
                mbar_arrive(&scr_ready[par]);                         // release: this thread's residues of the tile are written
            }
            if (timing && tid == 0) { g.tim[blockIdx.x * 16 + 4] = tacc[0]; g.tim[blockIdx.x * 16 + 5] = tacc[1]; g.tim[blockIdx.x * 16 + 6] = tacc[2]; }
        } else {
            uint32_t js = 0, nt = 0;
            typename Epi::State st;
            for (int grp = pair; grp < ngroups; grp += npairs, ++nt) {
                const uint32_t par = nt & 1;
                const int tm = (grp % mgroups) * mstep + int(cq) * 2 + int(rank);
                const int tn2 = n2_lo + grp / mgroups;
                if (tm < g.tiles_m && !(g.flags & 1)) {                // not the dummy tile of an odd tile count
                    const int m = tm * OZ_TM + ml;
                    const int ex = g.eX[size_t(tm) * OZ_TM + ml];
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        const int tn = tn2 * 2 + h;
                        if (tn < g.tile_n0 || tn >= g.tile_n0 + g.tiles_n) continue;     // uniform over the CTA
                        asm volatile("bar.sync 2, %0;" ::"n"(CRT_EPI_HALF) : "memory");
                        if (etid == 0) sYexp[OZ_TN] = 1;                 // all column exponents of the half within +-450
                        asm volatile("bar.sync 2, %0;" ::"n"(CRT_EPI_HALF) : "memory");
                        if (etid < OZ_TN) {
                            const int ey = g.eY[size_t(tn) * OZ_TN + etid];
                            sYexp[etid] = ey;
                            if (ey < -450 || ey > 450) sYexp[OZ_TN] = 0;
                        }
                        epi.prepare(epi_smem, tn, etid);
                        asm volatile("bar.sync 2, %0;" ::"n"(CRT_EPI_HALF) : "memory");
                        epi.begin(st, tn);
                        // |ex + eY| <= 900: 2^(ex + eY) is a normal number and t P 2^(ex + eY) cannot overflow early (P < 2^126)
                        const bool fast = sYexp[OZ_TN] != 0 && ex >= -450 && ex <= 450;
#pragma unroll 1
                        for (int sl = 0; sl < 16; ++sl, ++js) {
                            const int c2 = sl >> 3, w = sl & 7, cg = grp2 * 2 + c2;
                            const int s = js % CRT_SCR_STAGES;
                            const long long t0 = CRT_T0();
                            crt_wait(&sfull[s], (js / CRT_SCR_STAGES) & 1);
                            CRT_ACC(0, t0);
                            const long long t1 = CRT_T0();
                            const uint32_t* sw = sring + size_t(s) * (CRT_SLICE_BYTES / 4) + etid;
                            double s1[4], s2[4];
                            if (g.flags & 8) { s1[0] = s1[1] = s1[2] = s1[3] = 0.25; s2[0] = s2[1] = s2[2] = s2[3] = double(sw[0]) * 1e-30; }
                            else switch (N) {                                 // compile-time trip counts: all loads first, no branches
                                case 10: crt_slice_sums<10>(sw, g.c, s1, s2); break;
                                case 11: crt_slice_sums<11>(sw, g.c, s1, s2); break;
                                case 12: crt_slice_sums<12>(sw, g.c, s1, s2); break;
                                case 13: crt_slice_sums<13>(sw, g.c, s1, s2); break;
                                case 14: crt_slice_sums<14>(sw, g.c, s1, s2); break;
                                case 15: crt_slice_sums<15>(sw, g.c, s1, s2); break;
                                default: crt_slice_sums<16>(sw, g.c, s1, s2); break;
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sempty[s]);      // the slice has been read into registers
                            CRT_ACC(1, t1);
                            const long long t2 = CRT_T0();
                            double v[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const double r1 = (s1[e] + 6755399441055744.0) - 6755399441055744.0;   // nearest integer (|s1| < 2^51)
                                const double t = (s1[e] - r1) + s2[e];   // C'/P in (-1/4, 1/4)
                                // value = t P 2^(ex + eY[j]); one exponent-field add unless an exponent is far out of range
                                const int et = ex + sYexp[cg * OZ_CPT + w * 4 + e];
                                v[e] = fast ? (t * g.c.Pd) * oz_pow2(et) : oz_scale(t * g.c.Pd, et);
                            }
                            if (!(g.flags & 4)) epi.accept4(st, epi_smem, m, tn, cg * OZ_CPT + w * 4, v);
                            else if (v[0] == 0.123) sYexp[0] = 1;
                            CRT_ACC(2, t2);
                        }
                        { const long long t3 = CRT_T0();
                        epi.template finish_g<2, 2>(st, epi_smem, tm, tn, ml, grp2);
                        CRT_ACC(3, t3); }
                    }
                }
                // a tile without reconstruction (dummy tile of an odd tile count): stay in step with the drain warps
                if (tm >= g.tiles_m || (g.flags & 1)) crt_wait(&scr_ready[par], (nt >> 1) & 1);
                mbar_arrive(&scr_free[par]);      // every slice of the tile has left the global scratch half
            }
            if (timing && etid == 0) { g.tim[blockIdx.x * 16 + 8] = tacc[0]; g.tim[blockIdx.x * 16 + 9] = tacc[1]; g.tim[blockIdx.x * 16 + 10] = tacc[2];
                                       g.tim[blockIdx.x * 16 + 11] = tacc[3]; g.tim[blockIdx.x * 16 + 12] = clock64() - t_begin; }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();               // the peer may still signal this CTA's barriers / read its shared memory until here
    if (warp == OZ_EPI_WARPS + 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <class Epi>
struct CrtShape {
    static constexpr size_t SMEM = size_t(CRT_STAGES) * CRT_SLOT + size_t(CRT_SCR_STAGES) * CRT_SLICE_BYTES +
                                   (3 * CRT_STAGES + 2 * CRT_SCR_STAGES + 10) * sizeof(uint64_t) + OZ_TN * sizeof(double) + 16 + Epi::SMEM;
};

inline size_t crt_scratch_bytes(int N, int sm_count) { return size_t(sm_count) * 2 * N * CRT_SCR_WORDS * sizeof(uint32_t); }

template <class Epi>
inline cudaError_t launch_crt_gemm(const CrtOperands& g, const Epi& epi, cudaStream_t stream) {
    auto kern = crt_gemm_kernel<Epi>;
    constexpr size_t smem = CrtShape<Epi>::SMEM;
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
    const int n2 = ((g.tile_n0 + g.tiles_n + 1) >> 1) - (g.tile_n0 >> 1);
    // REACH_CRT_CLUSTER=4: two pairs per cluster, Y rows multicast (3/4 of the L2 -> SM bytes, but only 132 of 148 SMs host such clusters)
    static const int env_cs = getenv("REACH_CRT_CLUSTER") ? atoi(getenv("REACH_CRT_CLUSTER")) : 0;
    const int cs = env_cs == 4 && g.tiles_m >= 3 ? 4 : 2;     // (measured on C5: 2.37 ms with 33 clusters of 4 against 2.22 ms with 74 pairs)
    const int groups = ((g.tiles_m + cs - 1) / cs) * n2;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(CRT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static thread_local int max_clusters[5] = {0, 0, 0, 0, 0};           // co-resident clusters of this kernel, by cluster size
    if (max_clusters[cs] == 0) {
        cfg.gridDim = dim3(cs * (sms / cs));
        int nc = 0;
        e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
        if (e != cudaSuccess) return e;
        max_clusters[cs] = std::max(1, nc);
    }
    cfg.gridDim = dim3(cs * std::max(1, std::min(max_clusters[cs], groups)));
    return cudaLaunchKernelEx(&cfg, kern, g, epi);
}

// ---- residues: FP64 rows -> N signed-byte residue images in the UMMA core-matrix layout -------------------
// One CTA per 8-row group and k range (as oz_slice_kernel).  Phase 1: row maximum and sum of squares -> the scale s with
// ||rint(x 2^s)||_2 <= 2^bits.  Phase 2: thread = (row, 16-wide k chunk): x' = rint(x 2^s) as a 64-bit integer, its
// residue modulo every p_i (compile-time moduli: the divisions become multiplications), written as N 16-byte stores.
//   Q[((i * tiles_total + tile) * nkb + kb) * (128*128) + g * 1024 + r * 128 + ((c ^ r) * 16) + b],  g = (row % 128) / 8,
//   r = row % 8, c = 16-byte chunk of the k-block: the SWIZZLE_128B image the MMA descriptors expect
template <int NMAX>
__global__ void __launch_bounds__(256) crt_slice_kernel(const double* __restrict__ X, int ldx, int rows_valid, int k_valid, int N,
                                                        int bits, int nkb, int tiles_total, int8_t* __restrict__ Q,
                                                        int* __restrict__ eout, int tile0 = 0) {
    __shared__ int s_s[8];
    const int row0 = blockIdx.x * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {   // phase 1: warp w scans row row0 + w
        const int row = row0 + warp;
        double mx = 0.0;
        if (row < rows_valid) {
            const double* x = X + size_t(row) * ldx;
            double m4[4] = {0.0, 0.0, 0.0, 0.0};                       // four loads in flight (the scans are L2 round trips)
            int k = lane;
            for (; k + 96 < k_valid; k += 128) {
#pragma unroll
                for (int u = 0; u < 4; ++u) m4[u] = fmax(m4[u], fabs(x[k + 32 * u]));
            }
            for (; k < k_valid; k += 32) m4[0] = fmax(m4[0], fabs(x[k]));
            mx = fmax(fmax(m4[0], m4[1]), fmax(m4[2], m4[3]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        int s = 0;
        if (mx > 0.0 && mx < 1.0e308) {
            const int em = ilogb(mx);                                  // 2^em <= mx < 2^(em+1)
            double ss = 0.0;
            const double* x = X + size_t(row) * ldx;
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
            int k = lane;
            for (; k + 96 < k_valid; k += 128) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double v = oz_scale(x[k + 32 * u], -em);     // |v| < 2
                    s4[u] = fma(v, v, s4[u]);
                }
            }
            for (; k < k_valid; k += 32) {
                const double v = oz_scale(x[k], -em);
                s4[0] = fma(v, v, s4[0]);
            }
            ss = (s4[0] + s4[1]) + (s4[2] + s4[3]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            // ||x|| = 2^em sqrt(ss) < 2^(em + en),  en = ilogb(sqrt(ss) (1 + 2^-20)) + 1
            const int en = ilogb(sqrt(ss) * 1.000001) + 1;
            s = bits - (em + en);
        }
        if (lane == 0) {
            s_s[warp] = s;
            if (blockIdx.y == 0) eout[tile0 * 128 + row] = -s;
        }
    }
    __syncthreads();
    const int tile = tile0 + row0 / 128, gidx = (row0 % 128) / 8;
    const int nchunks = nkb * 8;
    const int per = (nchunks + gridDim.y - 1) / gridDim.y;
    const int item_lo = 8 * per * blockIdx.y, item_hi = min(8 * nchunks, item_lo + 8 * per);
    for (int item = item_lo + threadIdx.x; item < item_hi; item += blockDim.x) {
        const int r = item & 7, chunk = item >> 3;
        const int row = row0 + r;
        const int kb = chunk >> 3, c = chunk & 7;
        const int k0 = chunk * 16;
        const int s = s_s[r];
        uint32_t w[NMAX][4];
#pragma unroll
        for (int i = 0; i < NMAX; ++i) w[i][0] = w[i][1] = w[i][2] = w[i][3] = 0u;
        // the thread's 16 values: eight 16-byte loads when the chunk is whole and aligned (all loads in flight at once)
        double xv[16];
        const double* xrow = X + size_t(row) * ldx + k0;
        if (row < rows_valid && k0 + 16 <= k_valid && (ldx & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const double2 t = reinterpret_cast<const double2*>(xrow)[b];
                xv[2 * b] = t.x;
                xv[2 * b + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int b = 0; b < 16; ++b) xv[b] = (row < rows_valid && k0 + b < k_valid) ? xrow[b] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const double v = xv[b];
            const long long xi = __double2ll_rn(oz_scale(v, s));      // |xi| <= 2^bits (1 + eps): exact
            const unsigned long long ax = xi < 0 ? (unsigned long long)(-xi) : (unsigned long long)xi;
            // |x'| < 2^62 in three limbs a 2^40 + b 2^20 + c: a c40 + b c20 + c < 2^31 is congruent to it (c40 = 2^40 mod p,
            // c20 = 2^20 mod p), so every modulus costs two multiply-adds and ONE reduction by a compile-time constant
            const uint32_t la = uint32_t(ax >> 40), lb = uint32_t(ax >> 20) & 0xFFFFFu, lc = uint32_t(ax) & 0xFFFFFu;
#pragma unroll
            for (int i = 0; i < NMAX; ++i) {
                const uint32_t p = uint32_t(crt_modulus(i));
                const uint32_t c20 = uint32_t((uint64_t(1) << 20) % p), c40 = uint32_t((uint64_t(1) << 40) % p);
                const uint32_t rr = (la * c40 + lb * c20 + lc) % p;    // la < 2^22: < 2^30 + 2^28 + 2^20
                int sr = xi < 0 ? -int(rr) : int(rr);                  // (-p, p)
                const int hs = int(p - 1) >> 1;
                if (sr > hs) sr -= int(p);
                if (sr < -int(p >> 1)) sr += int(p);
                w[i][b >> 2] |= (uint32_t(sr) & 0xFFu) << (8 * (b & 3));
            }
        }
        int8_t* q = Q + (size_t(tile) * nkb + kb) * CRT_IMG + gidx * 1024 + r * 128 + ((c ^ r) * 16);
        const size_t mstride = size_t(tiles_total) * nkb * CRT_IMG;
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
            if (i < N) *reinterpret_cast<uint4*>(q + size_t(i) * mstride) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
    }
}

inline void launch_crt_slice(const double* X, int ldx, int rows_valid, int rows_padded, int k_valid, const CrtConsts& c, bool is_y, int nkb,
                             int tiles_total, int8_t* Q, int* eout, int tile0, int ksplit, cudaStream_t st) {
    // only instantiate the moduli in use.  The k split is the caller's (enough CTAs for few rows): every CTA of a row group
    // repeats the two scans of its 8 rows, so splitting further costs more than it fills -- 2 x 2048 rows x 2000 columns:
    // k split 1 / 2 / 4 / 8 -> 0.156 / 0.178 / 0.222 / 0.336 ms (REACH_CRT_KSPLIT overrides for experiments)
    const int groups = rows_padded / 8;
    if (const char* e = getenv("REACH_CRT_KSPLIT")) ksplit = std::max(1, atoi(e));
    const dim3 grid(groups, ksplit);
    const int bits = is_y ? c.beta : c.alpha;
#define CRT_SLICE_CASE(NN) case NN: crt_slice_kernel<NN><<<grid, 256, 0, st>>>(X, ldx, rows_valid, k_valid, c.N, bits, nkb, tiles_total, Q, eout, tile0); break
    switch (c.N) {
        CRT_SLICE_CASE(10); CRT_SLICE_CASE(11); CRT_SLICE_CASE(12); CRT_SLICE_CASE(13); CRT_SLICE_CASE(14); CRT_SLICE_CASE(15);
        default: crt_slice_kernel<CRT_MAX_N><<<grid, 256, 0, st>>>(X, ldx, rows_valid, k_valid, c.N, bits, nkb, tiles_total, Q, eout, tile0);
    }
#undef CRT_SLICE_CASE
}

}  // namespace reach
