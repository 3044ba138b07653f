// lgg09_kernels.cuh -- device side of the LGG09 support-function recurrence.
//
// Reference loop (src/Algorithms/LGG09/reach_inhomog.jl:49-64), per direction l:
//     r_0 = l;   rho[j,i] = rho(r_i, Omega0) + s_i;   s_{i+1} = s_i + rho(r_i, U);   r_{i+1} = Phi' r_i
//
// Device formulation.  Every support term of the flattened Omega0 / V program is a "feature"
//     F_q(r) = lin_q(r) + abs_q(r),   lin_q(r) = sum_i a_q[i] y_i,   abs_q(r) = sum_i b_q[i] |y_i|
// over the rows y = [I ; E] r (E = the mapped rows M' / G' of LinearMap / Zonotope terms).  One
// pass of nt_dgemm over the row stack [I;E](Phi')^s, s = 1..S, yields y for S consecutive steps
// of every direction chain; the epilogue below reduces the accumulator tile to per-row-tile
// partial (lin, abs) sums, and stores the s = S block as the next direction block.  The combine
// kernel then walks the S steps sequentially per chain: alternatives (max), Minkowski terms
// (sum), the running input sum s_i, and the +l / -l outputs of a shared chain
// (rho(-r) = -lin + abs).
#pragma once

#include "nt_dgemm.cuh"

namespace reach {

constexpr int LGG09_MAX_FEATS = 16;
constexpr int LGG09_MAX_ALTS = 16;

struct Lgg09Epi {
    double* Lnext;              // [mu_pad][ldl]: rows = chains; written from the store block
    int ldl;
    int n;                      // state dimension (columns < n of the store block are stored)
    int nbe;                    // rows per stack block (multiple of TN)
    int block0;                 // absolute index of the first block of this launch (0 or 1)
    int store_block;            // absolute block index whose C tile is L_{k0+S}; -1: none
    const double* Wl;           // [nq][nbe] linear weights per stack-block row
    const double* Wa;           // [nq][nbe] absolute-value weights
    const uint32_t* tile_mask;  // [nbe/TN] features with a non-zero weight in that row tile
    int nq;
    double* featpart;           // [blocks][nbe/TN][nq][2][mu_pad]
    int mu_pad;

    template <int MT, int NT, int WN, class TC>
    __device__ __forceinline__ void run(double (&acc)[MT][NT][2], const TC& tc, double* smem) const {
        constexpr int TM = TC::TM, TN = TC::TN;
        const int row0 = tc.tn * TN;                 // row inside this launch's stack slice
        const int blk = block0 + row0 / nbe;         // absolute block
        const int rib0 = row0 % nbe;                 // row in block of the tile's first row
        const int tib = rib0 / TN;                   // tile in block
        const int tpb = nbe / TN;

        if (blk == store_block) {
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int m = tc.tm * TM + tc.wm0 + i * 8 + tc.g4;
                double* o = Lnext + size_t(m) * ldl;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int c = rib0 + tc.wn0 + j * 8 + 2 * tc.t4;
                    if (c + 1 < n) {
                        *reinterpret_cast<double2*>(o + c) = make_double2(acc[i][j][0], acc[i][j][1]);
                    } else if (c < n) {
                        o[c] = acc[i][j][0];
                    }
                }
            }
        }

        const uint32_t mask = tile_mask[tib];
        if (mask == 0u) return;
        // red[q'][wn][part][TM] for the active features q' (compacted)
        double* red = smem;
        const int wn = tc.wn0 / (NT * 8);
        int qa = 0;
        for (int q = 0; q < nq; ++q) {
            if (!((mask >> q) & 1u)) continue;
            double wl[NT][2], wa[NT][2];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int c = rib0 + tc.wn0 + j * 8 + 2 * tc.t4;
                const double2 l2 = *reinterpret_cast<const double2*>(Wl + size_t(q) * nbe + c);
                const double2 a2 = *reinterpret_cast<const double2*>(Wa + size_t(q) * nbe + c);
                wl[j][0] = l2.x; wl[j][1] = l2.y;
                wa[j][0] = a2.x; wa[j][1] = a2.y;
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                double lin = 0.0, ab = 0.0;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    lin = fma(wl[j][0], acc[i][j][0], lin);
                    lin = fma(wl[j][1], acc[i][j][1], lin);
                    ab = fma(wa[j][0], fabs(acc[i][j][0]), ab);
                    ab = fma(wa[j][1], fabs(acc[i][j][1]), ab);
                }
                lin += __shfl_xor_sync(0xffffffffu, lin, 1);
                ab += __shfl_xor_sync(0xffffffffu, ab, 1);
                lin += __shfl_xor_sync(0xffffffffu, lin, 2);
                ab += __shfl_xor_sync(0xffffffffu, ab, 2);
                if (tc.t4 == 0) {
                    const int m = tc.wm0 + i * 8 + tc.g4;
                    red[((qa * WN + wn) * 2 + 0) * TM + m] = lin;
                    red[((qa * WN + wn) * 2 + 1) * TM + m] = ab;
                }
            }
            ++qa;
        }
        tc.sync();
        // fixed-order sum over the WN warps that share a direction row -> deterministic
        const int nact = qa;
        for (int idx = tc.tid; idx < nact * 2 * TM; idx += tc.nthr) {
            const int m = idx % TM;
            const int part = (idx / TM) & 1;
            const int qi = idx / (2 * TM);
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < WN; ++w) s += red[((qi * WN + w) * 2 + part) * TM + m];
            // map compacted index back to the feature id
            int q = 0, seen = 0;
            for (; q < nq; ++q)
                if ((mask >> q) & 1u) {
                    if (seen == qi) break;
                    ++seen;
                }
            featpart[(((size_t(blk) * tpb + tib) * nq + q) * 2 + part) * mu_pad + tc.tm * TM + m] = s;
        }
    }
};

// Packed variant of the epilogue (large n): stack blocks follow each other with a stride of
// nbe = round_up(n + ne, 8) rows instead of being padded to the tile height, so no DMMA work is spent
// on padding rows (n = 2000: 2.3 %).  A 128-row tile can then straddle ONE block boundary (nbe >= TN);
// the boundary is a multiple of 8, so every 8-row MMA sub-tile -- hence every (warp, j) -- belongs to
// one block, and the tile emits two partial sums per feature: segment 0 (rows of the tile's first
// block) and segment 1 (rows of the next block).
struct Lgg09PackedEpi {
    double* Lnext;
    int ldl, n, nbe;
    int slice_block0;           // absolute block index of the first row of this launch's stack slice
    int store_block;            // absolute block whose rows are L_{k0+S}; -1: none
    const double* Wl;           // [nq][nbe]
    const double* Wa;
    int nq;
    double* featpart;           // [tiles_n][2 segments][nq][2][mu_pad]
    int mu_pad;

    template <int MT, int NT, int WN, class TC>
    __device__ __forceinline__ void run(double (&acc)[MT][NT][2], const TC& tc, double* smem) const {
        constexpr int TM = TC::TM, TN = TC::TN;
        const int row0 = tc.tn * TN;
        const int blk0 = row0 / nbe;
        const int bnd = (blk0 + 1) * nbe - row0;     // tile-local row where the next block starts (>= TN: none)
        int rib[NT];
        bool seg1[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int cl = tc.wn0 + j * 8 + 2 * tc.t4;
            seg1[j] = (tc.wn0 + j * 8) >= bnd;       // warp-uniform
            rib[j] = row0 + cl - (blk0 + (seg1[j] ? 1 : 0)) * nbe;
        }
        // store the rows that belong to the store block
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (slice_block0 + blk0 + (seg1[j] ? 1 : 0) != store_block) continue;
            const int c = rib[j];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int m = tc.tm * TM + tc.wm0 + i * 8 + tc.g4;
                double* o = Lnext + size_t(m) * ldl + c;
                if (c + 1 < n) *reinterpret_cast<double2*>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
                else if (c < n) o[0] = acc[i][j][0];
            }
        }
        double* red = smem;                          // [q][seg][wn][part][TM]
        const int wn = tc.wn0 / (NT * 8);
        for (int q = 0; q < nq; ++q) {
            double wl[NT][2], wa[NT][2];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double2 l2 = *reinterpret_cast<const double2*>(Wl + size_t(q) * nbe + rib[j]);
                const double2 a2 = *reinterpret_cast<const double2*>(Wa + size_t(q) * nbe + rib[j]);
                wl[j][0] = l2.x; wl[j][1] = l2.y;
                wa[j][0] = a2.x; wa[j][1] = a2.y;
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                double lin[2] = {0.0, 0.0}, ab[2] = {0.0, 0.0};
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    double l = fma(wl[j][0], acc[i][j][0], 0.0);
                    l = fma(wl[j][1], acc[i][j][1], l);
                    double a = fma(wa[j][0], fabs(acc[i][j][0]), 0.0);
                    a = fma(wa[j][1], fabs(acc[i][j][1]), a);
                    if (seg1[j]) { lin[1] += l; ab[1] += a; } else { lin[0] += l; ab[0] += a; }
                }
#pragma unroll
                for (int sg = 0; sg < 2; ++sg) {
                    lin[sg] += __shfl_xor_sync(0xffffffffu, lin[sg], 1);
                    ab[sg] += __shfl_xor_sync(0xffffffffu, ab[sg], 1);
                    lin[sg] += __shfl_xor_sync(0xffffffffu, lin[sg], 2);
                    ab[sg] += __shfl_xor_sync(0xffffffffu, ab[sg], 2);
                }
                if (tc.t4 == 0) {
                    const int m = tc.wm0 + i * 8 + tc.g4;
#pragma unroll
                    for (int sg = 0; sg < 2; ++sg) {
                        red[(((q * 2 + sg) * WN + wn) * 2 + 0) * TM + m] = lin[sg];
                        red[(((q * 2 + sg) * WN + wn) * 2 + 1) * TM + m] = ab[sg];
                    }
                }
            }
        }
        tc.sync();
        for (int idx = tc.tid; idx < nq * 4 * TM; idx += tc.nthr) {
            const int m = idx % TM;
            const int part = (idx / TM) & 1;
            const int sg = (idx / (2 * TM)) & 1;
            const int q = idx / (4 * TM);
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < WN; ++w) v += red[(((q * 2 + sg) * WN + w) * 2 + part) * TM + m];
            featpart[((((size_t(tc.tn) * 2 + sg) * nq + q) * 2) + part) * mu_pad + tc.tm * TM + m] = v;
        }
    }
};

struct Lgg09Program {
    int nalt;
    int box;                        // 1: the two outputs of a chain are (lin, abs) = (centre, radius) of the box hull --
                                    // the BOX algorithm (Algorithms/BOX/reach_homog.jl, reach_inhomog.jl) -- instead of
                                    // (rho(+l), rho(-l)) = (abs + lin, abs - lin)
    int qV;                         // feature of rho(., V), -1 when homogeneous
    int8_t q0[LGG09_MAX_ALTS];      // per alternative: feature evaluated at r_k     (-1: none)
    int8_t q1[LGG09_MAX_ALTS];      // per alternative: feature evaluated at r_{k+1} (-1: none)
};

__device__ __forceinline__ double lgg09_out_pos(const Lgg09Program& p, double lin, double ab) { return p.box ? lin : lin + ab; }
__device__ __forceinline__ double lgg09_out_neg(const Lgg09Program& p, double lin, double ab) { return p.box ? ab : ab - lin; }

struct Lgg09Combine {
    Lgg09Program prog;
    const double* featpart;         // [blocks][tpb][nq][2][mu_pad]
    const uint32_t* tile_mask;      // [tpb]
    int tpb, nq, mu_pad, mu;
    int b_lo, b_hi;                 // blocks present in featpart for this pass (inclusive)
    int has_carry;                  // F(r_{k0}) comes from `carry` instead of block b_lo
    int tail;                       // also emit step k0 + b_hi from the last features alone
    long long k0;                   // r index of block 0 of this pass
    long long nsteps;
    double* s_pos;                  // [mu_pad] running input sums of the +chain / -chain outputs
    double* s_neg;
    const int* row_pos;             // [mu] output row of +chain (-1: absent)
    const int* row_neg;             // [mu] output row of -chain (-1: absent)
    double* rho;                    // ndirs x nsteps column-major
    int ndirs;
};

// ---- combine, phase A: per-block feature totals -----------------------------------------------
// Ftot[b][q][part][j] = sum over the row tiles of block b of the per-tile partials, in fixed tile
// order (deterministic).  One thread per (chain j, block b).
static __global__ void lgg09_tilesum_kernel(const Lgg09Combine c, double* __restrict__ ftot) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = c.b_lo + blockIdx.y;
    if (j >= c.mu || b > c.b_hi) return;
    for (int q = 0; q < c.nq; ++q) {
        double lin = 0.0, ab = 0.0;
        for (int t = 0; t < c.tpb; ++t) {
            if (!((c.tile_mask[t] >> q) & 1u)) continue;
            const size_t base = (((size_t(b) * c.tpb + t) * c.nq + q) * 2) * c.mu_pad + j;
            lin += c.featpart[base];
            ab += c.featpart[base + c.mu_pad];
        }
        ftot[((size_t(b) * c.nq + q) * 2 + 0) * c.mu_pad + j] = lin;
        ftot[((size_t(b) * c.nq + q) * 2 + 1) * c.mu_pad + j] = ab;
    }
}

// Packed layout: block b (relative to the launch slice) spans rows [rel*nbe, (rel+1)*nbe); it collects
// segment 0 of the tiles that start inside it and segment 1 of the tile that starts in the previous block.
static __global__ void lgg09_tilesum_packed_kernel(const Lgg09Combine c, const double* __restrict__ featpart,
                                                   double* __restrict__ ftot, int nbe, int tn_rows, int blk0) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = c.b_lo + blockIdx.y;
    if (j >= c.mu || b > c.b_hi) return;
    const int rel = b - blk0;                       // blk0: block that starts at row 0 of the tiled row range
    const int t0 = (rel * nbe) / tn_rows, t1 = ((rel + 1) * nbe - 1) / tn_rows;
    for (int q = 0; q < c.nq; ++q) {
        double lin = 0.0, ab = 0.0;
        for (int t = t0; t <= t1; ++t) {
            const int sg = ((t * tn_rows) / nbe == rel) ? 0 : 1;
            const size_t base = ((((size_t(t) * 2 + sg) * c.nq + q) * 2)) * c.mu_pad + j;
            lin += featpart[base];
            ab += featpart[base + c.mu_pad];
        }
        ftot[((size_t(b) * c.nq + q) * 2 + 0) * c.mu_pad + j] = lin;
        ftot[((size_t(b) * c.nq + q) * 2 + 1) * c.mu_pad + j] = ab;
    }
}

// ---- guard of the INT8 digit-plane path: feature totals of a pass against the same blocks on the DMMA kernel ----
// The first and the last block of the pass are sampled.  Per (chain, feature): max over the two blocks of
// |d lin| + |d abs|, relative to the largest |lin| + |abs| of that chain and feature in them; the worst ratio is kept with an integer atomicMax (non-negative doubles order
// like their bit patterns).
static __global__ void lgg09_guard_kernel(const double* __restrict__ ftot, const double* __restrict__ ftot_ref, int nq,
                                          int mu_pad, int mu, int b_first, int b_last, unsigned long long* __restrict__ worst,
                                          int stride = 1) {
    // stride > 1: row j of ftot_ref (a sentinel chain) is compared with chain j * stride of ftot
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= mu * nq) return;
    const int j = idx % mu, q = idx / mu;
    double scale = 0.0, err = 0.0;
    for (int b = b_first; b <= b_last; b += (b_last > b_first ? b_last - b_first : 1)) {   // the two sampled blocks
        const size_t row = ((size_t(b) * nq + q) * 2) * mu_pad;
        const size_t base = row + j, mbase = row + size_t(j) * stride;
        const double lr = ftot_ref[base], ar = ftot_ref[base + mu_pad];
        scale = fmax(scale, fabs(lr) + fabs(ar));
        err = fmax(err, fabs(ftot[mbase] - lr) + fabs(ftot[mbase + mu_pad] - ar));
    }
    double ratio = scale > 0.0 ? err / scale : (err > 0.0 ? 1.0e300 : 0.0);
    if (!(ratio >= 0.0)) ratio = 1.0e300;                  // NaN
    atomicMax(worst, (unsigned long long)__double_as_longlong(ratio));
}

// ---- combine, phase B: support values of Omega0 and V per (chain, step), fully parallel --------
// Emit slot e of a pass pairs the features of r_k ("prev") with those of r_{k+1} ("cur"):
//   has_carry: slot 0 = (carry, block b_lo), slot e = (block b_lo+e-1, block b_lo+e)
//   otherwise: slot e = (block b_lo+e, block b_lo+e+1)
//   tail     : one more slot (last, last) for a final step that needs nothing at r_{k+1}
// rho(r_k; Omega0) + s_k goes straight into the output matrix (s_k comes from the running-sum pass,
// which runs first).  The thread of the last slot also writes the next pass's carry.
template <int NQ>
__global__ void lgg09_eval_kernel(const Lgg09Combine c, const double* __restrict__ ftot,
                                  const double* __restrict__ carry_in, double* __restrict__ carry_out,
                                  const double* __restrict__ sbuf) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y;
    if (j >= c.mu) return;
    const int nblk = c.b_hi - c.b_lo + 1;                       // may be 0
    const int npair = c.has_carry ? nblk : (nblk > 0 ? nblk - 1 : 0);
    const bool is_tail = (e == npair);                          // only launched when c.tail
    int bp, bc;                                                 // block ids; -1 = carry
    if (is_tail) {
        bp = bc = (nblk > 0 ? c.b_hi : -1);
    } else if (c.has_carry) {
        bp = e == 0 ? -1 : c.b_lo + e - 1;
        bc = c.b_lo + e;
    } else {
        bp = c.b_lo + e;
        bc = c.b_lo + e + 1;
    }
    const long long k = is_tail ? c.k0 + c.b_hi : c.k0 + bc - 1;
    double Fp[NQ][2], Fc[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        Fp[q][0] = Fp[q][1] = Fc[q][0] = Fc[q][1] = 0.0;
        if (q < c.nq) {
            const double* sp = bp < 0 ? carry_in + (size_t(q) * 2) * c.mu_pad
                                      : ftot + ((size_t(bp) * c.nq + q) * 2) * c.mu_pad;
            const double* sc = bc < 0 ? carry_in + (size_t(q) * 2) * c.mu_pad
                                      : ftot + ((size_t(bc) * c.nq + q) * 2) * c.mu_pad;
            Fp[q][0] = sp[j]; Fp[q][1] = sp[c.mu_pad + j];
            Fc[q][0] = sc[j]; Fc[q][1] = sc[c.mu_pad + j];
        }
    }
    if (k < c.nsteps) {
        double vp = 0.0, vn = 0.0;
        for (int a = 0; a < c.prog.nalt; ++a) {
            double ap = 0.0, an = 0.0;
            const int q0 = c.prog.q0[a], q1 = c.prog.q1[a];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (q == q0) { ap += lgg09_out_pos(c.prog, Fp[q][0], Fp[q][1]); an += lgg09_out_neg(c.prog, Fp[q][0], Fp[q][1]); }
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (q == q1) { ap += lgg09_out_pos(c.prog, Fc[q][0], Fc[q][1]); an += lgg09_out_neg(c.prog, Fc[q][0], Fc[q][1]); }
            if (a == 0) { vp = ap; vn = an; } else { vp = fmax(vp, ap); vn = fmax(vn, an); }
        }
        if (c.prog.qV >= 0) {                                   // + s_k from the running-sum pass
            vp += sbuf[(size_t(e) * 2 + 0) * c.mu_pad + j];
            vn += sbuf[(size_t(e) * 2 + 1) * c.mu_pad + j];
        }
        const int rp = c.row_pos[j], rn = c.row_neg[j];
        if (rp >= 0) c.rho[size_t(k) * c.ndirs + rp] = vp;
        if (rn >= 0) c.rho[size_t(k) * c.ndirs + rn] = vn;
    }
    // the features of the last r of this pass seed the next one
    const int nslots = npair + (c.tail ? 1 : 0);
    if (e == nslots - 1) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (q < c.nq) {
                carry_out[(size_t(q) * 2 + 0) * c.mu_pad + j] = Fc[q][0];
                carry_out[(size_t(q) * 2 + 1) * c.mu_pad + j] = Fc[q][1];
            }
    }
}

// ---- combine, running input sum, strictly in step order (runs between tile-sum and eval) ------------
// s_{k+1} = s_k + rho(r_k; V)   (reach_inhomog.jl:56-57); sbuf[e][0/1][j] = s_k of the +chain / -chain for
// emit slot e.  rho(+-r_k; V) = abs_qV(r_k) +- lin_qV(r_k) is read from the block totals (or the carry).
// A CTA owns 16 chains.  Warps 1..7 stream the (lin, abs) totals of the slots into a 4-deep shared-memory
// ring with cp.async, three chunks of 32 slots ahead; warp 0 runs the dependent adds (lane = chain and sign,
// one add per step: the reference's order), so the chain of adds never waits for a global load.
constexpr int PREFIX_CB = 16, PREFIX_CHK = 32, PREFIX_NBUF = 4, PREFIX_THREADS = 256;
static __global__ void __launch_bounds__(PREFIX_THREADS) lgg09_prefix_kernel(const Lgg09Combine c, const double* __restrict__ ftot,
                                                                             const double* __restrict__ carry_in,
                                                                             double* __restrict__ sbuf, int nslots, long long kfirst) {
    __shared__ __align__(16) double sv[PREFIX_NBUF][PREFIX_CHK][2][PREFIX_CB];
    const int lane = threadIdx.x & 31;
    const bool scanner = threadIdx.x < 32;
    constexpr int NLOAD = PREFIX_THREADS - 32, ITEMS = PREFIX_CHK * 2 * PREFIX_CB;
    const int nblk = c.b_hi - c.b_lo + 1;
    const int npair = c.has_carry ? nblk : (nblk > 0 ? nblk - 1 : 0);
    const int qv = c.prog.qV;
    // Loader fast path.  In an interior chunk every slot is regular (not the carried first slot, not the tail slot, inside
    // the run), so the source block advances by one per slot and a thread's items keep fixed offsets from the chunk's
    // base pointer: a few instructions per copy instead of ~50 (ncu on Building: the seven loader warps, not the adds of
    // the scanner warp, bounded the kernel -- 2.6 k warp instructions per 32-slot chunk, 'wait' stalls).
    constexpr int NT_LOAD = (ITEMS + NLOAD - 1) / NLOAD;
    long long f_off[NT_LOAD];
    unsigned f_dst[NT_LOAD];
    bool f_ok[NT_LOAD];
#pragma unroll
    for (int t = 0; t < NT_LOAD; ++t) {
        const int item = int(threadIdx.x) - 32 + NLOAD * t;
        const int i = item / (2 * PREFIX_CB), part = (item / PREFIX_CB) & 1, ch = item % PREFIX_CB;
        const int j = blockIdx.x * PREFIX_CB + ch;
        f_ok[t] = !scanner && item < ITEMS && j < c.mu;
        f_off[t] = ((long long)i * c.nq * 2 + part) * c.mu_pad + j;
        f_dst[t] = f_ok[t] ? smem_u32(&sv[0][i][part][ch]) : 0u;
    }
    const long long regular_end = min((long long)min(npair, nslots), c.nsteps - kfirst);   // slots [.., regular_end) are regular
    auto stage = [&](int buf, int e0) {                      // loader warps: totals of the "prev" block of each slot
        if (e0 < nslots && e0 + PREFIX_CHK <= regular_end && (e0 > 0 || !c.has_carry)) {
            const int bp0 = c.has_carry ? c.b_lo + e0 - 1 : c.b_lo + e0;
            const double* base = ftot + ((size_t(bp0) * c.nq + qv) * 2) * c.mu_pad;
            const unsigned boff = unsigned(buf) * unsigned(sizeof(double) * PREFIX_CHK * 2 * PREFIX_CB);
#pragma unroll
            for (int t = 0; t < NT_LOAD; ++t)
                if (f_ok[t])
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(f_dst[t] + boff), "l"(base + f_off[t]) : "memory");
        } else if (e0 < nslots) {
#pragma unroll
            for (int t = 0; t < (ITEMS + NLOAD - 1) / NLOAD; ++t) {
                const int item = int(threadIdx.x) - 32 + NLOAD * t;
                if (item >= ITEMS) break;
                const int i = item / (2 * PREFIX_CB), part = (item / PREFIX_CB) & 1, ch = item % PREFIX_CB;
                const int e = e0 + i, j = blockIdx.x * PREFIX_CB + ch;
                double* dst = &sv[buf][i][part][ch];
                if (j < c.mu && e < nslots && kfirst + e < c.nsteps) {
                    int bp;
                    if (e == npair) bp = nblk > 0 ? c.b_hi : -1;               // tail slot
                    else if (c.has_carry) bp = e == 0 ? -1 : c.b_lo + e - 1;
                    else bp = c.b_lo + e;
                    const double* src = (bp < 0 ? carry_in + (size_t(qv) * 2) * c.mu_pad
                                                : ftot + ((size_t(bp) * c.nq + qv) * 2) * c.mu_pad) + size_t(part) * c.mu_pad + j;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
                } else {
                    *dst = 0.0;
                }
            }
        }
        cp_async_commit();                                   // one group per chunk, empty or not
    };
    const int ch = lane % PREFIX_CB, sign = lane / PREFIX_CB;
    const int j = blockIdx.x * PREFIX_CB + ch;
    const bool jok = j < c.mu;
    double s = 0.0;
    if (scanner) {
        if (jok) s = sign ? c.s_neg[j] : c.s_pos[j];
    } else {
        for (int cidx = 0; cidx < PREFIX_NBUF - 1; ++cidx) stage(cidx, cidx * PREFIX_CHK);
        cp_async_wait<PREFIX_NBUF - 2>();
    }
    __syncthreads();
    for (int cidx = 0; cidx * PREFIX_CHK < nslots; ++cidx) {
        const int e0 = cidx * PREFIX_CHK, buf = cidx % PREFIX_NBUF;
        if (!scanner) {
            stage((cidx + PREFIX_NBUF - 1) % PREFIX_NBUF, (cidx + PREFIX_NBUF - 1) * PREFIX_CHK);
            cp_async_wait<PREFIX_NBUF - 2>();                // chunk cidx + 1 has landed
        } else if (jok) {
            const long long left = min((long long)(nslots - e0), c.nsteps - kfirst - e0);
            double* out = sbuf + (size_t(e0) * 2 + sign) * c.mu_pad + j;
            const size_t ostride = size_t(2) * c.mu_pad;
            if (left >= PREFIX_CHK) {                        // full chunk: branch-free, operands in registers first
                double d[PREFIX_CHK];
#pragma unroll
                for (int i = 0; i < PREFIX_CHK; ++i) {
                    const double lin = sv[buf][i][0][ch], ab = sv[buf][i][1][ch];
                    d[i] = sign ? lgg09_out_neg(c.prog, lin, ab) : lgg09_out_pos(c.prog, lin, ab);       // rho(-r_k; V) : rho(+r_k; V)
                }
#pragma unroll
                for (int i = 0; i < PREFIX_CHK; ++i) {
                    out[i * ostride] = s;
                    s += d[i];
                }
            } else {
                for (int i = 0; i < int(left); ++i) {
                    const double lin = sv[buf][i][0][ch], ab = sv[buf][i][1][ch];
                    out[i * ostride] = s;
                    s += sign ? lgg09_out_neg(c.prog, lin, ab) : lgg09_out_pos(c.prog, lin, ab);
                }
            }
        }
        __syncthreads();
    }
    if (scanner && jok) { if (sign) c.s_neg[j] = s; else c.s_pos[j] = s; }
}

// ---- persistent warp-per-chain kernel for small / sparse Phi ---------------------------------------
// When Phi' fits in shared memory in ELL form (dense n <~ 100, or a sparse Phi such as ISS: 540
// non-zeros of 72 900 -- the reference runs that model with `sparse=true`, LGG09/post.jl:25-27),
// one warp walks one direction chain through ALL time steps without ever leaving the SM:
//     r_{k+1} = Phi' r_k  (row i on lane i%32; per row the non-zeros in ascending column order:
//                          the strictly sequential association of the reference's mul!)
//     features of r_{k+1} (lane-strided partial sums + butterfly), alternatives / max, running s,
//     two 8-byte stores per step.
// Cost per chain-step is O(nnz/32 + n*NQ/32) FMAs plus ~10 shuffle rounds -- for ISS ~10^2 cycles
// where the dense GEMM path spends 135x the necessary flops.
struct Lgg09Chain {
    Lgg09Program prog;
    int n, npad;                    // npad = RPL*32 >= n
    int maxnnz;                     // ELL width
    const double* ell_val;          // [maxnnz][npad]  (row i of Phi' = column i of Phi)
    const int* ell_col;             // [maxnnz][npad]
    const double* Wl;               // [nq][wstride] direct weights first (n), mapped weights at offset n
    const double* Wa;
    int wstride, nq, ne;
    const double* E;                // [ne][n] mapped rows
    const double* L0;               // [mu][ldl] chains
    int ldl, mu;
    long long nsteps;
    int need_next;
    const int* row_pos;
    const int* row_neg;
    double* rho;
    int ndirs;
    long long* dbg;                 // optional [8] phase cycle counters of CTA 0 (NULL: off)
};

// One CTA (4 warps) per chain: thread t owns rows t, t+128, ... of r (RPL of them).  Only r_{k+1} = Phi' r_k is truly
// sequential, everything else is instruction count -- the kernel is issue-bound (ncu: ~3400 SASS instructions of flat
// stall profile, two warps per sub-partition), so it works in batches of CHAIN_B = 4 steps and splits the rest:
//   (1) advance r through the batch, one CTA barrier per step (two shared loads, maxnnz FMAs, a store);
//   (2) every thread: feature partials of its rows for all 4 vectors, parked in shared memory [value][thread];
//       96..128 threads then add them up, 64 threads' worth each, in fixed order (deterministic; this replaces a
//       5-round butterfly of 48 doubles per warp, which was a third of all instructions);
//   (3) warp b evaluates step b of the batch (mapped rows, alternatives, rho(+-r; V)); the running input sum is
//       propagated through the batch in step order from the four warps' increments.
// (v1 ran a chain in one warp, v2/v3 one CTA per chain with butterflies and the evaluation replicated in every warp:
//  all ~3000-3400 cycles per step on ISS.)
constexpr int CHAIN_WARPS = 4, CHAIN_THREADS = CHAIN_WARPS * 32, CHAIN_B = 4;
constexpr int CHAIN_RED_LD = CHAIN_THREADS + 2;      // row stride of the partials: == 2 (mod 16) doubles, see reduce_batch

template <int RPL, int NQ, int NE>
__global__ void __launch_bounds__(CHAIN_THREADS) lgg09_chain_kernel(const Lgg09Chain c) {
    constexpr int NV = 2 * NQ + NE;                                      // raw sums per vector: lin_q, abs_q, y_e
    static_assert(2 * CHAIN_B * NV <= CHAIN_THREADS, "two reducer threads per (vector, value)");
    extern __shared__ __align__(16) unsigned char smraw[];
    const int npad = c.npad;                                             // RPL * 128
    double* s_val = reinterpret_cast<double*>(smraw);                   // [maxnnz][npad]
    double* s_a = s_val + size_t(c.maxnnz) * npad;                      // [NQ][npad]
    double* s_b = s_a + NQ * npad;                                      // [NQ][npad]
    double* s_E = s_b + NQ * npad;                                      // [NE][npad]
    double* s_r = s_E + NE * npad;                                      // [CHAIN_B + 1][npad] ring of r vectors
    double* s_red = s_r + (CHAIN_B + 1) * npad;                         // [CHAIN_B * NV][CHAIN_RED_LD]
    double* s_tot = s_red + CHAIN_B * NV * CHAIN_RED_LD;                // [CHAIN_B + 1][NV] totals; [0] = carried r_k
    double* s_dv = s_tot + (CHAIN_B + 1) * NV;                          // [CHAIN_B][2] increments of the running sums
    int* s_col = reinterpret_cast<int*>(s_dv + CHAIN_B * 2);            // [maxnnz][npad]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x;                                            // chain of this CTA

    for (int idx = tid; idx < c.maxnnz * npad; idx += CHAIN_THREADS) { s_val[idx] = c.ell_val[idx]; s_col[idx] = c.ell_col[idx]; }
    for (int idx = tid; idx < NQ * npad; idx += CHAIN_THREADS) {
        const int q = idx / npad, i = idx % npad;
        const bool ok = q < c.nq && i < c.n;
        s_a[idx] = ok ? c.Wl[size_t(q) * c.wstride + i] : 0.0;
        s_b[idx] = ok ? c.Wa[size_t(q) * c.wstride + i] : 0.0;
    }
    for (int idx = tid; idx < NE * npad; idx += CHAIN_THREADS) {
        const int e = idx / npad, i = idx % npad;
        s_E[idx] = (e < c.ne && i < c.n) ? c.E[size_t(e) * c.n + i] : 0.0;
    }
    for (int i = tid; i < npad; i += CHAIN_THREADS) s_r[i] = i < c.n ? c.L0[size_t(j) * c.ldl + i] : 0.0;

    // mapped-row weights of every feature, replicated per thread (small)
    double wlE[NQ][NE > 0 ? NE : 1], waE[NQ][NE > 0 ? NE : 1];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const bool ok = q < c.nq && e < c.ne;
            wlE[q][e] = ok ? c.Wl[size_t(q) * c.wstride + c.n + e] : 0.0;
            waE[q][e] = ok ? c.Wa[size_t(q) * c.wstride + c.n + e] : 0.0;
        }
    __syncthreads();

    // raw sums of the vectors in ring slots slot0+1 .. slot0+nb -> s_tot[1..nb]   (two CTA barriers inside).
    // The loop over the vectors is deliberately NOT unrolled: with everything unrolled the steady-state loop was ~1300
    // SASS instructions (20 KB), beyond the instruction cache of a sub-partition that runs one or two warps -- ncu showed
    // a flat profile of stall_no_inst and every phase ran ~6x slower than its instruction latencies add up to.
    auto reduce_batch = [&](int slot0, int nb) {
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            int sl = slot0 + b + 1;
            if (sl > CHAIN_B) sl -= CHAIN_B + 1;
            const double* r = s_r + sl * npad;
            double v[NV];
#pragma unroll
            for (int x = 0; x < NV; ++x) v[x] = 0.0;
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int i = tid + CHAIN_THREADS * t;
                const double ri = r[i], ai = fabs(ri);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    v[2 * q] = fma(s_a[q * npad + i], ri, v[2 * q]);
                    v[2 * q + 1] = fma(s_b[q * npad + i], ai, v[2 * q + 1]);
                }
#pragma unroll
                for (int e = 0; e < NE; ++e) v[2 * NQ + e] = fma(s_E[e * npad + i], ri, v[2 * NQ + e]);
            }
#pragma unroll
            for (int x = 0; x < NV; ++x) s_red[(b * NV + x) * CHAIN_RED_LD + tid] = v[x];
        }
        __syncthreads();
        {   // thread pair (2o, 2o+1) adds the 128 partials of value o: thread h takes the partials of threads == h (mod 2)
            // in four interleaved fixed-order streams (16 dependent adds each), then a fixed tree.  Address
            // o*130 + 2i + h: the 16 threads of a half-warp hit 16 different 8-byte banks.
            const int o = tid >> 1, h = tid & 1;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (o < nb * NV) {
                const double* src = s_red + o * CHAIN_RED_LD + h;
#pragma unroll 4
                for (int i = 0; i < 64; i += 4) {
                    a0 += src[2 * i];
                    a1 += src[2 * i + 2];
                    a2 += src[2 * i + 4];
                    a3 += src[2 * i + 6];
                }
            }
            const double acc = (a0 + a1) + (a2 + a3);
            const double other = __shfl_xor_sync(0xffffffffu, acc, 1);
            if (h == 0 && o < nb * NV) s_tot[NV + o] = acc + other;
        }
        __syncthreads();
    };
    // features F_q = (lin, abs) of the vector whose raw sums are s_tot[slot]
    auto features = [&](int slot, double (&F)[NQ][2]) {
        const double* v = s_tot + slot * NV;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double lin = v[2 * q], ab = v[2 * q + 1];
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                lin = fma(wlE[q][e], v[2 * NQ + e], lin);
                ab = fma(waE[q][e], fabs(v[2 * NQ + e]), ab);
            }
            F[q][0] = lin;
            F[q][1] = ab;
        }
    };

    const int rp = c.row_pos[j], rn = c.row_neg[j];
    double sp = 0.0, sn = 0.0;
    int base = 0;                                                        // ring slot of r_k at the start of a batch
    reduce_batch(CHAIN_B, 1);                                            // slot0 + 1 wraps to slot 0 = r_0
    if (tid < NV) s_tot[tid] = s_tot[NV + tid];
    __syncthreads();
    long long tph[5] = {0, 0, 0, 0, 0};
    for (long long k0 = 0; k0 < c.nsteps; k0 += CHAIN_B) {
        const int nb = int(min((long long)CHAIN_B, c.nsteps - k0));
        long long tq0 = clock64();
        // (1) r_{k0+1} .. r_{k0+nb}: ring slots base+1 .. base+nb (mod CHAIN_B+1)
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            int sc = base + b, sx = base + b + 1;
            if (sc > CHAIN_B) sc -= CHAIN_B + 1;
            if (sx > CHAIN_B) sx -= CHAIN_B + 1;
            const double* rc = s_r + sc * npad;
            double* rnx = s_r + sx * npad;
            double acc[RPL];
#pragma unroll
            for (int t = 0; t < RPL; ++t) acc[t] = 0.0;
#pragma unroll 4
            for (int z = 0; z < c.maxnnz; ++z) {
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    const int i = tid + CHAIN_THREADS * t;
                    acc[t] = fma(s_val[z * npad + i], rc[s_col[z * npad + i]], acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < RPL; ++t) rnx[tid + CHAIN_THREADS * t] = acc[t];
            __syncthreads();
        }
        long long tq1 = clock64();
        // (2) raw sums of the nb new vectors -> s_tot[1..nb]
        reduce_batch(base, nb);
        long long tq2 = clock64();
        // (3) warp b: step k0 + b from the features of r_{k0+b} (s_tot[b]) and r_{k0+b+1} (s_tot[b+1])
        double vp = 0.0, vn = 0.0;
        if (warp < nb) {
            double Fp[NQ][2], Fc[NQ][2];
            features(warp, Fp);
            features(warp + 1, Fc);
            for (int a = 0; a < c.prog.nalt; ++a) {
                double ap = 0.0, an = 0.0;
                const int q0 = c.prog.q0[a], q1 = c.prog.q1[a];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (q == q0) { ap += lgg09_out_pos(c.prog, Fp[q][0], Fp[q][1]); an += lgg09_out_neg(c.prog, Fp[q][0], Fp[q][1]); }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (q == q1) { ap += lgg09_out_pos(c.prog, Fc[q][0], Fc[q][1]); an += lgg09_out_neg(c.prog, Fc[q][0], Fc[q][1]); }
                if (a == 0) { vp = ap; vn = an; } else { vp = fmax(vp, ap); vn = fmax(vn, an); }
            }
            double dp = 0.0, dn = 0.0;
            const int qv = c.prog.qV;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (q == qv) { dp = lgg09_out_pos(c.prog, Fp[q][0], Fp[q][1]); dn = lgg09_out_neg(c.prog, Fp[q][0], Fp[q][1]); }
            if (lane == 0) { s_dv[warp * 2] = dp; s_dv[warp * 2 + 1] = dn; }
        }
        __syncthreads();
        long long tq3 = clock64();
        {   // running input sum, strictly in step order: s_{k+1} = s_k + rho(r_k; V)
            double spw = sp, snw = sn;
#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                if (b == warp && lane == 0) {
                    const long long k = k0 + b;
                    if (rp >= 0) c.rho[size_t(k) * c.ndirs + rp] = vp + spw;
                    if (rn >= 0) c.rho[size_t(k) * c.ndirs + rn] = vn + snw;
                }
                spw += s_dv[b * 2];
                snw += s_dv[b * 2 + 1];
            }
            sp = spw;
            sn = snw;
        }
        if (tid < NV) s_tot[tid] = s_tot[nb * NV + tid];                // features of r_{k0+nb} carry into the next batch
        base += nb;
        if (base > CHAIN_B) base -= CHAIN_B + 1;
        __syncthreads();                                                 // s_tot, s_dv and the ring slots are reused
        long long tq4 = clock64();
        tph[0] += tq1 - tq0; tph[1] += tq2 - tq1; tph[2] += tq3 - tq2; tph[3] += tq4 - tq3;
    }
    if (c.dbg && blockIdx.x == 0 && tid == 0) { c.dbg[0] = tph[0]; c.dbg[1] = tph[1]; c.dbg[2] = tph[2]; c.dbg[3] = tph[3]; }
}

// ---- decoupled blocks: compact chains, parallel in time ---------------------------------------------
// r_k = (Phi')^k l only ever touches the rows reachable from supp(l) in the sparsity graph of Phi.  When that closure is
// small for every chain (ISS: Phi = expm(A dt) is 135 decoupled 2x2 oscillators, so a box direction lives in 2 rows for
// ever; any tiny system n <= 16), the chain is a G-dimensional recurrence with its own dense G x G matrix -- exact, the
// rows outside the closure are exactly zero and add exact zeros to every sum of the reference.  A lane group of G lanes
// owns one (chain, time block): local Phi' row, feature weights and mapped-row entries in registers, the mat-vec by
// G shuffles + FMAs (ascending column order: the sequential association of the reference's mul!), feature sums by a
// log2(G)-round butterfly.  Time is cut into blocks of sblk = 2^s steps: the block start vectors come from
// (Phi_C')^sblk (s squarings, lgg09_compact_starts_kernel) -- the same reassociation through stored powers as the GEMM
// path's time blocking -- so chains x blocks groups run at once instead of one 33 334-step dependent loop per chain.
// The running input sum s_k needs the totals of the earlier blocks: pass ONLY_V sums rho(+-r_k; V) per block in step
// order, lgg09_compact_scan_kernel turns the totals into block offsets (fixed order), the second pass writes rho once.
struct Lgg09Compact {
    Lgg09Program prog;
    int mu, mu_g;                   // chains; chains padded to whole warps (32/G per warp)
    int nq, ne, nblk, need_next, ndirs;
    long long nsteps, sblk;         // steps per time block (power of two)
    int log2_sblk;
    const double* M;                // [mu][G][G]  M[c][i][l] = Phi'(idx_i, idx_l) = Phi(idx_l, idx_i)
    const double* r0;               // [mu][G]
    const double* wl;               // [mu][NS][G] direct weights of the local rows per evaluation slot (NS = 2 nalt + 1)
    const double* wa;               // [mu][NS][G]
    const double* E;                // [mu][NE][G] mapped rows restricted to the closure
    const double* WlE;              // [NS][ne] mapped-row weights per slot
    const double* WaE;
    unsigned slot_used;             // bit q: slot q has a feature
    double* rstart;                 // [nblk][mu][G]
    double* voff;                   // [nblk][mu][2] block totals of rho(+-r; V), then exclusive offsets
    const int* row_pos;
    const int* row_neg;
    double* rho;
};

template <int G>
__global__ void lgg09_compact_starts_kernel(const Lgg09Compact c) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c.mu) return;
    double P[G * G], T[G * G];
#pragma unroll
    for (int x = 0; x < G * G; ++x) P[x] = c.M[size_t(ch) * G * G + x];
    for (int s = 0; s < c.log2_sblk; ++s) {                              // P <- P * P
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int l = 0; l < G; ++l) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < G; ++m) acc = fma(P[i * G + m], P[m * G + l], acc);
                T[i * G + l] = acc;
            }
#pragma unroll
        for (int x = 0; x < G * G; ++x) P[x] = T[x];
    }
    double r[G], rn[G];
#pragma unroll
    for (int i = 0; i < G; ++i) r[i] = c.r0[size_t(ch) * G + i];
    for (int b = 0; b < c.nblk; ++b) {
#pragma unroll
        for (int i = 0; i < G; ++i) c.rstart[(size_t(b) * c.mu + ch) * G + i] = r[i];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < G; ++l) acc = fma(P[i * G + l], r[l], acc);
            rn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < G; ++i) r[i] = rn[i];
    }
}

static __global__ void lgg09_compact_scan_kernel(double* __restrict__ voff, int mu, int nblk) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;                 // (chain, sign)
    if (x >= 2 * mu) return;
    double s = 0.0;
    for (int b0 = 0; b0 < nblk; b0 += 16) {                              // 16 independent loads, then the ordered adds
        double t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = b0 + i < nblk ? voff[size_t(b0 + i) * 2 * mu + x] : 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (b0 + i < nblk) voff[size_t(b0 + i) * 2 * mu + x] = s;
            s += t[i];
        }
    }
}

// Evaluation slots instead of feature indices: slot 2a is the feature alternative a evaluates at r_k, slot 2a+1 the one
// it evaluates at r_{k+1}, the last slot is rho(.; V).  The host gathers the weights per slot (zeros for an absent
// feature), so the step loop has no runtime feature selection (the first version spent a third of its ~424 warp
// instructions per step on selects) and skips unused slots / mapped rows with warp-uniform branches.
template <int G, int NALT, int NE, bool ONLY_V>
__global__ void __launch_bounds__(128) lgg09_compact_kernel(const Lgg09Compact c) {
    constexpr int NEP = NE > 0 ? NE : 1;
    constexpr int NS = 2 * NALT + 1, SV = 2 * NALT;                      // slots; the V slot
    constexpr int CPW = 32 / G;                                          // chains per warp
    __shared__ double s_wlE[NS][NEP], s_waE[NS][NEP];
    for (int x = threadIdx.x; x < NS * NEP; x += blockDim.x) {
        const int q = x / NEP, e = x % NEP;
        const bool ok = e < c.ne;
        s_wlE[q][e] = ok ? c.WlE[q * c.ne + e] : 0.0;
        s_waE[q][e] = ok ? c.WaE[q * c.ne + e] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, i = lane % G, gbase = lane - i;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int wpb = c.mu_g / CPW;                                        // warps per time block
    const int b = int(warp / wpb);
    if (b >= c.nblk) return;                                             // whole warps only
    const int ch = int(warp % wpb) * CPW + lane / G;
    const bool live = ch < c.mu;
    const int cc = live ? ch : 0;                                        // padding groups replay chain 0, store nothing
    const unsigned used = c.slot_used;

    double Mrow[G];
#pragma unroll
    for (int l = 0; l < G; ++l) Mrow[l] = c.M[(size_t(cc) * G + i) * G + l];
    double wl[NS], wa[NS], Ei[NEP];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        wl[q] = c.wl[(size_t(cc) * NS + q) * G + i];
        wa[q] = c.wa[(size_t(cc) * NS + q) * G + i];
    }
#pragma unroll
    for (int e = 0; e < NEP; ++e) Ei[e] = (NE > 0 && e < c.ne) ? c.E[(size_t(cc) * NE + e) * G + i] : 0.0;
    double r = c.rstart[(size_t(b) * c.mu + cc) * G + i];
    const long long kb = b * c.sblk, ke = min(c.nsteps, kb + c.sblk);
    // outputs of a (lin, abs) pair without selects: pos = lin + kpos*abs, neg = abs + kneg*lin
    //   support values: (abs + lin, abs - lin)      box output: (lin, abs)      -- fma(x, 1, y) = x + y exactly
    const double kpos = c.prog.box ? 0.0 : 1.0, kneg = c.prog.box ? 0.0 : -1.0;

    auto next = [&](double v) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < G; ++l) acc = fma(Mrow[l], __shfl_sync(0xffffffffu, v, gbase + l), acc);
        return acc;
    };
    auto group_sum = [&](double v) {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    };
    // padded mapped rows (e >= ne) have zero entries and zero weights: they add exact zeros, no guards needed
    auto mapped = [&](double v, double (&y)[NEP]) {
#pragma unroll
        for (int e = 0; e < NE; ++e) y[e] = group_sum(Ei[e] * v);
    };
    // (lin, abs) of slot q at the vector v with mapped rows y
    auto slot = [&](int q, double v, double av, const double (&y)[NEP], double& lin, double& ab) {
        lin = group_sum(wl[q] * v);
        ab = group_sum(wa[q] * av);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            lin = fma(s_wlE[q][e], y[e], lin);
            ab = fma(s_waE[q][e], fabs(y[e]), ab);
        }
    };

    double y[NEP];
#pragma unroll
    for (int e = 0; e < NEP; ++e) y[e] = 0.0;
    if (ONLY_V) {
        // block totals of rho(+r_k; V), rho(-r_k; V), summed in step order
        double tp = 0.0, tn = 0.0;
        for (long long k = kb; k < ke; ++k) {
            double lin, ab;
            mapped(r, y);
            slot(SV, r, fabs(r), y, lin, ab);
            tp += fma(ab, kpos, lin);
            tn += fma(lin, kneg, ab);
            r = next(r);
        }
        if (live && i == 0) {
            c.voff[(size_t(b) * c.mu + ch) * 2] = tp;
            c.voff[(size_t(b) * c.mu + ch) * 2 + 1] = tn;
        }
        return;
    }

    const int rp = live ? c.row_pos[ch] : -1, rn = live ? c.row_neg[ch] : -1;
    const bool has_v = (used >> SV) & 1u;
    bool use0[NALT], use1[NALT];
#pragma unroll
    for (int a = 0; a < NALT; ++a) { use0[a] = (used >> (2 * a)) & 1u; use1[a] = (used >> (2 * a + 1)) & 1u; }
    // output pointers advance by one column (ndirs doubles) per step; lanes that do not store keep a null pointer
    double* out_p = (i == 0 && rp >= 0) ? c.rho + size_t(kb) * c.ndirs + rp : nullptr;
    double* out_n = (i == 0 && rn >= 0) ? c.rho + size_t(kb) * c.ndirs + rn : nullptr;
    const size_t ostride = size_t(c.ndirs);
    double sp = 0.0, sn = 0.0;
    if (has_v) { sp = c.voff[(size_t(b) * c.mu + cc) * 2]; sn = c.voff[(size_t(b) * c.mu + cc) * 2 + 1]; }
    // features of the "current" vector: slot 2a of every alternative and the V slot, carried from step to step
    double c_lin[NALT + 1], c_ab[NALT + 1];
    mapped(r, y);
    {
        const double av = fabs(r);
#pragma unroll
        for (int a = 0; a <= NALT; ++a) {
            const int q = a < NALT ? 2 * a : SV;
            c_lin[a] = c_ab[a] = 0.0;
            if ((used >> q) & 1u) slot(q, r, av, y, c_lin[a], c_ab[a]);
        }
    }
    for (long long k = kb; k < ke; ++k) {
        r = next(r);                                                     // r_{k+1}
        mapped(r, y);
        const double av = fabs(r);
        double vp = 0.0, vn = 0.0;
#pragma unroll
        for (int a = 0; a < NALT; ++a) {
            double ap = 0.0, an = 0.0;
            if (use0[a]) {
                ap += fma(c_ab[a], kpos, c_lin[a]);
                an += fma(c_lin[a], kneg, c_ab[a]);
            }
            if (use1[a]) {                                               // evaluated at r_{k+1}
                double lin, ab;
                slot(2 * a + 1, r, av, y, lin, ab);
                ap += fma(ab, kpos, lin);
                an += fma(lin, kneg, ab);
            }
            if (a == 0) { vp = ap; vn = an; } else { vp = fmax(vp, ap); vn = fmax(vn, an); }
        }
        if (out_p) { *out_p = vp + sp; out_p += ostride; }
        if (out_n) { *out_n = vn + sn; out_n += ostride; }
        if (has_v) {                                                     // s_{k+1} = s_k + rho(r_k; V)
            sp += fma(c_ab[NALT], kpos, c_lin[NALT]);
            sn += fma(c_lin[NALT], kneg, c_ab[NALT]);
        }
        // the current-vector slots of the next step
#pragma unroll
        for (int a = 0; a < NALT; ++a)
            if (use0[a]) slot(2 * a, r, av, y, c_lin[a], c_ab[a]);
        if (has_v) slot(SV, r, av, y, c_lin[NALT], c_ab[NALT]);
    }
}

// stack block 0 starts with the identity: out[i][i] = 1 (the buffer is zero-filled)
static __global__ void set_identity_kernel(double* __restrict__ out, int ld, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[size_t(i) * ld + i] = 1.0;
}
// L0[c][:] = sign_c * dirs[:, src_c] + 0.0  (canonical orientation; -0.0 normalised to +0.0)
static __global__ void build_chains_kernel(const double* __restrict__ dirs, int n, const int* __restrict__ src,
                                           const double* __restrict__ sign, double* __restrict__ L0, int ld) {
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += blockDim.x) L0[size_t(c) * ld + i] = sign[c] * dirs[size_t(src[c]) * n + i] + 0.0;
}

// rho(d_j, flowpipe) = max_k rho[j,k]  (Flowpipes/AbstractFlowpipe.jl:35-37); first maximiser.
static __global__ void max_over_steps_kernel(const double* __restrict__ rho, int ndirs, long long nsteps,
                                      double* __restrict__ vmax, long long* __restrict__ amax) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ndirs) return;
    double best = -INFINITY;
    long long at = -1;
    for (long long k = 0; k < nsteps; ++k) {
        const double v = rho[size_t(k) * ndirs + j];
        if (v > best || at < 0) { best = v; at = k; }
    }
    vmax[j] = best;
    if (amax) amax[j] = at;
}

// Early termination (reach_homog.jl:418, reach_inhomog.jl:210): first step whose template reach-set
// P_k = {x : d_j.x <= rho_j[k]} is certified disjoint from the invariant.  A face {a.x <= b} of the invariant is
// disjoint from P_k iff min_{x in P_k} a.x > b; every lambda >= 0 with sum_j lambda_j d_j = -a gives the lower bound
// min a.x >= -sum_j lambda_j rho_j[k] (weak duality), so "sum_j lambda_j rho_j[k] < -b" certifies it.  The host picks
// lambda: one template row (-a is a positive multiple of it) or the +-e_i rows of a box template -- both exact there,
// because the support values are attained (rho_j[k] = rho(d_j, Omega_k) and Omega_k is inside P_k).
// Checks are stored in CSR form: check i = rows/weights [ptr[i], ptr[i+1]), bound[i] = -b.
struct InvChecks {
    const int* ptr;
    const int* rows;
    const double* w;
    const double* bound;
    int n;
};
static __global__ void first_disjoint_kernel(const double* __restrict__ rho, int ndirs, long long k0, long long k1,
                                             const InvChecks c, unsigned long long* __restrict__ first) {
    const long long k = k0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const double* col = rho + size_t(k) * ndirs;
    bool bad = false;
    for (int i = 0; i < c.n && !bad; ++i) {
        double acc = 0.0;
        for (int e = c.ptr[i]; e < c.ptr[i + 1]; ++e) acc = fma(c.w[e], col[c.rows[e]], acc);
        bad = acc < c.bound[i];
    }
    if (bad) atomicMin(first, (unsigned long long)k);
}

}  // namespace reach
