// glgm06_kernels.cuh -- device side of the GLGM06 zonotope flowpipe.
//
// Reference loop (src/Algorithms/GLGM06/reach_inhomog.jl:28-41), per initial set:
//     R_k = reduce_order(P Omega0 (+) W, r, GIR05);  W = reduce_order(P U (+) W, r, GIR05);  P = P Phi
// with linear_map_minkowski_sum (Continuous/setops.jl:7-25): c = P c1 + c2, G = [P G1 | G2].
//
// Device formulation.  P_k and W_k do not depend on the initial set, so the run is split in
//   (1) the shared chain, sequential in k and small:  P_{k+1} = P_k Phi,  P_k [G_U | c_U],
//       GIR05 of [P_k G_U | W_k]  -> W_{k+1}, plus, per step, the sort of W_k's generator weights
//       and the prefix table of |W_k| in that order;
//   (2) the batched part, parallel over (split, step): ONE nt_dgemm over the stack of all powers
//       gives P_k [G0 | c0] for every split and step (generators are the rows of the X operand),
//       its epilogue also emits each generator's sum|.| and max|.|; a per-(split,step) kernel
//       ranks the split's own generators, merges them with the shared sorted list to find the
//       GIR05 cut, and writes the selection, the box diagonal and the centre.
// A reduced zonotope is stored as [own mapped generators | selection | diagonal] + a reference to
// the step's shared generators, never as a dense n x (r n) matrix (device-resident Flowpipe).
#pragma once

#include "nt_dgemm.cuh"

namespace reach {

// ---- GEMM epilogue: store C with a per-N-block stride and emit per-row sum|c| / max|c| ----------
struct GenEpi {
    double* out;            // element (x, blk*nb + i) -> out[blk*blk_stride + x*ldc + i], i < n
    long long blk_stride;
    int ldc;
    int n, nb;              // valid / padded columns per block
    int m_valid;
    double* psum;           // [nblk*tpb][m_pad] partial sum |c| over the tile's valid columns (may be NULL)
    double* pmax;           // [nblk*tpb][m_pad] partial max |c|
    int m_pad;

    template <int MT, int NT, int WN, class TC>
    __device__ __forceinline__ void run(double (&acc)[MT][NT][2], const TC& tc, double* smem) const {
        constexpr int TM = TC::TM, TN = TC::TN;
        const int col0 = tc.tn * TN;
        const int blk = col0 / nb;
        const int cib0 = col0 % nb;
        double* o_blk = out + size_t(blk) * blk_stride;
        double* red = smem;                              // [WN][2][TM]
        const int wn = tc.wn0 / (NT * 8);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int ml = tc.wm0 + i * 8 + tc.g4;
            const int m = tc.tm * TM + ml;
            double s = 0.0, mx = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int c = cib0 + tc.wn0 + j * 8 + 2 * tc.t4;
                const double v0 = acc[i][j][0], v1 = acc[i][j][1];
                if (m < m_valid) {
                    double* o = o_blk + size_t(m) * ldc + c;
                    if (c + 1 < n && (ldc & 1) == 0) *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
                    else {
                        if (c < n) o[0] = v0;
                        if (c + 1 < n) o[1] = v1;
                    }
                }
                if (c < n) { s += fabs(v0); mx = fmax(mx, fabs(v0)); }
                if (c + 1 < n) { s += fabs(v1); mx = fmax(mx, fabs(v1)); }
            }
            if (psum) {
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                if (tc.t4 == 0) {
                    red[(wn * 2 + 0) * TM + ml] = s;
                    red[(wn * 2 + 1) * TM + ml] = mx;
                }
            }
        }
        if (!psum) return;
        tc.sync();
        const int tpb = nb / TN;
        const size_t prow = (size_t(blk) * tpb + cib0 / TN) * m_pad + size_t(tc.tm) * TM;
        for (int ml = tc.tid; ml < TM; ml += tc.nthr) {
            double s = 0.0, mx = 0.0;
#pragma unroll
            for (int w = 0; w < WN; ++w) {
                s += red[(w * 2 + 0) * TM + ml];
                mx = fmax(mx, red[(w * 2 + 1) * TM + ml]);
            }
            psum[prow + ml] = s;
            pmax[prow + ml] = mx;
        }
    }
};

// ---- sorting: (weight, index) ascending, i.e. the stable order Julia's sortperm guarantees ------
struct WKey {
    double w;
    int idx;
};
// bitwise, not short-circuit: the compiler turned `||` / `&&` into divergent branches with reconvergence barriers, which
// cost ~150 cycles per comparison in the rank / sort loops of the chain kernel
__device__ __forceinline__ bool wkey_less(const WKey& a, const WKey& b) {
    return bool(int(a.w < b.w) | (int(a.w == b.w) & int(a.idx < b.idx)));
}
// bitonic sort of `np2` (power of two) keys in shared memory by all threads of the CTA
__device__ inline void bitonic_sort(WKey* keys, int np2) {
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    WKey a = keys[i], b = keys[ixj];
                    if (wkey_less(b, a) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}
// The same sort for np2 <= 64 keys by ONE warp: two keys per lane in registers (element e lives in lane e & 31, slot
// e >> 5), compare-exchange partners by shuffle (j < 32) or in-lane (j = 32): ~20 stages of a few shuffles instead of
// ~20 CTA-wide barriers.  (weight, index) keys are distinct, so the result is the unique sorted order.
__device__ inline void bitonic_sort_warp64(WKey* keys, int np2) {
    const int lane = threadIdx.x & 31;
    WKey v[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int e = lane + 32 * s;
        if (e < np2) v[s] = keys[e]; else { v[s].w = INFINITY; v[s].idx = 0x7fffffff; }
    }
    for (int k = 2; k <= 64; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {
                // partner of element e is e ^ 32: the other slot of the same lane; (e & k) == 0 for k = 64: ascending
                if (wkey_less(v[1], v[0])) { WKey t = v[0]; v[0] = v[1]; v[1] = t; }
            } else {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int e = lane + 32 * s;
                    WKey o;
                    o.w = __shfl_xor_sync(0xffffffffu, v[s].w, j);
                    o.idx = __shfl_xor_sync(0xffffffffu, v[s].idx, j);
                    const bool up = (e & k) == 0;
                    const bool lower = (e & j) == 0;               // this element is the lower index of the pair
                    // ascending pair: lower keeps the smaller; descending pair: lower keeps the larger
                    const bool take_min = (lower == up);
                    const bool o_less = wkey_less(o, v[s]);
                    if (take_min ? o_less : !o_less) v[s] = o;
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int e = lane + 32 * s;
        if (e < np2) keys[e] = v[s];
    }
}
__device__ __forceinline__ int lower_bound_d(const double* a, int n, double v) {   // #{a_i < v}
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int upper_bound_key(const WKey* a, int n, double v) {   // #{a_i.w <= v}
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid].w <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

// ---- shared chain, per step: weights + stable sort of the shared generators W_k, prefix of |W_k| --
// GW: [pW][ldg] one generator per row.  Outputs: wB[pW] sorted weights, permB[pW] (sorted -> original),
// bpref[(pW+1)][n]: bpref[t][i] = sum over the first t generators in sorted order of |g[i]|.
// Single CTA; dynamic shared memory: np2 WKeys.
static __global__ void glgm06_shared_prep_kernel(const double* __restrict__ GW, int ldg, int pW, int n, int np2,
                                          double* __restrict__ wB, int* __restrict__ permB,
                                          double* __restrict__ bpref) {
    extern __shared__ __align__(16) unsigned char smraw[];
    WKey* keys = reinterpret_cast<WKey*>(smraw);
    for (int j = threadIdx.x; j < np2; j += blockDim.x) {
        if (j < pW) {
            double s = 0.0, mx = 0.0;
            for (int i = 0; i < n; ++i) {
                const double a = fabs(GW[size_t(j) * ldg + i]);
                s += a;
                mx = fmax(mx, a);
            }
            keys[j].w = s - mx;
            keys[j].idx = j;
        } else {
            keys[j].w = INFINITY;
            keys[j].idx = 0x7fffffff;
        }
    }
    __syncthreads();
    bitonic_sort(keys, np2);
    for (int j = threadIdx.x; j < pW; j += blockDim.x) {
        wB[j] = keys[j].w;
        permB[j] = keys[j].idx;
    }
    // prefix table, one thread per coordinate i (sequential in sorted order)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        bpref[i] = 0.0;
        for (int t = 0; t < pW; ++t) {
            s += fabs(GW[size_t(keys[t].idx) * ldg + i]);
            bpref[size_t(t + 1) * n + i] = s;
        }
    }
}

// ---- GIR05 of [own | shared] given the shared side's sorted weights ------------------------------
// own: pa generators (rows of `A`, leading dimension lda) with weights computed here or supplied;
// shared: pb generators already sorted (wB, permB, bpref).  kept = floor(n (r-1)) smallest-weight
// generators are NOT boxed; m = pa + pb - kept are.  Results:
//   pos_own[j]  output column of own generator j, or -1 if it went into the box
//   *tb         number of boxed shared generators (the first tb in sorted order)
//   diag[i]     sum over boxed generators of |g[i]|
// Cooperative over one CTA; `keys` is shared memory for np2 >= pa keys; `sw` for pa doubles.
__device__ inline void gir05_cut(const double* __restrict__ A, int lda, int pa, int n, WKey* keys, double* sw,
                                 int np2, const double* __restrict__ wB, int pb, const double* __restrict__ bpref,
                                 int kept, int* __restrict__ pos_own, int* tb_out, double* __restrict__ diag,
                                 int* s_ta) {
    const int m = pa + pb - kept;                      // > 0 by construction
    bitonic_sort(keys, np2);
    for (int i = threadIdx.x; i < pa; i += blockDim.x) sw[i] = keys[i].w;
    if (threadIdx.x == 0) *s_ta = 0;
    __syncthreads();
    // own generator at sorted position i has union rank i + #{shared with weight < a_i} (ties: own first)
    for (int i = threadIdx.x; i < pa; i += blockDim.x) {
        const int lb = lower_bound_d(wB, pb, sw[i]);
        if (i + lb < m) atomicMax(s_ta, i + 1);        // boxed; monotone in i, so ta = 1 + last boxed
    }
    __syncthreads();
    const int ta = *s_ta;
    const int tb = m - ta;
    for (int i = threadIdx.x; i < pa; i += blockDim.x) {
        int pos = -1;
        if (i >= ta) {
            const int lb = lower_bound_d(wB, pb, sw[i]);
            pos = (i - ta) + max(0, lb - tb);
        }
        pos_own[keys[i].idx] = pos;
    }
    if (threadIdx.x == 0) *tb_out = tb;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < ta; ++t) s += fabs(A[size_t(keys[t].idx) * lda + i]);   // own, sorted order
        diag[i] = s + (pb > 0 ? bpref[size_t(tb) * n + i] : 0.0);
    }
    __syncthreads();
}

// ---- batched per-(split, step) selection ----------------------------------------------------------
struct SelectParams {
    const double* A;        // [nblk][m_pad_rows][lda]: row = split*(p0+1) + g ; g == p0 is P c0
    long long a_blk_stride;
    int lda;
    const double* psum;     // [nblk*tpb][m_pad]
    const double* pmax;
    int tpb, m_pad;
    int n, p0, nsplits, nblk, r;
    const int* pW;          // [nblk] shared generator count per step
    const double* wB;       // [nblk][pWcap]
    const double* bpref;    // [nblk][(pWcap+1)][n]
    const double* cW;       // [nblk][n] centre of W at that step
    int pWcap;
    int* pos_own;           // [nblk][nsplits][p0]
    int* tb;                // [nblk][nsplits]  (-1: no reduction at this step)
    double* diag;           // [nblk][nsplits][n]
    double* centre;         // [nblk][nsplits][n]
};

static __global__ void glgm06_select_kernel(const SelectParams s, int np2) {
    extern __shared__ __align__(16) unsigned char smraw[];
    WKey* keys = reinterpret_cast<WKey*>(smraw);
    double* sw = reinterpret_cast<double*>(keys + np2);
    __shared__ int s_ta;
    const int split = blockIdx.x, b = blockIdx.y;
    const int row0 = split * (s.p0 + 1);
    const double* A = s.A + size_t(b) * s.a_blk_stride + size_t(row0) * s.lda;
    const size_t o = size_t(b) * s.nsplits + split;
    const int pb = s.pW[b];
    // centre: P c0 + c_W
    for (int i = threadIdx.x; i < s.n; i += blockDim.x)
        s.centre[o * s.n + i] = A[size_t(s.p0) * s.lda + i] + s.cW[size_t(b) * s.n + i];
    const int p = s.p0 + pb;
    if (p <= s.r * s.n) {                               // r n >= p: returned unchanged
        for (int j = threadIdx.x; j < s.p0; j += blockDim.x) s.pos_own[o * s.p0 + j] = j;
        if (threadIdx.x == 0) s.tb[o] = -1;
        return;
    }
    for (int j = threadIdx.x; j < np2; j += blockDim.x) {
        if (j < s.p0) {
            double sum = 0.0, mx = 0.0;
            for (int t = 0; t < s.tpb; ++t) {
                sum += s.psum[(size_t(b) * s.tpb + t) * s.m_pad + row0 + j];
                mx = fmax(mx, s.pmax[(size_t(b) * s.tpb + t) * s.m_pad + row0 + j]);
            }
            keys[j].w = sum - mx;
            keys[j].idx = j;
        } else {
            keys[j].w = INFINITY;
            keys[j].idx = 0x7fffffff;
        }
    }
    __syncthreads();
    const int kept = s.n * (s.r - 1);
    gir05_cut(A, s.lda, s.p0, s.n, keys, sw, np2, s.wB + size_t(b) * s.pWcap, pb,
              s.bpref + size_t(b) * (s.pWcap + 1) * s.n, kept, s.pos_own + o * s.p0, s.tb + o,
              s.diag + o * s.n, &s_ta);
}

// The same selection with ONE WARP per (split, step) for up to 32 * KPL own generators: keys live in registers (element
// e = 32 s + lane in slot s), the bitonic network exchanges by shuffle (j < 32) or between slots of a lane (j >= 32), the
// cut is a ballot count, the box diagonal is accumulated by the lanes over the coordinates with the sorted indices
// broadcast by shuffle.  No shared memory and no barrier; ncu of the CTA version: 1567 instructions per warp x 4 warps
// per (split, step), 28 barriers, half of it the sort.  Same arithmetic and the same order of additions.
__device__ __forceinline__ int lower_bound_g(const double* __restrict__ a, int n, double v) {   // #{a_i < v}, global array
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid; }
    return lo;
}
template <int KPL>
static __global__ void __launch_bounds__(128) glgm06_select_warp_kernel(const SelectParams s, int nbh) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (long long)s.nsplits * nbh) return;
    const int split = int(w % s.nsplits), b = int(w / s.nsplits);
    const int row0 = split * (s.p0 + 1);
    const double* A = s.A + size_t(b) * s.a_blk_stride + size_t(row0) * s.lda;
    const size_t o = size_t(b) * s.nsplits + split;
    const int pb = s.pW[b], n = s.n, pa = s.p0;
    for (int i = lane; i < n; i += 32) s.centre[o * n + i] = A[size_t(pa) * s.lda + i] + s.cW[size_t(b) * n + i];
    if (pa + pb <= s.r * n) {                           // r n >= p: returned unchanged
        for (int j = lane; j < pa; j += 32) s.pos_own[o * pa + j] = j;
        if (lane == 0) s.tb[o] = -1;
        return;
    }
    WKey key[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int j = 32 * q + lane;
        if (j < pa) {
            double sum = 0.0, mx = 0.0;
            for (int t = 0; t < s.tpb; ++t) {
                sum += s.psum[(size_t(b) * s.tpb + t) * s.m_pad + row0 + j];
                mx = fmax(mx, s.pmax[(size_t(b) * s.tpb + t) * s.m_pad + row0 + j]);
            }
            key[q].w = sum - mx;
            key[q].idx = j;
        } else {
            key[q].w = INFINITY;
            key[q].idx = 0x7fffffff;
        }
    }
    // bitonic sort of the 32 KPL keys, ascending by (weight, index)
#pragma unroll
    for (int k = 2; k <= 32 * KPL; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int dq = j >> 5;
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    if ((q & dq) == 0) {
                        const bool up = ((32 * q + lane) & k) == 0;
                        const bool sw2 = wkey_less(key[q + dq], key[q]) == up;
                        const WKey lo = sw2 ? key[q + dq] : key[q], hi = sw2 ? key[q] : key[q + dq];
                        key[q] = lo;
                        key[q + dq] = hi;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    const int e = 32 * q + lane;
                    WKey other;
                    other.w = __shfl_xor_sync(0xffffffffu, key[q].w, j);
                    other.idx = __shfl_xor_sync(0xffffffffu, key[q].idx, j);
                    const bool take_min = (((e & j) == 0) == ((e & k) == 0));
                    const bool o_less = wkey_less(other, key[q]);
                    if (take_min ? o_less : !o_less) key[q] = other;
                }
            }
        }
    }
    // the cut: own generator at sorted position e has union rank e + #{shared with smaller weight} (ties: own first)
    const double* wB = s.wB + size_t(b) * s.pWcap;
    const int kept = n * (s.r - 1);
    const int m = pa + pb - kept;
    int lb[KPL];
    int ta = 0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int e = 32 * q + lane;
        lb[q] = e < pa ? lower_bound_g(wB, pb, key[q].w) : pb;
        ta += __popc(__ballot_sync(0xffffffffu, e < pa && e + lb[q] < m));      // boxed; monotone in e
    }
    const int tb = m - ta;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int e = 32 * q + lane;
        if (e < pa) s.pos_own[o * pa + key[q].idx] = e >= ta ? (e - ta) + max(0, lb[q] - tb) : -1;
    }
    if (lane == 0) s.tb[o] = tb;
    // box diagonal: boxed own generators in sorted order, then the prefix-table row of the tb boxed shared ones
    const double* bpref = s.bpref + size_t(b) * (s.pWcap + 1) * n;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int cnt = min(32, ta - 32 * q);
            for (int l = 0; l < cnt; ++l) {
                const int idx = __shfl_sync(0xffffffffu, key[q].idx, l);
                if (i < n) acc += fabs(A[size_t(idx) * s.lda + i]);
            }
        }
        if (i < n) s.diag[o * n + i] = acc + (pb > 0 ? bpref[size_t(tb) * n + i] : 0.0);
    }
}

// ---- shared chain: W_{k+1} = reduce_order([P G_U | W_k]) materialised -----------------------------
// PU: [(pU+1)][ldp] rows = P g_U ..., last row = P c_U.  GW/cW in, GWn/cWn out.  pWn = new count.
static __global__ void glgm06_w_reduce_kernel(const double* __restrict__ PU, int ldp, int pU,
                                       const double* __restrict__ GW, int ldg, int pW,
                                       const double* __restrict__ cW, const double* __restrict__ wB,
                                       const int* __restrict__ permB, const double* __restrict__ bpref,
                                       int n, int r, int np2, double* __restrict__ GWn, double* __restrict__ cWn,
                                       int* __restrict__ scratch_pos) {
    extern __shared__ __align__(16) unsigned char smraw[];
    WKey* keys = reinterpret_cast<WKey*>(smraw);
    double* sw = reinterpret_cast<double*>(keys + np2);
    double* sdiag = sw + np2;
    __shared__ int s_ta, s_tb;
    for (int i = threadIdx.x; i < n; i += blockDim.x) cWn[i] = PU[size_t(pU) * ldp + i] + cW[i];
    const int p = pU + pW;
    if (p <= r * n) {                                   // unchanged: [P G_U | W]
        for (int idx = threadIdx.x; idx < p * n; idx += blockDim.x) {
            const int g = idx / n, i = idx % n;
            GWn[size_t(g) * ldg + i] = g < pU ? PU[size_t(g) * ldp + i] : GW[size_t(g - pU) * ldg + i];
        }
        return;
    }
    for (int j = threadIdx.x; j < np2; j += blockDim.x) {
        if (j < pU) {
            double sum = 0.0, mx = 0.0;
            for (int i = 0; i < n; ++i) {
                const double a = fabs(PU[size_t(j) * ldp + i]);
                sum += a;
                mx = fmax(mx, a);
            }
            keys[j].w = sum - mx;
            keys[j].idx = j;
        } else {
            keys[j].w = INFINITY;
            keys[j].idx = 0x7fffffff;
        }
    }
    __syncthreads();
    const int kept = n * (r - 1);
    gir05_cut(PU, ldp, pU, n, keys, sw, np2, wB, pW, bpref, kept, scratch_pos, &s_tb, sdiag, &s_ta);
    __syncthreads();
    const int ta = s_ta, tb = s_tb;
    // kept own generators
    for (int j = threadIdx.x; j < pU; j += blockDim.x) {
        const int pos = scratch_pos[j];
        if (pos >= 0)
            for (int i = 0; i < n; ++i) GWn[size_t(pos) * ldg + i] = PU[size_t(j) * ldp + i];
    }
    // kept shared generators: sorted position j >= tb lands after the kept own ones with weight <= w_j
    for (int j = tb + threadIdx.x; j < pW; j += blockDim.x) {
        const int ub = upper_bound_key(keys, pU, wB[j]);
        const int pos = (j - tb) + max(0, ub - ta);
        const double* src = GW + size_t(permB[j]) * ldg;
        for (int i = 0; i < n; ++i) GWn[size_t(pos) * ldg + i] = src[i];
    }
    // the box: n axis-aligned generators
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int g = idx / n, i = idx % n;
        GWn[size_t(kept + g) * ldg + i] = (g == i) ? sdiag[i] : 0.0;
    }
}

// ---- shared chain, fused: ONE persistent CTA walks all steps (small n) -----------------------------
// For n <= 64 (Building: n = 48) the whole per-step state of the shared chain -- P_k, Phi, [G_U|c_U],
// the sorted weights of W_k -- fits in shared memory, and one step is ~10^5 FMAs: launching four
// kernels per step costs far more than the arithmetic.  This kernel keeps P_k and Phi on chip for the
// entire run and performs, per step: store P_k; P_k [G_U|c_U]; GIR05 cut of [P_k G_U | W_k]; W_{k+1}
// materialised (global, L2-resident); the sorted order of W_{k+1} WITHOUT sorting (kept generators are
// already ascending, the n box generators have weight exactly 0 and slot in after the zero-weight kept
// ones -- the stable (weight, index) order); the prefix table of |W_{k+1}|; P_{k+1} = P_k Phi.
// All powers P_b = Phi^{b+1} (rows), sequentially P_{b+1} = P_b Phi as `mul!(cache, P, Phi)` does
// (reach_inhomog.jl:38), by one CTA with P_b and Phi' in shared memory.  Running this first lets the
// batched split GEMM (which only needs the powers) overlap with the W chain on a second stream.
static __global__ void __launch_bounds__(1024, 1) glgm06_powers_kernel(const double* __restrict__ P0, const double* __restrict__ PhiT,
                                                                       int n, int ldk, int nb, long long nblk,
                                                                       double* __restrict__ Pstack, int* __restrict__ progress) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int ldp = n + 1;
    double* sP[2];
    sP[0] = reinterpret_cast<double*>(smraw);
    sP[1] = sP[0] + n * ldp;
    double* sPhiT = sP[1] + n * ldp;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int idx = tid; idx < n * n; idx += nthr) {
        const int a = idx / n, b = idx % n;
        sP[0][a * ldp + b] = P0[size_t(a) * ldk + b];
        sPhiT[a * ldp + b] = PhiT[size_t(a) * ldk + b];
    }
    __syncthreads();
    for (long long b = 0; b < nblk; ++b) {
        const double* P = sP[b & 1];
        double* Pn = sP[(b + 1) & 1];
        double* Pg = Pstack + size_t(b) * nb * ldk;
        for (int idx = tid; idx < n * n; idx += nthr) {
            const int i = idx / n, m = idx - i * n;
            Pg[size_t(i) * ldk + m] = P[i * ldp + m];
        }
        if (b + 1 < nblk) {
            // four rows of P_{b+1} per thread: one (conflict-free) load of Phi'(m, l) feeds four FMAs, the P_b entries
            // are warp-wide broadcasts -- the plain one-output-per-thread loop was bound by shared-memory wavefronts
            // (3 per FMA).  Every dot product keeps its sequential order over l.
            const int nrb = (n + 3) / 4;
            for (int idx = tid; idx < nrb * n; idx += nthr) {
                const int ib = idx / n, m = idx - ib * n;
                const int i0 = 4 * ib;
                const double* p0 = P + min(i0, n - 1) * ldp;
                const double* p1 = P + min(i0 + 1, n - 1) * ldp;
                const double* p2 = P + min(i0 + 2, n - 1) * ldp;
                const double* p3 = P + min(i0 + 3, n - 1) * ldp;
                const double* ph = sPhiT + m * ldp;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                for (int l = 0; l < n; ++l) {
                    const double f = ph[l];
                    s0 = fma(p0[l], f, s0);
                    s1 = fma(p1[l], f, s1);
                    s2 = fma(p2[l], f, s2);
                    s3 = fma(p3[l], f, s3);
                }
                Pn[i0 * ldp + m] = s0;
                if (i0 + 1 < n) Pn[(i0 + 1) * ldp + m] = s1;
                if (i0 + 2 < n) Pn[(i0 + 2) * ldp + m] = s2;
                if (i0 + 3 < n) Pn[(i0 + 3) * ldp + m] = s3;
            }
        }
        __syncthreads();
        if (tid == 0) {                                   // blocks 0..b are in global memory: publish (release)
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress), "r"(int(b + 1)) : "memory");
        }
    }
}

// The chain CTA is launched BEFORE the powers kernel (so that it owns an SM before the batched GEMM floods the
// device) and consumes the powers as they are published.
__device__ __forceinline__ void glgm06_wait_progress(const int* progress, int need) {
    if (threadIdx.x == 0) {
        int v = 0;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(progress) : "memory");
            if (v < need) {
                __nanosleep(200);
                if (++spins > (1u << 26)) __trap();
            }
        } while (v < need);
    }
    __syncthreads();
}

static __global__ void glgm06_iota_kernel(int* __restrict__ out, size_t count) {
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < count) out[i] = int(i);
}

constexpr int GLGM06_NSEG = 21;      // prefix-table segments per coordinate (48 x 21 = 1008 threads busy at n = 48)

// The powers kernel runs several times faster than the chain, so one poll usually covers many steps: `known` (the
// same value in every thread) caches the last count read and the CTA only polls -- an L2 round trip plus a barrier --
// when it is not enough.
__device__ __forceinline__ void glgm06_wait_progress_cached(const int* progress, int need, int& known, int* s_known) {
    if (known >= need) return;
    if (threadIdx.x == 0) {
        int v = 0;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(progress) : "memory");
            if (v < need) {
                __nanosleep(200);
                if (++spins > (1u << 26)) __trap();
            }
        } while (v < need);
        *s_known = v;
    }
    __syncthreads();
    known = *s_known;
    __syncthreads();                                  // s_known may be rewritten by the next poll
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}

struct ChainParams {
    int n, ldk, nb;             // ldk = Kp (row pitch of Phi/P/XU buffers), nb = rows per P block
    int pU, r, pWcap;
    int pa_max;                 // most generators any consumer of the prefix table brings along: max(p0, pU)
    long long nblk;
    const double* XU;           // [(pU+1)][ldk]
    const double* Pstack;       // [nblk][nb][ldk], produced concurrently by glgm06_powers_kernel
    const int* progress;        // number of blocks of Pstack published so far
    double* GWall;              // [(nblk+1)][pWcap][n]
    double* cWall;              // [(nblk+1)][n]
    const int* pW;              // [nblk+1]
    double* wB;                 // [nblk][pWcap]
    int* permB;                 // [nblk][pWcap]
    double* bpref;              // [nblk][(pWcap+1)][n]
    int* slotB;                 // [(nblk+1)][pWcap] row of the generator pool (= GWall) that holds column r of W_b
    long long* dbg;             // optional [8] phase cycle counters (NULL: off)
};

static __global__ void __launch_bounds__(1024, 1) glgm06_chain_kernel(const ChainParams c) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = c.n, pU = c.pU, ldp = n + 1;
    const int np2 = [&] { int p = 32; while (p < pU) p <<= 1; return p; }();
    // carve shared memory
    double* sP[2];
    sP[0] = reinterpret_cast<double*>(smraw);
    sP[1] = sP[0] + n * ldp;
    double* sPhiT = sP[1] + n * ldp;                  // [n][ldp]: sPhiT[m][l] = Phi(l, m)
    double* sXU = sPhiT + n * ldp;                    // [(pU+1)][ldp]
    double* sPU = sXU + (pU + 1) * ldp;               // [(pU+1)][ldp]
    double* swB[2];
    swB[0] = sPU + (pU + 1) * ldp;                    // sorted weights of the current / next W
    swB[1] = swB[0] + c.pWcap;
    double* swm = swB[1] + c.pWcap;                   // weights of the kept generators in merged order
    double* sw = swm + c.pWcap;                       // sorted own weights [np2]
    double* sdiag = sw + np2;                         // [n]
    double* sseg = sdiag + n;                         // [n][NSEG] segment totals for the prefix table
    constexpr int NSEG = GLGM06_NSEG;
    WKey* keys = reinterpret_cast<WKey*>(sseg + n * NSEG);      // [np2]
    int* sperm[2];
    sperm[0] = reinterpret_cast<int*>(keys + np2);
    sperm[1] = sperm[0] + c.pWcap;
    int* spos = sperm[1] + c.pWcap;                   // [pWcap] destination row of every kept generator
    __shared__ int s_ta, s_z, s_known;
    int* srow = spos + c.pWcap + pU;                  // [pU+1] column of a single-entry row of [G_U|c_U], -1: general
    int* sslot[2];                                    // pool row of every column of the current / next W
    sslot[0] = srow + pU + 1;
    sslot[1] = sslot[0] + c.pWcap;

    const int tid = threadIdx.x, nthr = blockDim.x;
    int known = 0;
    glgm06_wait_progress_cached(c.progress, 1, known, &s_known);
    for (int idx = tid; idx < n * n; idx += nthr) {
        const int a = idx / n, b = idx % n;
        sP[0][a * ldp + b] = __ldcg(c.Pstack + size_t(a) * c.ldk + b);  // P_0
    }
    for (int idx = tid; idx < (pU + 1) * n; idx += nthr) {
        const int g = idx / n, l = idx % n;
        sXU[g * ldp + l] = c.XU[size_t(g) * c.ldk + l];
    }
    __syncthreads();
    // axis-aligned input generators (a box or ball input: all but a few rows of [G_U|c_U]) have one non-zero entry: their
    // image under P_b is a scaled column of P_b.  Skipping exact zeros leaves every fma chain's value unchanged.
    for (int g = tid; g <= pU; g += nthr) {
        int nz = 0, at = -1;
        for (int l = 0; l < n; ++l)
            if (sXU[g * ldp + l] != 0.0) { ++nz; at = l; }
        srow[g] = nz == 1 ? at : -1;
    }
    __syncthreads();

    // sorted order of W_0 = U: a real sort (once)
    int cur = 0;
    {
        const int pW0 = c.pW[0];
        const double* GW = c.GWall;
        // reuse the bitonic sorter in chunks of np2w keys held in the `swm`/`sperm[1]` area is not possible
        // (WKey storage), so sort through global scratch when pW0 > np2: W_0 has pU generators, pU <= np2.
        for (int j = tid; j < np2; j += nthr) {
            if (j < pW0) {
                double s = 0.0, mx = 0.0;
                for (int i = 0; i < n; ++i) { const double a = fabs(GW[size_t(j) * n + i]); s += a; mx = fmax(mx, a); }
                keys[j].w = s - mx; keys[j].idx = j;
            } else { keys[j].w = INFINITY; keys[j].idx = 0x7fffffff; }
        }
        __syncthreads();
        if (np2 <= 64) { if (tid < 32) bitonic_sort_warp64(keys, np2); __syncthreads(); }
        else bitonic_sort(keys, np2);
        for (int j = tid; j < pW0; j += nthr) { swB[0][j] = keys[j].w; sperm[0][j] = keys[j].idx; }
        __syncthreads();
    }

    // W_0 = U sits in rows 0.. of the pool.  From here on a kept generator is never moved: W_{b+1} lists the pool rows of
    // its columns (slotB), only the <= pU + n new generators of a step are written (into that step's own rows).
    for (int r = tid; r < c.pW[0]; r += nthr) { sslot[0][r] = r; c.slotB[r] = r; }
    __syncthreads();
    long long tph[6] = {0, 0, 0, 0, 0, 0};
    for (long long b = 0; b < c.nblk; ++b) {
        const long long tq0 = clock64();
        const int pWb = c.pW[b];
        const double* P = sP[b & 1];
        double* Pn = sP[(b + 1) & 1];
        const double* pool = c.GWall;
        const int* slotc = sslot[cur];
        int* slotn = sslot[1 - cur];
        const int base_n = int(b + 1) * c.pWcap;          // pool row of column 0 of W_{b+1}'s own block
        double* GWn = c.GWall + size_t(b + 1) * c.pWcap * n;
        const double* wBc = swB[cur];
        const int* permc = sperm[cur];
        double* bprefb = c.bpref + size_t(b) * (c.pWcap + 1) * n;
        if (b + 1 < c.nblk) {
            // P_{b+1} is only needed by the next step: start its copy into the idle buffer now (cp.async) and collect
            // it after this step's own work
            glgm06_wait_progress_cached(c.progress, int(b) + 2, known, &s_known);
            const double* Pg = c.Pstack + size_t(b + 1) * c.nb * c.ldk;
            for (int idx = tid; idx < n * n; idx += nthr) {
                const int a = idx / n, l = idx - a * n;
                cp_async8(Pn + a * ldp + l, Pg + size_t(a) * c.ldk + l);
            }
            cp_async_commit();
        }

        // (a) publish P_b, the sorted weights / permutation of W_b, and the prefix table of |W_b|
        {
            for (int j = tid; j < pWb; j += nthr) {
                c.wB[size_t(b) * c.pWcap + j] = wBc[j];
                c.permB[size_t(b) * c.pWcap + j] = permc[j];
            }
            // prefix over the sorted order, two-level: NSEG segments per coordinate.  W_b lives in global memory (L2);
            // the loads go out in batches of eight independent requests (one dependent L2 round trip per element was
            // half of the step time at 480 generators), the additions keep their order.
            // Only row tb of the table is ever read (the box diagonal of a reduction that boxes tb shared generators),
            // and tb <= m = pa + pW_b - n (r - 1) for a consumer with pa own generators: the table stops at the largest
            // m (C4: 167 of 480 rows), which is the L2 traffic of this phase cut by the same factor.
            const int tmax = max(0, min(pWb, c.pa_max + pWb - n * (c.r - 1)));
            const int seglen = (tmax + NSEG - 1) / NSEG;
            for (int idx = tid; idx < n * NSEG; idx += nthr) {
                const int i = idx % n, sg = idx / n;
                double s = 0.0;
                const int t1 = min(tmax, (sg + 1) * seglen);
                for (int t = sg * seglen; t < t1; t += 8) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = t + u < t1 ? fabs(__ldcg(pool + size_t(slotc[permc[t + u]]) * n + i)) : 0.0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (t + u < t1) s += v[u];
                }
                sseg[i * NSEG + sg] = s;
            }
            __syncthreads();
            for (int idx = tid; idx < n * NSEG; idx += nthr) {
                const int i = idx % n, sg = idx / n;
                double s = 0.0;
                for (int q = 0; q < sg; ++q) s += sseg[i * NSEG + q];
                if (sg == 0) bprefb[i] = 0.0;
                const int t1 = min(tmax, (sg + 1) * seglen);
                for (int t = sg * seglen; t < t1; t += 8) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = t + u < t1 ? fabs(__ldcg(pool + size_t(slotc[permc[t + u]]) * n + i)) : 0.0;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (t + u < t1) {
                            s += v[u];
                            bprefb[size_t(t + u + 1) * n + i] = s;
                        }
                }
            }
        }
        if (b + 1 >= c.nblk) break;                      // W_{nblk} and P_{nblk} are never used
        const long long tq1 = clock64();

        // (b) P_b [G_U | c_U]  and  P_{b+1} = P_b Phi
        for (int idx = tid; idx < (pU + 1) * n; idx += nthr) {
            const int g = idx / n, i = idx - g * n;
            const int l1 = srow[g];
            double s = 0.0;
            if (l1 >= 0) s = fma(sXU[g * ldp + l1], P[i * ldp + l1], s);
            else
                for (int l = 0; l < n; ++l) s = fma(sXU[g * ldp + l], P[i * ldp + l], s);
            sPU[g * ldp + i] = s;
        }
        const long long tq2 = clock64();
        cp_async_wait<0>();                               // P_{b+1} (issued at the top of the step) has landed
        __syncthreads();
        const long long tq3 = clock64();
        // centre of W_{b+1}
        for (int i = tid; i < n; i += nthr) c.cWall[size_t(b + 1) * n + i] = sPU[pU * ldp + i] + c.cWall[size_t(b) * n + i];

        // (c) own weights, sorted
        for (int j = tid; j < np2; j += nthr) {
            if (j < pU) {
                double s = 0.0, mx = 0.0;
                for (int i = 0; i < n; ++i) { const double a = fabs(sPU[j * ldp + i]); s += a; mx = fmax(mx, a); }
                keys[j].w = s - mx; keys[j].idx = j;
            } else { keys[j].w = INFINITY; keys[j].idx = 0x7fffffff; }
        }
        __syncthreads();
        if (pU <= nthr) {
            // rank sort: thread j counts the keys below its own (pU broadcast reads, no barriers in between) and drops its
            // key at that rank -- (weight, index) keys are distinct, so the ranks are a permutation.  One barrier pair
            // instead of the ~20 dependent shuffle stages of a bitonic network.
            WKey me{0.0, 0};
            int rank = 0;
            if (tid < pU) {
                me = keys[tid];
                for (int i = 0; i < pU; ++i) rank += wkey_less(keys[i], me) ? 1 : 0;
            }
            __syncthreads();
            if (tid < pU) keys[rank] = me;
            __syncthreads();
        } else {
            bitonic_sort(keys, np2);
        }
        for (int i = tid; i < pU; i += nthr) sw[i] = keys[i].w;
        if (tid == 0) { s_ta = 0; s_z = 0; }
        __syncthreads();

        const long long tq4 = clock64();
        const int p = pU + pWb;
        const bool reduce = p > c.r * n;
        const int kept = reduce ? n * (c.r - 1) : p;
        const int m = p - kept;
        // (d) the cut: own generator at sorted position i has union rank i + #{shared with smaller weight}
        if (reduce)
            for (int i = tid; i < pU; i += nthr)
                if (i + lower_bound_d(wBc, pWb, sw[i]) < m) atomicMax(&s_ta, i + 1);
        __syncthreads();
        const int ta = s_ta, tb = m - ta;
        // (e) W_{b+1}: kept generators in merged (weight, index) order, then the box
        // destination rows first (one binary search per kept generator, all at once), then two flat copies
        for (int i = ta + tid; i < pU; i += nthr) {
            const int pos = (i - ta) + max(0, lower_bound_d(wBc, pWb, sw[i]) - tb);
            swm[pos] = sw[i];
            sperm[1 - cur][pos] = reduce ? pos : keys[i].idx;
            spos[i] = pos;
            if (reduce) slotn[pos] = base_n + pos;                          // new generator, written below
        }
        for (int j = tb + tid; j < pWb; j += nthr) {
            const int pos = (j - tb) + max(0, upper_bound_key(keys, pU, wBc[j]) - ta);
            swm[pos] = wBc[j];
            sperm[1 - cur][pos] = reduce ? pos : pU + permc[j];
            spos[pU + j] = pos;
            if (reduce) slotn[pos] = slotc[permc[j]];                       // kept generator: stays where it is
        }
        __syncthreads();
        if (reduce) {
            // only the kept own generators are written (row-wise, lanes over the coordinates)
            const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
            for (int i = ta + warp; i < pU; i += nwarps) {
                const double* src = sPU + keys[i].idx * ldp;
                double* dst = GWn + size_t(spos[i]) * n;
                for (int q = lane; q < n; q += 32) dst[q] = src[q];
            }
        }
        const long long tq5 = clock64();
        if (reduce) {
            for (int i = tid; i < n; i += nthr) {
                double s = 0.0;
                for (int t = 0; t < ta; ++t) s += fabs(sPU[keys[t].idx * ldp + i]);
                sdiag[i] = s + bprefb[size_t(tb) * n + i];
            }
        }
        __syncthreads();
        if (!reduce) {
            // unreduced W_{b+1} = [P G_U | W_b] keeps its ORIGINAL column order; only the sort order is merged.
            // Rewrite the generators in original order (own first, then shared).
            for (int idx = tid; idx < pU * n; idx += nthr) {
                const int g = idx / n, q = idx - g * n;
                GWn[size_t(g) * n + q] = sPU[g * ldp + q];
            }
            for (int g = tid; g < p; g += nthr) slotn[g] = g < pU ? base_n + g : slotc[g - pU];
            for (int t = tid; t < p; t += nthr) swB[1 - cur][t] = swm[t];
        } else {
            for (int idx = tid; idx < n * n; idx += nthr) {
                const int g = idx / n, i = idx % n;
                GWn[size_t(kept + g) * n + i] = (g == i) ? sdiag[i] : 0.0;
            }
            for (int g = tid; g < n; g += nthr) slotn[kept + g] = base_n + kept + g;
            // stable order of [kept ascending | n zero-weight box generators]:
            //   zero-weight kept (z of them), then the box, then the rest
            for (int t = tid; t < kept; t += nthr)
                if (swm[t] == 0.0) atomicMax(&s_z, t + 1);
            __syncthreads();
            const int z = s_z;
            for (int t = tid; t < kept + n; t += nthr) {
                int src;                                  // column of W_{b+1} at sorted position t
                if (t < z) src = t;
                else if (t < z + n) src = kept + (t - z);
                else src = t - n;
                sperm[1 - cur][t] = src;
                swB[1 - cur][t] = src < kept ? swm[src] : 0.0;
            }
        }
        __syncthreads();
        for (int r = tid; r < c.pW[b + 1]; r += nthr) c.slotB[size_t(b + 1) * c.pWcap + r] = slotn[r];
        cur ^= 1;
        const long long tq6 = clock64();
        tph[0] += tq1 - tq0; tph[1] += tq2 - tq1; tph[2] += tq3 - tq2; tph[3] += tq4 - tq3; tph[4] += tq5 - tq4; tph[5] += tq6 - tq5;
    }
    if (c.dbg && tid == 0)
        for (int i = 0; i < 6; ++i) c.dbg[i] = tph[i];
}

// ---- stand-alone reduce_order(Z, r, GIR05) for a batch of zonotopes (post.jl:26) ------------------
// G: [batch][p][ldg] rows = generators.  Out: Gr [batch][r n][ldg], sel [batch][r n] (source index or -1).
static __global__ void glgm06_reduce_batch_kernel(const double* __restrict__ G, int ldg, int p, int n, int r, int np2,
                                           double* __restrict__ Gr, int* __restrict__ sel, int* __restrict__ scratch_pos) {
    extern __shared__ __align__(16) unsigned char smraw[];
    WKey* keys = reinterpret_cast<WKey*>(smraw);
    double* sw = reinterpret_cast<double*>(keys + np2);
    double* sdiag = sw + np2;
    __shared__ int s_ta, s_tb;
    const int z = blockIdx.x;
    const double* Gz = G + size_t(z) * p * ldg;
    double* Go = Gr + size_t(z) * (r * n) * ldg;
    int* pos = scratch_pos + size_t(z) * p;
    for (int j = threadIdx.x; j < np2; j += blockDim.x) {
        if (j < p) {
            double sum = 0.0, mx = 0.0;
            for (int i = 0; i < n; ++i) {
                const double a = fabs(Gz[size_t(j) * ldg + i]);
                sum += a;
                mx = fmax(mx, a);
            }
            keys[j].w = sum - mx;
            keys[j].idx = j;
        } else {
            keys[j].w = INFINITY;
            keys[j].idx = 0x7fffffff;
        }
    }
    __syncthreads();
    const int kept = n * (r - 1);
    gir05_cut(Gz, ldg, p, n, keys, sw, np2, nullptr, 0, nullptr, kept, pos, &s_tb, sdiag, &s_ta);
    __syncthreads();
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
        const int q = pos[j];
        if (q >= 0) {
            for (int i = 0; i < n; ++i) Go[size_t(q) * ldg + i] = Gz[size_t(j) * ldg + i];
            if (sel) sel[size_t(z) * (r * n) + q] = j;
        }
    }
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int g = idx / n, i = idx % n;
        Go[size_t(kept + g) * ldg + i] = (g == i) ? sdiag[i] : 0.0;
        if (sel && i == 0) sel[size_t(z) * (r * n) + kept + g] = -1;
    }
}

// ---- materialise one reduced zonotope (fetch) -----------------------------------------------------
// Output: Gout [ngens][n] rows = generators (== n x ngens column-major), sel[ngens].
static __global__ void glgm06_materialize_kernel(const double* __restrict__ A, int lda, int p0,
                                          const int* __restrict__ pos_own, int tb,
                                          const double* __restrict__ pool, const int* __restrict__ rowslot, int ldg, int pW,
                                          const int* __restrict__ permB, const double* __restrict__ diag,
                                          int n, int r, double* __restrict__ Gout, int* __restrict__ sel,
                                          int* __restrict__ occupied /* [ngens] zeroed */) {
    // single CTA
    if (tb < 0) {                                       // not reduced: [own | shared] in original order
        const int p = p0 + pW;
        for (int idx = threadIdx.x; idx < p * n; idx += blockDim.x) {
            const int g = idx / n, i = idx % n;
            Gout[size_t(g) * n + i] = g < p0 ? A[size_t(g) * lda + i] : pool[size_t(rowslot[g - p0]) * ldg + i];
            if (i == 0) sel[g] = g;
        }
        return;
    }
    const int kept = n * (r - 1);
    for (int j = threadIdx.x; j < p0; j += blockDim.x) {
        const int q = pos_own[j];
        if (q >= 0) {
            occupied[q] = 1;
            sel[q] = j;
            for (int i = 0; i < n; ++i) Gout[size_t(q) * n + i] = A[size_t(j) * lda + i];
        }
    }
    __syncthreads();
    // kept shared generators fill the free slots in sorted order (thread 0 walks; counts are small)
    if (threadIdx.x == 0) {
        int q = 0;
        for (int j = tb; j < pW; ++j) {
            while (q < kept && occupied[q]) ++q;
            occupied[q] = 2 + j;                        // remember which sorted shared generator
            ++q;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < kept; q += blockDim.x) {
        const int tag = occupied[q];
        if (tag >= 2) {
            const int j = tag - 2;
            const int src = permB[j];
            sel[q] = p0 + src;
            for (int i = 0; i < n; ++i) Gout[size_t(q) * n + i] = pool[size_t(rowslot[src]) * ldg + i];
        }
    }
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int g = idx / n, i = idx % n;
        Gout[size_t(kept + g) * n + i] = (g == i) ? diag[i] : 0.0;
        if (i == 0) sel[kept + g] = -1;
    }
}

// ---- support function of every stored zonotope in one direction -----------------------------------
// Shared part per step: sfx[b][t] = sum over sorted shared generators j >= t of |g_j . d| (t = 0..pW),
// and for unreduced steps the plain total sits at t = 0.
static __global__ void glgm06_shared_dots_kernel(const double* __restrict__ pool, const int* __restrict__ rowslot, int ldg,
                                          const int* __restrict__ pW, const int* __restrict__ permB, int pWcap,
                                          const double* __restrict__ d, int n, double* __restrict__ sfx) {
    // one CTA per step; thread 0 does the suffix scan after a parallel dot phase
    extern __shared__ __align__(16) unsigned char smraw[];
    double* dots = reinterpret_cast<double*>(smraw);
    const int b = blockIdx.x;
    const int pb = pW[b];
    for (int j = threadIdx.x; j < pb; j += blockDim.x) {
        const double* g = pool + size_t(rowslot[size_t(b) * pWcap + permB[size_t(b) * pWcap + j]]) * ldg;
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += g[i] * d[i];
        dots[j] = fabs(s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        sfx[size_t(b) * (pWcap + 1) + pb] = 0.0;
        for (int j = pb - 1; j >= 0; --j) {
            s += dots[j];
            sfx[size_t(b) * (pWcap + 1) + j] = s;
        }
    }
}

struct SupportParams {
    const double* A; long long a_blk_stride; int lda;
    int n, p0, nsplits, nblk;
    const int* pos_own; const int* tb; const double* diag; const double* centre;
    const double* sfx; int pWcap;
    const double* d;
    double* out;            // [nblk+1][nsplits] ; row 0 is step 0 (filled separately)
};
static __global__ void glgm06_support_kernel(const SupportParams s) {
    // one warp per (split, step)
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= s.nsplits * s.nblk) return;
    const int split = warp % s.nsplits, b = warp / s.nsplits;
    const size_t o = size_t(b) * s.nsplits + split;
    const double* A = s.A + size_t(b) * s.a_blk_stride + size_t(split) * (s.p0 + 1) * s.lda;
    const int tb = s.tb[o];
    double acc = 0.0;
    for (int i = lane; i < s.n; i += 32) {
        acc += s.d[i] * s.centre[o * s.n + i];
        if (tb >= 0) acc += fabs(s.d[i]) * s.diag[o * s.n + i];
    }
    for (int j = lane; j < s.p0; j += 32) {
        if (tb < 0 || s.pos_own[o * s.p0 + j] >= 0) {
            double dot = 0.0;
            for (int i = 0; i < s.n; ++i) dot += A[size_t(j) * s.lda + i] * s.d[i];
            acc += fabs(dot);
        }
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) s.out[size_t(b + 1) * s.nsplits + split] = acc + s.sfx[size_t(b) * (s.pWcap + 1) + (tb < 0 ? 0 : tb)];
}

// rho(d, Z) for dense zonotopes stored as rows [z][(p+1)][ld] (generators then centre): step 0 and
// the homogeneous flowpipe.
static __global__ void zono_support_dense_kernel(const double* __restrict__ Z, long long z_stride, int ld, int p, int n,
                                          const double* __restrict__ d, int count, double* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= count) return;
    const double* z = Z + size_t(warp) * z_stride;
    double acc = 0.0;
    for (int j = lane; j <= p; j += 32) {
        double dot = 0.0;
        for (int i = 0; i < n; ++i) dot += z[size_t(j) * ld + i] * d[i];
        acc += j < p ? fabs(dot) : dot;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[warp] = acc;
}

}  // namespace reach
