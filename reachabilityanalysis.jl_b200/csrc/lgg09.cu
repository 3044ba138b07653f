// lgg09.cu -- host side of the LGG09 device path: flattening of the support program into
// feature rows, direction-chain sharing, the (Phi')^s row stack, launch configuration and the
// C ABI entry points declared in include/reach.h.
//
// Replaces the inside of reach_homog_LGG09! (src/Algorithms/LGG09/reach_homog.jl:6-33,49-62) and
// reach_inhomog_LGG09! (reach_inhomog.jl:5-33,49-81); the invariant variants
// (reach_homog.jl:382-424, reach_inhomog.jl:174-218) are served by the same pass plus a
// first-disjoint-step reduction.
#include "reach_common.cuh"
#include "lgg09_kernels.cuh"
#include "oz_gemm.cuh"
#include "crt_gemm.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

using namespace reach;

namespace {

// ---- launch configurations ------------------------------------------------------------------
struct TileCfg {
    int id, TM, TN, occ;     // occ: CTAs that fit one SM (registers / shared memory)
};
//                         id  TM   TN  occ
const TileCfg kCfgs[] = {{0, 128, 128, 1},    // 2x4 warps, 64x32 warp tile  (large blocks)
                         {1, 64, 64, 2},      // 2x2 warps, 32x32 warp tile  (small n / stack build)
                         {2, 16, 128, 2},     // 1x4 warps, 16x32 warp tile  (a handful of directions)
                         {3, 64, 128, 1},     // 2x4 warps, 32x32 warp tile
                         {4, 128, 128, 1},    // INT8 tcgen05 digit-plane GEMM (oz_gemm.cuh), packed blocks
                         {5, 64, 48, 3}};     // 2x2 warps, 32x24 warp tile: stack blocks of <= 48 rows without padding (Building)
constexpr int kNumCfgs = 6;
constexpr int kOzCfg = 4;

template <class Epi>
cudaError_t launch_cfg(int cfg, const GemmOperands& g, const Epi& epi, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_nt_dgemm_ws<2, 4, 8, 4, 5>(g, epi, st);   // warp-specialised, 5-stage bulk-copy ring
        case 100: return launch_nt_dgemm<2, 4, 8, 4, 4>(g, epi, st);    // v1 (cp.async ring), kept for A/B timing
        case 1:
            // K <= 48: three k-tiles in all -- a 3-stage ring holds the whole panel and a third CTA fits beside (61 KB each)
            if (g.K <= 3 * BK) return launch_nt_dgemm<2, 2, 4, 4, 3>(g, epi, st);
            return launch_nt_dgemm<2, 2, 4, 4, 4>(g, epi, st);
        case 2: return launch_nt_dgemm<1, 4, 2, 4, 4>(g, epi, st);
        case 3: return launch_nt_dgemm<2, 4, 4, 4, 4>(g, epi, st);
        case 5:
            if (g.K <= 3 * BK) return launch_nt_dgemm<2, 2, 4, 3, 3>(g, epi, st);
            return launch_nt_dgemm<2, 2, 4, 3, 4>(g, epi, st);
    }
    return cudaErrorInvalidValue;
}

// ---- flattened support program -> features --------------------------------------------------
struct Feature {
    std::vector<double> a, b;          // direct weights on r (length n)
    std::map<int, std::pair<double, double>> rows;   // mapped row -> (lin, abs) weight
    bool empty() const {
        for (double v : a) if (v != 0.0) return false;
        for (double v : b) if (v != 0.0) return false;
        return rows.empty();
    }
    bool operator==(const Feature& o) const { return a == o.a && b == o.b && rows == o.rows; }
};

struct Program {
    int n = 0;
    std::vector<std::vector<double>> erows;          // mapped rows E (each length n)
    std::map<std::string, int> erow_index;
    std::vector<Feature> feats;
    Lgg09Program dev{};
    bool need_next = false;                          // some term is evaluated at r_{k+1}

    int add_row(const double* v) {
        std::vector<double> row(v, v + n);
        for (double& x : row) x += 0.0;              // -0.0 -> +0.0 so equal rows compare equal
        std::string key(reinterpret_cast<const char*>(row.data()), sizeof(double) * n);
        auto it = erow_index.find(key);
        if (it != erow_index.end()) return it->second;
        int id = int(erows.size());
        erows.push_back(std::move(row));
        erow_index.emplace(std::move(key), id);
        return id;
    }
    int add_feature(const Feature& f) {
        for (size_t i = 0; i < feats.size(); ++i)
            if (feats[i] == f) return int(i);
        feats.push_back(f);
        return int(feats.size()) - 1;
    }
};

// contribution of one direction-space vector v (a column of M, a generator, ...) with weights
// (wl, wa): a single-entry vector folds into the direct weights, anything else becomes a mapped row
void add_vector(Program& P, Feature& f, const double* v, double wl, double wa) {
    int nz = 0, at = -1;
    for (int i = 0; i < P.n; ++i)
        if (v[i] != 0.0) { ++nz; at = i; }
    if (nz == 0) return;
    if (nz == 1) {
        f.a[at] += v[at] * wl;
        f.b[at] += std::fabs(v[at]) * wa;
        return;
    }
    int id = P.add_row(v);
    auto& w = f.rows[id];
    w.first += wl;
    w.second += wa;
}

int add_term(reach_ctx* ctx, Program& P, Feature& f, const reach_term& t) {
    const int n = P.n;
    const int d = t.q > 0 ? t.q : n;
    if (t.q < 0) return set_error(ctx, REACH_EINVAL, "term: q < 0");
    if (t.q > 0 && !t.M) return set_error(ctx, REACH_EINVAL, "term: q > 0 but M is NULL");
    if (t.kind == REACH_TERM_BOX) {
        if (t.q == 0) {
            for (int i = 0; i < n; ++i) {
                if (t.c) f.a[i] += t.c[i];
                if (t.r) {
                    if (t.r[i] < 0) return set_error(ctx, REACH_EINVAL, "term: negative radius");
                    f.b[i] += t.r[i];
                }
            }
        } else {
            for (int j = 0; j < d; ++j) {
                const double wl = t.c ? t.c[j] : 0.0, wa = t.r ? t.r[j] : 0.0;
                if (wa < 0) return set_error(ctx, REACH_EINVAL, "term: negative radius");
                if (wl == 0.0 && wa == 0.0) continue;
                add_vector(P, f, t.M + size_t(j) * n, wl, wa);
            }
        }
        return REACH_OK;
    }
    if (t.kind == REACH_TERM_ZONO) {
        if (t.p < 0 || (t.p > 0 && !t.G)) return set_error(ctx, REACH_EINVAL, "term: bad generators");
        if (t.c) {
            if (t.q == 0) {
                for (int i = 0; i < n; ++i) f.a[i] += t.c[i];
            } else {
                for (int j = 0; j < d; ++j)
                    if (t.c[j] != 0.0) add_vector(P, f, t.M + size_t(j) * n, t.c[j], 0.0);
            }
        }
        std::vector<double> v(n);
        for (int g = 0; g < t.p; ++g) {
            const double* col = t.G + size_t(g) * d;
            if (t.q == 0) {
                add_vector(P, f, col, 0.0, 1.0);
            } else {                                   // v = M g
                std::fill(v.begin(), v.end(), 0.0);
                for (int j = 0; j < d; ++j) {
                    const double gj = col[j];
                    if (gj == 0.0) continue;
                    const double* mc = t.M + size_t(j) * n;
                    for (int i = 0; i < n; ++i) v[i] += mc[i] * gj;
                }
                add_vector(P, f, v.data(), 0.0, 1.0);
            }
        }
        return REACH_OK;
    }
    return set_error(ctx, REACH_EINVAL, "term: unknown kind %d", t.kind);
}

int build_program(reach_ctx* ctx, int n, const reach_setexpr* O, const reach_setexpr* V, Program& P) {
    P.n = n;
    if (!O || O->nalt < 1) return set_error(ctx, REACH_EINVAL, "Omega0 must have at least one alternative");
    if (O->nalt > LGG09_MAX_ALTS)
        return set_error(ctx, REACH_EUNSUPPORTED, "Omega0 has %d alternatives (max %d)", O->nalt, LGG09_MAX_ALTS);
    P.dev.nalt = O->nalt;
    P.dev.qV = -1;
    const reach_term* t = O->terms;
    for (int a = 0; a < O->nalt; ++a) {
        Feature f[2];
        for (auto& x : f) { x.a.assign(n, 0.0); x.b.assign(n, 0.0); }
        const int len = O->alt_len ? O->alt_len[a] : 0;
        if (len < 0) return set_error(ctx, REACH_EINVAL, "negative alt_len");
        for (int i = 0; i < len; ++i, ++t) {
            if (t->which != 0 && t->which != 1) return set_error(ctx, REACH_EINVAL, "term: which must be 0 or 1");
            REACH_TRY(add_term(ctx, P, f[t->which], *t));
        }
        P.dev.q0[a] = f[0].empty() ? int8_t(-1) : int8_t(P.add_feature(f[0]));
        P.dev.q1[a] = f[1].empty() ? int8_t(-1) : int8_t(P.add_feature(f[1]));
        if (P.dev.q1[a] >= 0) P.need_next = true;
    }
    if (V) {
        if (V->nalt != 1) return set_error(ctx, REACH_EINVAL, "V must be a single sum of terms (nalt == 1)");
        Feature f;
        f.a.assign(n, 0.0);
        f.b.assign(n, 0.0);
        const int len = V->alt_len ? V->alt_len[0] : 0;
        for (int i = 0; i < len; ++i) {
            if (V->terms[i].which != 0) return set_error(ctx, REACH_EINVAL, "V terms must have which == 0");
            REACH_TRY(add_term(ctx, P, f, V->terms[i]));
        }
        // an all-zero V still is an input: s stays 0, identical to the reference adding zeros
        P.dev.qV = int8_t(P.add_feature(f));
    }
    if (int(P.feats.size()) > LGG09_MAX_FEATS)
        return set_error(ctx, REACH_EUNSUPPORTED, "support program needs %zu features (max %d)",
                         P.feats.size(), LGG09_MAX_FEATS);
    return REACH_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
struct reach_lgg09_plan {
    reach_ctx* ctx = nullptr;
    int n = 0, ndirs = 0;
    int64_t nsteps = 0;
    reach_lgg09_opts opts{};

    // derived sizes
    int Kp = 0;            // n padded to BK
    int mu = 0, mu_pad = 0;
    int mu_ld = 0;               // stride of the per-chain feature arrays: mu_pad * oz_J (mu_pad without pass groups)
    int ne = 0, nbe = 0;   // mapped rows, rows per stack block
    int nq = 0;
    int S = 1;
    int cfg = 0, TM = 0, TN = 0, tpb = 0;
    int nq_pad = 0;        // rows of the Q (row-major power) buffers
    bool uploaded = false, ran = false;
    // row stack built from outside in steps (reach_lgg09_plan_stack_step): the ranks of a sharded run split the rows of
    // every doubling step and all-gather them instead of every rank building the whole stack
    bool stack_ready = false;
    int stack_steps_done = 0;
    Program prog;

    DevBuf stack;          // [(S+1)*nbe][Kp]
    DevBuf Qbuf[2];        // [nq_pad][Kp] : Phi^s row-major, ping-pong
    DevBuf L0, Lbuf[2];    // [mu_pad][Kp]
    DevBuf Wl, Wa;         // [nq][nbe]
    DevBuf tile_mask;      // [tpb] uint32
    DevBuf featpart;       // [(S+1)*tpb][nq][2][mu_pad]
    DevBuf ftot;           // [(S+1)][nq][2][mu_pad] per-block feature totals
    DevBuf carry[2];       // [nq][2][mu_pad], ping-pong between passes
    DevBuf dv;             // [S+2][2][mu_pad] rho(r_k; V) per emit slot of a pass
    DevBuf s_pos, s_neg;   // [mu_pad]
    DevBuf row_pos, row_neg;   // [mu] int
    DevBuf rho_own;        // ndirs x nsteps
    double* rho = nullptr; // rho_own or caller-provided
    // invariant: dual certificates per face (CSR, see first_disjoint_kernel)
    DevBuf chk_ptr, chk_rows, chk_w, chk_bound;
    int nchecks = 0;
    int inv_faces = 0, inv_covered = 0, inv_exact = 0;
    DevBuf first_bad;      // unsigned long long
    unsigned long long* fb_host = nullptr;   // pinned: polled copies of first_bad (early exit of the pass loop)
    cudaEvent_t ev_fb[2] = {nullptr, nullptr};
    int64_t nsteps_done = 0;
    int64_t steps_computed = 0;              // steps whose support values the last run produced (== nsteps without early exit)
    bool synced = false;
    DevBuf phi_cm;         // n x n column-major copy of Phi (source of the transpose)
    // persistent warp-per-chain path (small / sparse Phi)
    bool use_chain = false;
    bool packed = false;   // stack blocks packed with stride round_up(n+ne, 8) (large n, 128x128 tiles)
    // INT8 tcgen05 digit-plane path (path == 3)
    bool use_oz = false;
    bool oz_guard = false;        // auto-selected: sampled passes are re-run on the DMMA kernel and compared
    int oz_T = OZ_DEFAULT_T, oz_nkb = 0, oz_tiles_n = 0;
    // residue (Chinese remainder) form of the same path (crt_gemm.cuh): oz_crt moduli = INT8 products per FP64 product
    // (0: digit planes).  Same launches, guard and pass groups; Xq / Yq hold residue images instead of digit planes.
    int oz_crt = 0;
    int crt_lo = 0, crt_hi = 0;    // guarded auto path: start with crt_lo moduli, go to crt_hi if the first guard check is marginal
    double crt_escalated_at = 0.0; // the pass-0 guard value that made the run switch to crt_hi (0: it did not)
    CrtConsts crt{};
    DevBuf crt_scr;               // per-CTA residue scratch of the kernel
    DevBuf Xq, Yq, eX, eY;
    // Pass groups (few chains, e.g. one shard of a multi-GPU run): oz_J consecutive passes become ONE pair of launches.
    // The carry stack holds (Phi')^(jS), j = 1..oz_J; L_t x carry stack gives the direction rows of the next oz_J passes
    // at once (Lst), and the stacked rows [L_t; L_{t+1}; ...; L_{t+J-1}] x blocks 1..S is a GEMM with oz_J times the
    // rows -- the shape (and tensor-pipe efficiency) of the unsharded problem, and oz_J times fewer dependent launches.
    int oz_J = 1;
    int oz_ctiles = 0;            // 128-row tiles per carry block (round_up(n, 128) / 128)
    DevBuf Lst;                   // [oz_J][mu_pad][Kp]: L_{t+1} .. L_{t+J}
    DevBuf cstack;                // [oz_J][oz_ctiles*128][Kp] FP64: first n rows of block jS
    DevBuf Yq_c, eY_c;            // digit planes / exponents of the carry stack
    DevBuf ftot_ref, guard_val;
    // State after the last verified pass (direction block, running input sums, carried features): a failed check
    // restores it and replays every pass since then on the FP64 tensor pipe.
    DevBuf L_snap, s_snap, carry_snap;
    // Sentinel chains: up to 128 chains (evenly spaced over the template) are also propagated pass by pass on the FP64
    // tensor pipe (block S only, on their own stream).  At a checked pass their features are computed from the FP64
    // direction rows and compared with the INT8 path's features of the same chains, which started from the INT8 path's
    // own rows: the difference is the drift accumulated since pass 0, not only the error of one pass.
    DevBuf Lsh[2], ftot_sen, drift_val;
    int nsen = 0, sen_stride = 1;
    cudaStream_t st3 = nullptr;
    cudaEvent_t ev_sen[2] = {nullptr, nullptr};
    double oz_drift_max = 0.0;    // largest sentinel feature difference (INT8 chain vs FP64 chain), relative to the feature scale
    int64_t oz_replay_from = -1;  // first pass recomputed on the DMMA kernel after a failed check (-1: none)
    int64_t oz_replayed = 0;      // INT8 passes whose results were discarded and recomputed
    double oz_guard_max = 0.0;    // largest sampled (int8 - DMMA) feature difference, relative to the feature scale
    int64_t oz_fallback_pass = -1;   // pass at which the run fell back to the DMMA kernel (-1: never)
    int64_t oz_checks = 0;
    int ch_maxnnz = 0, ch_rpl = 0, ch_ne = 0;
    size_t ch_smem = 0;
    DevBuf ell_val, ell_col, Edev;
    // decoupled blocks: compact chains, parallel in time (path 4)
    bool use_compact = false;
    int cp_G = 0, cp_ne = 0, cp_nblk = 0, cp_log2 = 0, cp_mu_g = 0;
    unsigned cp_slot_used = 0;
    int64_t cp_sblk = 0;
    DevBuf cpM, cpr0, cpwl, cpwa, cpE, cpWlE, cpWaE, cprstart, cpvoff;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // side stream: the combine kernels (tile sums, running sum, evaluation) of pass t run beside the GEMM of pass t+1;
    // the per-tile partials are double-buffered by pass parity
    cudaStream_t st2 = nullptr;
    cudaEvent_t ev_gemm[2] = {nullptr, nullptr}, ev_comb[2] = {nullptr, nullptr};
    size_t fp_stride = 0;               // doubles per featpart buffer
    std::vector<cudaEvent_t> gemm_ev;   // pairs around sampled main-loop GEMM launches
    int gemm_sampled = 0;
    int64_t gemm_total = 0;
    int64_t launches = 0;
    double last_ms = 0.0, gemm_ms = 0.0;

    ~reach_lgg09_plan() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (auto e : ev_gemm) if (e) cudaEventDestroy(e);
        for (auto e : ev_comb) if (e) cudaEventDestroy(e);
        for (auto e : ev_sen) if (e) cudaEventDestroy(e);
        for (auto e : ev_fb) if (e) cudaEventDestroy(e);
        if (fb_host) cudaFreeHost(fb_host);
        if (st2) cudaStreamDestroy(st2);
        if (st3) cudaStreamDestroy(st3);
        for (auto e : gemm_ev) cudaEventDestroy(e);
    }
};

namespace {

template <int RPL, int NE>
int launch_chain_t(reach_ctx* ctx, const Lgg09Chain& c, size_t smem, cudaStream_t st) {
    auto kern = lgg09_chain_kernel<RPL, 4, NE>;
    REACH_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<c.mu, CHAIN_THREADS, smem, st>>>(c);                       // one CTA per chain
    REACH_CUDA(ctx, cudaGetLastError());
    return REACH_OK;
}
template <int RPL>
int launch_chain_r(reach_ctx* ctx, int ne, const Lgg09Chain& c, size_t smem, cudaStream_t st) {
    if (ne == 0) return launch_chain_t<RPL, 0>(ctx, c, smem, st);
    if (ne == 4) return launch_chain_t<RPL, 4>(ctx, c, smem, st);
    return launch_chain_t<RPL, 8>(ctx, c, smem, st);
}
int launch_chain(reach_ctx* ctx, int rpl, int ne, const Lgg09Chain& c, size_t smem, cudaStream_t st) {
    switch (rpl) {
        case 1: return launch_chain_r<1>(ctx, ne, c, smem, st);
        case 2: return launch_chain_r<2>(ctx, ne, c, smem, st);
        case 3: return launch_chain_r<3>(ctx, ne, c, smem, st);
        case 4: return launch_chain_r<4>(ctx, ne, c, smem, st);
    }
    return set_error(ctx, REACH_EINVAL, "bad rows-per-thread %d", rpl);
}

template <int G, int NALT, int NE>
int launch_compact_t(reach_ctx* ctx, const Lgg09Compact& c, bool inhomog, cudaStream_t st, int64_t& launches) {
    lgg09_compact_starts_kernel<G><<<(c.mu + 63) / 64, 64, 0, st>>>(c);
    REACH_CUDA(ctx, cudaGetLastError());
    ++launches;
    const long long warps = (long long)(c.mu_g / (32 / G)) * c.nblk;
    const unsigned grid = unsigned((warps + 3) / 4);
    if (inhomog && c.nblk > 1) {
        lgg09_compact_kernel<G, NALT, NE, true><<<grid, 128, 0, st>>>(c);
        REACH_CUDA(ctx, cudaGetLastError());
        lgg09_compact_scan_kernel<<<(2 * c.mu + 127) / 128, 128, 0, st>>>(c.voff, c.mu, c.nblk);
        REACH_CUDA(ctx, cudaGetLastError());
        launches += 2;
    } else {
        REACH_CUDA(ctx, cudaMemsetAsync(c.voff, 0, size_t(c.nblk) * c.mu * 2 * sizeof(double), st));
    }
    lgg09_compact_kernel<G, NALT, NE, false><<<grid, 128, 0, st>>>(c);
    REACH_CUDA(ctx, cudaGetLastError());
    ++launches;
    return REACH_OK;
}
template <int G, int NALT>
int launch_compact_g(reach_ctx* ctx, int ne, const Lgg09Compact& c, bool inhomog, cudaStream_t st, int64_t& launches) {
    if (ne == 0) return launch_compact_t<G, NALT, 0>(ctx, c, inhomog, st, launches);
    if (ne == 4) return launch_compact_t<G, NALT, 4>(ctx, c, inhomog, st, launches);
    return launch_compact_t<G, NALT, 8>(ctx, c, inhomog, st, launches);
}
template <int NALT>
int launch_compact_a(reach_ctx* ctx, int G, int ne, const Lgg09Compact& c, bool inhomog, cudaStream_t st, int64_t& launches) {
    switch (G) {
        case 2: return launch_compact_g<2, NALT>(ctx, ne, c, inhomog, st, launches);
        case 4: return launch_compact_g<4, NALT>(ctx, ne, c, inhomog, st, launches);
        case 8: return launch_compact_g<8, NALT>(ctx, ne, c, inhomog, st, launches);
        case 16: return launch_compact_g<16, NALT>(ctx, ne, c, inhomog, st, launches);
    }
    return set_error(ctx, REACH_EINVAL, "bad compact group size %d", G);
}
int launch_compact(reach_ctx* ctx, int G, int ne, const Lgg09Compact& c, bool inhomog, cudaStream_t st, int64_t& launches) {
    if (c.prog.nalt == 1) return launch_compact_a<1>(ctx, G, ne, c, inhomog, st, launches);
    return launch_compact_a<2>(ctx, G, ne, c, inhomog, st, launches);
}

// Cost model (SM cycles per time step) used to pick the tile shape and the time block S.
// Calibrated on B200: a 128x128x2000 tile takes ~0.31 ms (~55 FP64 FMA/clk/SM sustained); smaller
// tiles lose a little to shared-memory traffic; every pass costs two launches; building the row
// stack costs about S + 2 log2(S) passes of an n x n x n product, amortised over the run.
double step_cost(const reach_lgg09_plan& p, const TileCfg& c, int S, int mu, int nrows, int sms) {
    // cfg 0 is the warp-specialised kernel; cfg 4 the digit-plane GEMM (7 planes of 8 bits: ~3x the DMMA rate)
    static const double eff[kNumCfgs] = {1.12, 0.85, 0.9, 0.92, 3.0, 0.85};
    const int mu_pad = round_up(mu, c.TM);
    const bool packed = (c.id == 0 || c.id == kOzCfg) && nrows >= 128;
    const int nbe = packed ? round_up(nrows, 8) : round_up(nrows, c.TN);
    const double slots = double(sms) * c.occ;
    auto pass = [&](double tiles) {
        const double waves = std::ceil(tiles / slots);
        const double mma = double(c.TM) * c.TN * p.Kp / (55.0 * eff[c.id]) * c.occ;
        const double mem = double(c.TM + c.TN) * p.Kp * 8.0 / 40.0 * c.occ;   // ~40 B/clk/SM from L2
        return waves * (std::max(mma, mem) + (c.id == 0 || c.id == kOzCfg ? 12000.0 : 4000.0));   // prologue + epilogue
    };
    // ~30 us per pass: GEMM + tile-sum + eval + prefix launches back to back.  The persistent digit-plane kernel pays about
    // one more tile time per launch (ring fill, the last tile's epilogue, the slice of L in front of it): measured on
    // shards of C5, 1000 chains: S = 14 / 20 / 28 -> 901 / 880 / 855 ms.
    // (the residue kernel: + the 0.11 ms residue conversion of L in front of every launch and a longer tail -- measured on C5, 10^4
    // steps: S = 20 / 32 / 40 / 48 / 64 / 80 -> 1398 / 1369 / 1357 / 1334 / 1362 / 1369 ms)
    const double launch = c.id == kOzCfg ? (p.oz_crt ? 450000.0 : 250000.0) : 60000.0;
    double gemm_pass = pass(double(mu_pad / c.TM) * std::ceil(double(S) * nbe / c.TN));
    if (c.id == kOzCfg && p.oz_crt) {
        // residue kernel: CTA pairs on 256 x 256 tiles, one pair per two SMs -- whole waves of those
        const double pair_tiles = std::ceil(mu_pad / 256.0) * (std::floor(double(S) * nbe / 256.0) + 1.0);
        const double waves = std::ceil(pair_tiles / std::max(1, sms / 2));
        const double mma = 128.0 * 128.0 * p.Kp / (55.0 * eff[c.id]);
        gemm_pass = waves * (4.0 * mma + 12000.0);
    }
    double cost = (gemm_pass + launch) / S;
    // stack build by doubling (64x64 tiles, occupancy 2), amortised: one launch per product, each as many waves as its
    // row count needs (for small n a whole doubling step is a single wave -- the build is ~2 log2(S) launches, not S)
    auto product = [&](double rows) {
        const double tiles = std::ceil(rows / 64.0) * std::ceil(p.n / 64.0);
        return std::ceil(tiles / (sms * 2.0)) * (64.0 * 64.0 * p.Kp / (55.0 * 0.85) * 2 + 4000.0) + 6000.0;
    };
    double build = product(nbe);
    for (int s = 1; s < S; s *= 2) {
        if (2 * s < S) build += product(p.n);
        build += product(double(std::min(s, S - s)) * nbe);
    }
    cost += build / double(std::max<int64_t>(1, p.nsteps));
    return cost;
}

int choose_config(reach_lgg09_plan& p) {
    reach_ctx* ctx = p.ctx;
    const int nrows = p.n + p.ne;
    const int smax = int(std::min<int64_t>(1024, std::max<int64_t>(1, p.nsteps)));
    int best_cfg = -1, best_S = 1;
    double best = 1e300;
    for (int ci = 0; ci < kNumCfgs; ++ci) {
        const TileCfg& c = kCfgs[ci];
        if ((ci == kOzCfg) != p.use_oz) continue;
        if (p.opts.tile_m && p.opts.tile_m != c.TM) continue;
        if (p.opts.tile_n && p.opts.tile_n != c.TN) continue;
        for (int S = 1; S <= std::max(smax, p.opts.time_block); ++S) {
            if (p.opts.time_block > 0 && S != p.opts.time_block) continue;
            if (p.opts.time_block == 0 && S > 16 && (S & 3)) continue;        // coarse scan for large S
            if (p.opts.time_block == 0 && S > 128 && (S & 31)) continue;
            const int nbe = round_up(nrows, c.TN);
            const int mu_pad = round_up(p.mu, c.TM);
            const double stack_bytes = double(S + 1) * nbe * p.Kp * 8.0;
            const double feat_bytes = double(S + 1) * (nbe / c.TN) * std::max(1, p.nq) * 2.0 * mu_pad * 8.0;
            if (S > 1 && (stack_bytes > 3e9 || feat_bytes > 1e9)) continue;
            const double cost = step_cost(p, c, S, p.mu * (ci == kOzCfg ? p.oz_J : 1), nrows, ctx->sm_count);
            // prefer the earlier (larger-tile, smaller-S) candidate unless clearly beaten
            if (cost < best * 0.985) { best = cost; best_cfg = ci; best_S = S; }
        }
    }
    if (best_cfg < 0)
        return set_error(ctx, REACH_EINVAL,
                         "no launch configuration matches time_block=%d tile_m=%d tile_n=%d "
                         "(tiles are 128x128, 64x64, 64x48, 16x128 or 64x128; the row stack must fit 3 GB)",
                         p.opts.time_block, p.opts.tile_m, p.opts.tile_n);
    p.cfg = best_cfg;
    p.S = best_S;
    p.TM = kCfgs[best_cfg].TM;
    p.TN = kCfgs[best_cfg].TN;
    return REACH_OK;
}

}  // namespace

extern "C" void reach_lgg09_default_opts(reach_lgg09_opts* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->time_block = 0;
    o->share_negated = 1;
}

extern "C" int reach_lgg09_plan_create(reach_ctx* ctx, int n, int ndirs, int64_t nsteps,
                                       const reach_lgg09_opts* opts, reach_lgg09_plan** out) {
    if (!ctx) return set_error(nullptr, REACH_EINVAL, "ctx is NULL");
    if (!out) return set_error(ctx, REACH_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 1) return set_error(ctx, REACH_EINVAL, "state dimension must be >= 1, got %d", n);
    if (ndirs < 1) return set_error(ctx, REACH_EINVAL, "need at least one template direction, got %d", ndirs);
    if (nsteps < 1) return set_error(ctx, REACH_EINVAL, "NSTEPS must be >= 1, got %lld", (long long)nsteps);
    if (opts && opts->time_block < 0) return set_error(ctx, REACH_EINVAL, "time_block < 0");
    if (opts && opts->time_block > 1024)
        return set_error(ctx, REACH_EINVAL, "time_block must be <= 1024, got %d", opts->time_block);
    reach_lgg09_plan* p = new (std::nothrow) reach_lgg09_plan;
    if (!p) return set_error(ctx, REACH_ENOMEM, "out of host memory");
    p->ctx = ctx;
    p->n = n;
    p->ndirs = ndirs;
    p->nsteps = nsteps;
    if (opts) p->opts = *opts; else reach_lgg09_default_opts(&p->opts);
    p->Kp = round_up(n, BK);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->st2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->st3, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p->ev_gemm[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_comb[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_sen[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        delete p;
        return set_error(ctx, REACH_ECUDA, "plan_create: %s", cudaGetErrorString(e));
    }
    *out = p;
    return REACH_OK;
}

extern "C" void reach_lgg09_plan_destroy(reach_lgg09_plan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    if (p->st2) cudaStreamSynchronize(p->st2);      // an error return can leave combine kernels in flight on the side stream
    if (p->st3) cudaStreamSynchronize(p->st3);
    cudaStreamSynchronize(p->ctx->stream);
    delete p;
}

extern "C" int reach_lgg09_plan_upload(reach_lgg09_plan* p, const double* Phi, const double* dirs,
                                       const reach_setexpr* Omega0, const reach_setexpr* V,
                                       const reach_invariant* inv) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!Phi || !dirs) return set_error(ctx, REACH_EINVAL, "Phi and dirs must not be NULL");
    NvtxRange nvtx_fn("reach_lgg09_plan_upload (flatten program, choose path, H2D)");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = p->n, Kp = p->Kp, ndirs = p->ndirs;
    for (size_t i = 0; i < size_t(n) * n; ++i)
        if (!std::isfinite(Phi[i])) return set_error(ctx, REACH_EINVAL, "Phi contains a non-finite entry");

    // ---- support program ----
    p->prog = Program();
    REACH_TRY(build_program(ctx, n, Omega0, V, p->prog));
    p->ne = int(p->prog.erows.size());
    p->nq = int(p->prog.feats.size());
    p->prog.dev.box = p->opts.reserved[1] ? 1 : 0;
    if (p->prog.dev.box) {
        if (p->prog.dev.nalt != 1)
            return set_error(ctx, REACH_EINVAL, "box output (opts.reserved[1]) needs Omega0 to be a single sum of terms, got %d alternatives",
                             p->prog.dev.nalt);
        if (inv && inv->kind != REACH_INV_NONE)
            return set_error(ctx, REACH_EUNSUPPORTED, "box output does not take an invariant (the caller checks disjointness on (c, r))");
    }

    // ---- direction chains: l and -l share (Phi')^k l (reach_homog.jl:108-110) ----
    // chain c = sign_c * direction src_c (canonical orientation: first non-zero entry positive)
    std::vector<int> row_pos, row_neg, chain_src;
    std::vector<double> chain_sign;
    {
        std::unordered_multimap<uint64_t, int> index;
        index.reserve(size_t(ndirs) * 2);
        for (int j = 0; j < ndirs; ++j) {
            const double* d = dirs + size_t(j) * n;
            // canonical orientation first (sign of the first non-zero entry), then an FNV-style hash of sign * d in four
            // independent lanes (one dependent multiply per element was ~10 ms for 4000 x 2000 directions)
            double sign = 1.0;
            if (p->opts.share_negated)
                for (int i = 0; i < n; ++i)
                    if (d[i] != 0.0) { sign = d[i] < 0 ? -1.0 : 1.0; break; }
            uint64_t hl[4] = {1469598103934665603ull, 0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull};
            bool finite = true;
            for (int i = 0; i < n; ++i) {
                const double v = d[i];
                finite &= std::isfinite(v);
                const double cv = sign * v + 0.0;                 // -0 hashes as +0
                uint64_t bits;
                std::memcpy(&bits, &cv, 8);
                hl[i & 3] = (hl[i & 3] ^ bits) * 1099511628211ull;
            }
            if (!finite) return set_error(ctx, REACH_EINVAL, "direction %d is not finite", j);
            const uint64_t h = ((hl[0] * 31 + hl[1]) * 31 + hl[2]) * 31 + hl[3];
            int c = -1;
            if (p->opts.share_negated) {
                auto range = index.equal_range(h);
                for (auto it = range.first; it != range.second && c < 0; ++it) {
                    const int cand = it->second;
                    const double* e = dirs + size_t(chain_src[cand]) * n;
                    const double rel = sign * chain_sign[cand];   // d == rel * e ?
                    bool same = true;
                    for (int i = 0; i < n && same; ++i) same = (d[i] == rel * e[i]);
                    if (!same) continue;
                    int& slot = sign > 0 ? row_pos[cand] : row_neg[cand];
                    if (slot < 0) { slot = j; c = cand; }        // else: duplicate direction, own chain
                }
            }
            if (c < 0) {
                c = int(chain_src.size());
                chain_src.push_back(j);
                chain_sign.push_back(sign);
                row_pos.push_back(sign > 0 ? j : -1);
                row_neg.push_back(sign > 0 ? -1 : j);
                if (p->opts.share_negated) index.emplace(h, c);
            }
        }
    }
    p->mu = int(chain_src.size());

    p->use_compact = false;
    const bool forced_opts = p->opts.path == 1 || p->opts.path == 3 || (p->opts.path == 0 && (p->opts.time_block > 0 || p->opts.tile_m || p->opts.tile_n));
    // ---- decoupled blocks: closure of every chain's support under the sparsity graph of Phi (path 4) ----
    std::vector<int> cp_idx;                                  // [mu][G] closure rows, ascending, -1 padded
    if (!forced_opts && p->opts.path != 2 && p->prog.dev.nalt <= 2 && p->ne <= 8 && (p->opts.path == 4 || !getenv("REACH_NO_COMPACT"))) {
        constexpr int GMAX = 16;
        size_t nnz = 0;
        for (size_t i = 0; i < size_t(n) * n; ++i) nnz += Phi[i] != 0.0;
        if (nnz <= size_t(GMAX) * n) {                        // components of <= 16 rows hold <= 16 n non-zeros
            std::vector<int> parent(n);
            for (int i = 0; i < n; ++i) parent[i] = i;
            auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
            for (int i = 0; i < n; ++i)
                for (int l = 0; l < n; ++l)
                    if (Phi[size_t(i) * n + l] != 0.0) { const int a = find(i), b = find(l); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
            std::vector<int> comp(n), csize(n, 0);
            for (int i = 0; i < n; ++i) { comp[i] = find(i); ++csize[comp[i]]; }
            std::vector<std::vector<int>> members(n);
            for (int i = 0; i < n; ++i) members[comp[i]].push_back(i);
            // mapped rows and direct weights only matter inside the closure, so they do not enlarge it
            int gmax = 1;
            bool ok = true;
            std::vector<std::vector<int>> closure(p->mu);
            for (int c = 0; c < p->mu && ok; ++c) {
                const double* d = dirs + size_t(chain_src[c]) * n;
                std::vector<int> comps;
                int size = 0;
                for (int i = 0; i < n && ok; ++i) {
                    if (d[i] == 0.0) continue;
                    if (std::find(comps.begin(), comps.end(), comp[i]) != comps.end()) continue;
                    comps.push_back(comp[i]);
                    size += csize[comp[i]];
                    ok = size <= GMAX;
                }
                if (!ok) break;
                for (int cid : comps) closure[c].insert(closure[c].end(), members[cid].begin(), members[cid].end());
                std::sort(closure[c].begin(), closure[c].end());
                gmax = std::max(gmax, size);
            }
            if (ok) {
                int G = 2;
                while (G < gmax) G *= 2;
                p->use_compact = true;
                p->cp_G = G;
                cp_idx.assign(size_t(p->mu) * G, -1);
                for (int c = 0; c < p->mu; ++c) std::copy(closure[c].begin(), closure[c].end(), cp_idx.begin() + size_t(c) * G);
            }
        }
        if (p->opts.path == 4 && !p->use_compact)
            return set_error(ctx, REACH_EUNSUPPORTED, "the compact-chain path needs every direction's closure under the sparsity "
                             "graph of Phi to span at most %d rows", GMAX);
    } else if (p->opts.path == 4) {
        return set_error(ctx, REACH_EUNSUPPORTED, "the compact-chain path needs <= 2 alternatives in Omega0 and <= 8 mapped rows "
                         "(alternatives=%d, mapped rows=%d)", p->prog.dev.nalt, p->ne);
    }
    // ---- INT8 tcgen05 digit-plane GEMM requested (path 3) ----
    p->use_oz = false;
    p->oz_guard = false;
    if (!p->use_compact && p->opts.path == 0 && !p->opts.tile_m && !p->opts.tile_n && n >= 512 && n <= OZ_MAX_K && p->mu >= 96 &&
        n + p->ne >= 128 && p->nq >= 1 && p->nq <= 4 && !getenv("REACH_NO_INT8")) {
        // large dense problem: the INT8 tcgen05 digit-plane GEMM (~2x the FP64 tensor pipe), guarded: sampled passes
        // are recomputed with the DMMA kernel and the run falls back to it if they differ by more than kOzGuardTol
        p->use_oz = true;
        p->oz_guard = true;
    }
    if (p->opts.path == 3) {
        if (n + p->ne < 128 || p->nq < 1 || p->nq > 4 || n > OZ_MAX_K)
            return set_error(ctx, REACH_EUNSUPPORTED,
                             "the INT8 GEMM path needs 128 <= n + mapped rows, n <= 16384 and 1..4 support features "
                             "(n=%d, mapped rows=%d, features=%d)", n, p->ne, p->nq);
        p->use_oz = true;
    }
    p->oz_crt = 0;
    if (p->use_oz) {
        // opts.reserved[0]: 0 = default (residues modulo CRT_DEFAULT_N moduli), 4..8 = that many 8-bit digit planes,
        // 10..16 = residues modulo that many moduli
        int sel = p->opts.reserved[0];
        p->crt_lo = p->crt_hi = 0;
        if (sel == 0) {
            sel = getenv("REACH_NO_CRT") ? OZ_DEFAULT_T : CRT_DEFAULT_N;
            if (const char* e = getenv("REACH_CRT_MODULI")) sel = atoi(e);
            else if (sel == CRT_DEFAULT_N && p->oz_guard && !getenv("REACH_CRT_NO_ADAPT")) {
                // guarded: one modulus fewer (the digit planes' accuracy class, 7 % less work) unless the first guard check
                // says the problem needs the full count -- then the run restarts at pass 0 with CRT_DEFAULT_N moduli
                p->crt_lo = CRT_DEFAULT_N - 1;
                p->crt_hi = CRT_DEFAULT_N;
                sel = p->crt_lo;
            }
        }
        if (sel >= 4 && sel <= OZ_MAX_T) {
            p->oz_T = sel;
        } else if (sel >= 10 && sel <= CRT_MAX_N) {
            p->oz_crt = sel;
            p->oz_T = sel;
            p->crt = crt_make_consts(sel);
        } else {
            return set_error(ctx, REACH_EINVAL, "opts.reserved[0] must be 0, 4..%d (digit planes) or 10..%d (moduli), got %d", OZ_MAX_T,
                             CRT_MAX_N, sel);
        }
    }
    p->oz_J = 1;
    if (p->use_oz && !(inv && inv->kind != REACH_INV_NONE) && !getenv("REACH_OZ_NO_GROUPS")) {
        // pass groups: stack passes until the GEMM has ~2048 rows (16 row tiles).  Not with an invariant: the early exit
        // is decided pass by pass.
        const int mu_pad128 = round_up(p->mu, OZ_TM);
        // measured on shards of C5 (n = 2000): 250 chains J = 8: 325 -> 247 ms, 500 chains J = 4: 484 -> 462 ms, 1000 chains
        // J = 2: no gain over single passes (the carry GEMM costs what the fuller waves save)
        int J = mu_pad128 <= 512 ? std::max(1, std::min(16, 2048 / mu_pad128)) : 1;
        if (const char* e = getenv("REACH_OZ_GROUP")) J = std::max(1, std::min(16, atoi(e)));
        p->oz_J = J;
    }
    REACH_TRY(choose_config(*p));
    // ---- small / sparse Phi: ELL form of Phi' (row i of Phi' = column i of Phi) ----
    std::vector<double> ell_val;
    std::vector<int> ell_col;
    p->use_chain = false;
    const bool forced_gemm = p->opts.path == 1 || p->opts.path == 3 || p->use_oz || (p->opts.path == 0 && (p->opts.time_block > 0 || p->opts.tile_m || p->opts.tile_n));
    if (p->use_compact) p->S = 1;
    if (!p->use_compact && !forced_gemm && n <= 512 && p->nq <= 4 && p->ne <= 8) {
        int maxnnz = 0;
        for (int i = 0; i < n; ++i) {
            int cnt = 0;
            for (int l = 0; l < n; ++l) cnt += Phi[size_t(i) * n + l] != 0.0;
            maxnnz = std::max(maxnnz, cnt);
        }
        maxnnz = std::max(maxnnz, 1);
        const int rpl = (n + CHAIN_THREADS - 1) / CHAIN_THREADS;
        const int npad = rpl * CHAIN_THREADS, nep = p->ne == 0 ? 0 : (p->ne <= 4 ? 4 : 8);
        const size_t smem = size_t(maxnnz) * npad * 12 + size_t(8 + nep) * npad * 8 + size_t(CHAIN_B + 1) * npad * 8 +
                            size_t(CHAIN_B) * (8 + nep) * CHAIN_RED_LD * 8 + size_t(CHAIN_B + 1) * (8 + nep) * 8 +
                            size_t(CHAIN_B) * 2 * 8 + 64;
        // cycles per time step of one CTA (4 warps, one chain): SpMV + feature partials of RPL rows per thread, a
        // butterfly, one CTA barrier; several CTAs share an SM
        // (measured: ISS rpl 3, 2 non-zeros per row: ~1370 cycles; Building rpl 1, 48 per row: ~2240)
        const double chain_cycles = 700.0 + rpl * (30.0 * maxnnz + 200.0);
        const double per_sm = std::max(1.0, std::min(8.0, std::floor(200.0 * 1024 / double(smem))));
        const double waves = std::ceil(double(p->mu) / (ctx->sm_count * per_sm));
        const double chain_cost = waves * chain_cycles;
        const double gemm_cost = step_cost(*p, kCfgs[p->cfg], p->S, p->mu, n + p->ne, ctx->sm_count);
        if (smem <= ctx->smem_optin && (p->opts.path == 2 || chain_cost < gemm_cost)) {
            p->use_chain = true;
            p->ch_maxnnz = maxnnz;
            p->ch_rpl = rpl;
            p->ch_ne = nep;
            p->ch_smem = smem;
            ell_val.assign(size_t(maxnnz) * npad, 0.0);
            ell_col.assign(size_t(maxnnz) * npad, 0);
            for (int i = 0; i < n; ++i) {
                int z = 0;
                for (int l = 0; l < n; ++l) {
                    const double v = Phi[size_t(i) * n + l];       // Phi(l, i) = Phi'(i, l)
                    if (v != 0.0) { ell_val[size_t(z) * npad + i] = v; ell_col[size_t(z) * npad + i] = l; ++z; }
                }
            }
            p->S = 1;
        }
    }
    if (p->opts.path == 2 && !p->use_chain)
        return set_error(ctx, REACH_EUNSUPPORTED,
                         "the persistent chain kernel needs n <= 512, <= 4 support features, <= 8 mapped rows and Phi' "
                         "in ELL form within shared memory (n=%d, features=%d, mapped rows=%d)", n, p->nq, p->ne);
    const int TM = p->TM, TN = p->TN, S = p->S;
    p->mu_pad = round_up(p->mu, TM);
    p->packed = !p->use_chain && (p->cfg == 0 || p->use_oz) && (n + p->ne) >= 128 && p->nq >= 1 && p->nq <= 8;
    p->nbe = p->packed ? round_up(n + p->ne, 8) : round_up(n + p->ne, TN);
    p->tpb = p->packed ? 1 : p->nbe / TN;
    const size_t slice_tiles = (size_t(S + 1) * p->nbe + TN - 1) / TN;      // row tiles of the longest launch
    p->nq_pad = round_up(n, 128);
    p->mu_ld = p->mu_pad * (p->use_oz ? p->oz_J : 1);
    const int nbe = p->nbe, mu_pad = p->mu_pad, mu_ld = p->mu_ld, nq = p->nq;

    // ---- device buffers ----
    REACH_TRY(p->stack.alloc(ctx, (size_t(S + 1) * nbe + 256) * Kp * sizeof(double)));     // + slack rows (zero)
    REACH_TRY(p->Qbuf[0].alloc(ctx, size_t(p->nq_pad) * Kp * sizeof(double)));
    REACH_TRY(p->Qbuf[1].alloc(ctx, size_t(p->nq_pad) * Kp * sizeof(double)));
    REACH_TRY(p->L0.alloc(ctx, size_t(mu_pad) * Kp * sizeof(double)));
    REACH_TRY(p->Lbuf[0].alloc(ctx, size_t(mu_pad) * Kp * sizeof(double)));
    REACH_TRY(p->Lbuf[1].alloc(ctx, size_t(mu_pad) * Kp * sizeof(double)));
    REACH_TRY(p->Wl.alloc(ctx, size_t(std::max(1, nq)) * nbe * sizeof(double)));
    REACH_TRY(p->Wa.alloc(ctx, size_t(std::max(1, nq)) * nbe * sizeof(double)));
    REACH_TRY(p->tile_mask.alloc(ctx, size_t(p->tpb) * sizeof(uint32_t)));
    p->fp_stride = (p->packed ? slice_tiles * 2 : size_t(S + 1) * p->tpb) * std::max(1, nq) * 2 * mu_ld;
    REACH_TRY(p->featpart.alloc(ctx, 2 * p->fp_stride * sizeof(double)));
    REACH_TRY(p->ftot.alloc(ctx, size_t(S + 1) * std::max(1, nq) * 2 * mu_ld * sizeof(double)));
    REACH_TRY(p->carry[0].alloc(ctx, size_t(std::max(1, nq)) * 2 * mu_ld * sizeof(double)));
    REACH_TRY(p->carry[1].alloc(ctx, size_t(std::max(1, nq)) * 2 * mu_ld * sizeof(double)));
    REACH_TRY(p->dv.alloc(ctx, size_t(S + 2) * 2 * mu_ld * sizeof(double)));
    REACH_TRY(p->s_pos.alloc(ctx, size_t(mu_pad) * sizeof(double)));
    REACH_TRY(p->s_neg.alloc(ctx, size_t(mu_pad) * sizeof(double)));
    REACH_TRY(p->row_pos.alloc(ctx, size_t(p->mu) * sizeof(int)));
    REACH_TRY(p->row_neg.alloc(ctx, size_t(p->mu) * sizeof(int)));
    REACH_TRY(p->phi_cm.alloc(ctx, size_t(n) * n * sizeof(double), false));
    REACH_TRY(p->first_bad.alloc(ctx, sizeof(unsigned long long)));
    if (!p->rho) {
        REACH_TRY(p->rho_own.alloc(ctx, size_t(ndirs) * p->nsteps * sizeof(double), false));
        p->rho = p->rho_own.as<double>();
    }

    if (p->use_oz) {
        const int kblock = p->oz_crt ? CRT_KB : OZ_KB;
        p->oz_nkb = (n + kblock - 1) / kblock;
        p->oz_tiles_n = int((size_t(S + 1) * nbe + OZ_TN - 1) / OZ_TN);
        // bytes of all planes (digit planes or residue images) of one 128-row tile and one k-block
        const int crt_alloc = std::max(p->oz_crt, p->crt_hi);           // room for the escalated modulus count
        const size_t plane_blk = p->oz_crt ? size_t(crt_alloc) * CRT_IMG : size_t(p->oz_T) * OZ_PLANE;
        if (p->oz_crt) REACH_TRY(p->crt_scr.alloc(ctx, crt_scratch_bytes(crt_alloc, ctx->sm_count), false));
        REACH_TRY(p->Yq.alloc(ctx, size_t(p->oz_nkb) * p->oz_tiles_n * plane_blk, false));
        REACH_TRY(p->eY.alloc(ctx, size_t(p->oz_tiles_n) * OZ_TN * sizeof(int)));
        REACH_TRY(p->Xq.alloc(ctx, size_t(p->oz_nkb) * (mu_ld / OZ_TM) * plane_blk, false));
        REACH_TRY(p->eX.alloc(ctx, size_t(mu_ld) * sizeof(int)));
        if (p->oz_J > 1) {
            p->oz_ctiles = round_up(n, OZ_TN) / OZ_TN;
            const size_t crows = size_t(p->oz_J) * p->oz_ctiles * OZ_TN;
            REACH_TRY(p->Lst.alloc(ctx, size_t(p->oz_J) * mu_pad * Kp * sizeof(double)));
            REACH_TRY(p->cstack.alloc(ctx, crows * Kp * sizeof(double)));
            REACH_TRY(p->Yq_c.alloc(ctx, size_t(p->oz_nkb) * (crows / OZ_TN) * plane_blk, false));
            REACH_TRY(p->eY_c.alloc(ctx, crows * sizeof(int)));
        }
        if (p->oz_guard) {
            REACH_TRY(p->ftot_ref.alloc(ctx, p->ftot.bytes));
            REACH_TRY(p->guard_val.alloc(ctx, sizeof(unsigned long long)));
            REACH_TRY(p->L_snap.alloc(ctx, size_t(mu_pad) * Kp * sizeof(double)));
            REACH_TRY(p->s_snap.alloc(ctx, size_t(2) * mu_pad * sizeof(double)));
            REACH_TRY(p->carry_snap.alloc(ctx, p->carry[0].bytes));
            p->nsen = std::min(p->mu, OZ_TM);
            p->sen_stride = p->mu / p->nsen;
            REACH_TRY(p->Lsh[0].alloc(ctx, size_t(OZ_TM) * Kp * sizeof(double)));
            REACH_TRY(p->Lsh[1].alloc(ctx, size_t(OZ_TM) * Kp * sizeof(double)));
            REACH_TRY(p->ftot_sen.alloc(ctx, p->ftot.bytes));
            REACH_TRY(p->drift_val.alloc(ctx, sizeof(unsigned long long)));
        }
    }
    if (p->use_chain) {
        REACH_TRY(p->ell_val.alloc(ctx, ell_val.size() * 8, false));
        REACH_TRY(p->ell_col.alloc(ctx, ell_col.size() * 4, false));
        REACH_CUDA(ctx, cudaMemcpyAsync(p->ell_val.p, ell_val.data(), ell_val.size() * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaMemcpyAsync(p->ell_col.p, ell_col.data(), ell_col.size() * 4, cudaMemcpyHostToDevice, st));
        std::vector<double> E(size_t(std::max(1, p->ne)) * n, 0.0);
        for (int e = 0; e < p->ne; ++e) std::memcpy(&E[size_t(e) * n], p->prog.erows[e].data(), sizeof(double) * n);
        REACH_TRY(p->Edev.alloc(ctx, E.size() * 8, false));
        REACH_CUDA(ctx, cudaMemcpyAsync(p->Edev.p, E.data(), E.size() * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (p->use_compact) {
        const int G = p->cp_G, mu = p->mu;
        const int nep = p->ne == 0 ? 0 : (p->ne <= 4 ? 4 : 8);
        p->cp_ne = nep;
        p->cp_mu_g = round_up(mu, 32 / G);
        // time blocks: enough (chain, block) groups for ~16 warps per SM, blocks of 2^s >= 64 steps
        const int64_t wpb = p->cp_mu_g / (32 / G);
        const int64_t want_blocks = std::max<int64_t>(1, (int64_t(ctx->sm_count) * 16 + wpb - 1) / wpb);
        int64_t sblk = 64;
        int lg = 6;
        while (sblk * want_blocks < p->nsteps) { sblk *= 2; ++lg; }
        p->cp_sblk = sblk;
        p->cp_log2 = lg;
        p->cp_nblk = int((p->nsteps + sblk - 1) / sblk);
        // evaluation slots: 2a = feature of alternative a at r_k, 2a+1 = at r_{k+1}, last = rho(.; V); -1 = absent
        const int nalt = p->prog.dev.nalt, NS = 2 * nalt + 1;
        std::vector<int> slot_q(NS, -1);
        for (int a = 0; a < nalt; ++a) { slot_q[2 * a] = p->prog.dev.q0[a]; slot_q[2 * a + 1] = p->prog.dev.q1[a]; }
        slot_q[2 * nalt] = p->prog.dev.qV;
        p->cp_slot_used = 0;
        for (int q = 0; q < NS; ++q)
            if (slot_q[q] >= 0) p->cp_slot_used |= 1u << q;
        std::vector<double> M(size_t(mu) * G * G, 0.0), r0(size_t(mu) * G, 0.0), wl(size_t(mu) * NS * G, 0.0),
            wa(size_t(mu) * NS * G, 0.0), E(size_t(mu) * std::max(1, nep) * G, 0.0), WlE(size_t(NS) * std::max(1, p->ne), 0.0),
            WaE(size_t(NS) * std::max(1, p->ne), 0.0);
        for (int c = 0; c < mu; ++c) {
            const int* idx = &cp_idx[size_t(c) * G];
            const double* d = dirs + size_t(chain_src[c]) * n;
            for (int i = 0; i < G; ++i) {
                if (idx[i] < 0) continue;
                r0[size_t(c) * G + i] = chain_sign[c] * d[idx[i]] + 0.0;
                for (int l = 0; l < G; ++l)
                    if (idx[l] >= 0) M[(size_t(c) * G + i) * G + l] = Phi[size_t(idx[i]) * n + idx[l]];   // Phi(idx_l, idx_i)
                for (int q = 0; q < NS; ++q) {
                    if (slot_q[q] < 0) continue;
                    wl[(size_t(c) * NS + q) * G + i] = p->prog.feats[slot_q[q]].a[idx[i]];
                    wa[(size_t(c) * NS + q) * G + i] = p->prog.feats[slot_q[q]].b[idx[i]];
                }
                for (int e = 0; e < p->ne; ++e) E[(size_t(c) * nep + e) * G + i] = p->prog.erows[e][idx[i]];
            }
        }
        for (int q = 0; q < NS; ++q) {
            if (slot_q[q] < 0) continue;
            for (auto& kv : p->prog.feats[slot_q[q]].rows) {
                WlE[size_t(q) * p->ne + kv.first] = kv.second.first;
                WaE[size_t(q) * p->ne + kv.first] = kv.second.second;
            }
        }
        auto ship = [&](DevBuf& b, const std::vector<double>& v) -> int {
            REACH_TRY(b.alloc(ctx, v.size() * 8, false));
            REACH_CUDA(ctx, cudaMemcpyAsync(b.p, v.data(), v.size() * 8, cudaMemcpyHostToDevice, st));
            return REACH_OK;
        };
        REACH_TRY(ship(p->cpM, M));
        REACH_TRY(ship(p->cpr0, r0));
        REACH_TRY(ship(p->cpwl, wl));
        REACH_TRY(ship(p->cpwa, wa));
        REACH_TRY(ship(p->cpE, E));
        REACH_TRY(ship(p->cpWlE, WlE));
        REACH_TRY(ship(p->cpWaE, WaE));
        REACH_TRY(p->cprstart.alloc(ctx, size_t(p->cp_nblk) * mu * G * 8, false));
        REACH_TRY(p->cpvoff.alloc(ctx, size_t(p->cp_nblk) * mu * 2 * 8, false));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }
    // ---- host staging + H2D ----
    // Phi: column-major n x n.  Row k of the row-major Phi (the seed Q_1) is produced on the
    // device by a transpose at run time; here we only ship the matrix.
    REACH_CUDA(ctx, cudaMemcpyAsync(p->phi_cm.p, Phi, size_t(n) * n * sizeof(double), cudaMemcpyHostToDevice, st));
    {   // block 0 of the stack = [I ; E]: ones on the diagonal (kernel), mapped rows copied row by row
        set_identity_kernel<<<(n + 255) / 256, 256, 0, st>>>(p->stack.as<double>(), Kp, n);
        REACH_CUDA(ctx, cudaGetLastError());
        if (p->ne > 0) {
            std::vector<double> E(size_t(p->ne) * n);
            for (int e = 0; e < p->ne; ++e) std::memcpy(&E[size_t(e) * n], p->prog.erows[e].data(), sizeof(double) * n);
            REACH_CUDA(ctx, cudaMemcpy2DAsync(p->stack.as<double>() + size_t(n) * Kp, size_t(Kp) * 8, E.data(),
                                              size_t(n) * 8, size_t(n) * 8, p->ne, cudaMemcpyHostToDevice, st));
            REACH_CUDA(ctx, cudaStreamSynchronize(st));
        }
    }
    {   // direction chains, one row each: ship the caller's direction matrix as is, orient on the device
        DevBuf ddev, src, sgn;
        REACH_TRY(ddev.alloc(ctx, size_t(n) * ndirs * 8, false));
        REACH_TRY(src.alloc(ctx, size_t(p->mu) * 4, false));
        REACH_TRY(sgn.alloc(ctx, size_t(p->mu) * 8, false));
        REACH_CUDA(ctx, cudaMemcpyAsync(ddev.p, dirs, size_t(n) * ndirs * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaMemcpyAsync(src.p, chain_src.data(), size_t(p->mu) * 4, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaMemcpyAsync(sgn.p, chain_sign.data(), size_t(p->mu) * 8, cudaMemcpyHostToDevice, st));
        build_chains_kernel<<<p->mu, 128, 0, st>>>(ddev.as<double>(), n, src.as<int>(),
                                                                          sgn.as<double>(), p->L0.as<double>(), Kp);
        REACH_CUDA(ctx, cudaGetLastError());
        REACH_CUDA(ctx, cudaMemcpyAsync(p->row_pos.p, row_pos.data(), sizeof(int) * p->mu, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaMemcpyAsync(p->row_neg.p, row_neg.data(), sizeof(int) * p->mu, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }
    {   // per-row feature weights and the per-tile activity masks
        std::vector<double> wl(size_t(std::max(1, nq)) * nbe, 0.0), wa(size_t(std::max(1, nq)) * nbe, 0.0);
        std::vector<uint32_t> mask(p->tpb, 0u);
        for (int q = 0; q < nq; ++q) {
            const Feature& f = p->prog.feats[q];
            for (int i = 0; i < n; ++i) { wl[size_t(q) * nbe + i] = f.a[i]; wa[size_t(q) * nbe + i] = f.b[i]; }
            for (auto& kv : f.rows) {
                wl[size_t(q) * nbe + n + kv.first] = kv.second.first;
                wa[size_t(q) * nbe + n + kv.first] = kv.second.second;
            }
            for (int r = 0; r < nbe && !p->packed; ++r)
                if (wl[size_t(q) * nbe + r] != 0.0 || wa[size_t(q) * nbe + r] != 0.0) mask[r / TN] |= 1u << q;
        }
        REACH_CUDA(ctx, cudaMemcpyAsync(p->Wl.p, wl.data(), wl.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaMemcpyAsync(p->Wa.p, wa.data(), wa.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaMemcpyAsync(p->tile_mask.p, mask.data(), mask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }

    // ---- invariant -> dual certificates over the template's support values (see first_disjoint_kernel) ----
    p->nchecks = 0;
    p->inv_faces = p->inv_covered = p->inv_exact = 0;
    if (inv && inv->kind != REACH_INV_NONE) {
        struct Face { std::vector<double> a; double b; };
        std::vector<Face> faces;
        if (inv->kind == REACH_INV_HALFSPACE) {
            if (!inv->a || !inv->b) return set_error(ctx, REACH_EINVAL, "half-space invariant needs a and b");
            faces.push_back({std::vector<double>(inv->a, inv->a + n), inv->b[0]});
        } else if (inv->kind == REACH_INV_BOX) {
            if (!inv->a || !inv->b) return set_error(ctx, REACH_EINVAL, "box invariant needs lo and hi");
            for (int i = 0; i < n; ++i) {
                if (inv->a[i] > inv->b[i]) return set_error(ctx, REACH_EINVAL, "box invariant: lo[%d] > hi[%d]", i, i);
                if (std::isfinite(inv->b[i])) { Face f{std::vector<double>(n, 0.0), inv->b[i]}; f.a[i] = 1.0; faces.push_back(std::move(f)); }
                if (std::isfinite(inv->a[i])) { Face f{std::vector<double>(n, 0.0), -inv->a[i]}; f.a[i] = -1.0; faces.push_back(std::move(f)); }
            }
        } else if (inv->kind == REACH_INV_POLYHEDRON) {
            if (!inv->a || !inv->b || inv->m < 1) return set_error(ctx, REACH_EINVAL, "polyhedral invariant needs m >= 1 faces, A and b");
            for (int f = 0; f < inv->m; ++f)
                faces.push_back({std::vector<double>(inv->a + size_t(f) * n, inv->a + size_t(f + 1) * n), inv->b[f]});
        } else {
            return set_error(ctx, REACH_EUNSUPPORTED, "invariant kind %d is not supported on the device", inv->kind);
        }
        // axis rows of the template: d_j = c e_i
        std::vector<int> pos_row(n, -1), neg_row(n, -1);
        std::vector<double> pos_c(n, 0.0), neg_c(n, 0.0);
        bool pure_box = true;
        for (int j = 0; j < ndirs; ++j) {
            const double* d = dirs + size_t(j) * n;
            int nz = 0, at = -1;
            for (int i = 0; i < n; ++i) if (d[i] != 0.0) { ++nz; at = i; }
            if (nz != 1) { pure_box = false; continue; }
            if (d[at] > 0 && pos_row[at] < 0) { pos_row[at] = j; pos_c[at] = d[at]; }
            if (d[at] < 0 && neg_row[at] < 0) { neg_row[at] = j; neg_c[at] = -d[at]; }
        }
        for (int i = 0; i < n && pure_box; ++i) pure_box = pos_row[i] >= 0 && neg_row[i] >= 0;
        std::vector<int> ptr{0}, rows;
        std::vector<double> w, bound;
        int covered = 0, exact_faces = 0;
        for (const Face& f : faces) {
            int at = -1, nnz = 0;
            for (int i = 0; i < n; ++i) {
                if (!std::isfinite(f.a[i])) return set_error(ctx, REACH_EINVAL, "invariant normal is not finite");
                if (f.a[i] != 0.0) { if (at < 0) at = i; ++nnz; }
            }
            if (at < 0 || std::isnan(f.b)) return set_error(ctx, REACH_EINVAL, "invariant face with a zero normal or NaN offset");
            if (f.b == INFINITY) { ++covered; ++exact_faces; continue; }          // never violated
            bool done = false;
            // (1) -a is a positive multiple of a template direction: d_j = -lam a
            for (int j = 0; j < ndirs && !done; ++j) {
                const double* d = dirs + size_t(j) * n;
                const double lam = -d[at] / f.a[at];
                if (!(lam > 0)) continue;
                bool same = true;
                for (int i = 0; i < n && same; ++i) same = (d[i] == -lam * f.a[i]);
                if (!same) continue;
                rows.push_back(j); w.push_back(1.0 / lam);
                done = true;
                ++exact_faces;
            }
            // (2) the +-e_i rows of the template: -a = sum_i max(-a_i,0) e_i + max(a_i,0) (-e_i)
            if (!done) {
                bool ok = true;
                for (int i = 0; i < n && ok; ++i)
                    if (f.a[i] > 0) ok = neg_row[i] >= 0; else if (f.a[i] < 0) ok = pos_row[i] >= 0;
                if (ok) {
                    for (int i = 0; i < n; ++i) {
                        if (f.a[i] > 0) { rows.push_back(neg_row[i]); w.push_back(f.a[i] / neg_c[i]); }
                        else if (f.a[i] < 0) { rows.push_back(pos_row[i]); w.push_back(-f.a[i] / pos_c[i]); }
                    }
                    done = true;
                    if (pure_box || nnz == 1) ++exact_faces;
                }
            }
            if (!done) continue;                                                  // no certificate for this face
            ptr.push_back(int(rows.size()));
            bound.push_back(-f.b);
            ++covered;
        }
        p->inv_faces = int(faces.size());
        p->inv_covered = covered;
        // equivalent to the reference's _isdisjoint(X, set(F[k])) (an exact polyhedral test): one face with an exact
        // certificate; several faces only when both sets are boxes (separable axis by axis)
        p->inv_exact = (covered == p->inv_faces && exact_faces == p->inv_faces &&
                        (p->inv_faces == 1 || (inv->kind == REACH_INV_BOX && pure_box))) ? 1 : 0;
        p->nchecks = int(bound.size());
        if (covered == 0)
            return set_error(ctx, REACH_EUNSUPPORTED,
                             "the invariant cannot be tested on the device with this template: no face normal is a template "
                             "direction and the template lacks the +-e_i rows its normals need (the reference's exact test is an LP "
                             "per step); add the invariant's normals to the template or check disjointness on the fetched values");
        if (p->nchecks) {
            auto ship = [&](DevBuf& b, const void* src, size_t bytes) -> int {
                REACH_TRY(b.alloc(ctx, bytes, false));
                REACH_CUDA(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
                return REACH_OK;
            };
            REACH_TRY(ship(p->chk_ptr, ptr.data(), ptr.size() * sizeof(int)));
            REACH_TRY(ship(p->chk_rows, rows.data(), rows.size() * sizeof(int)));
            REACH_TRY(ship(p->chk_w, w.data(), w.size() * sizeof(double)));
            REACH_TRY(ship(p->chk_bound, bound.data(), bound.size() * sizeof(double)));
            REACH_CUDA(ctx, cudaStreamSynchronize(st));
            if (!p->fb_host) {
                REACH_CUDA(ctx, cudaMallocHost(&p->fb_host, 2 * sizeof(unsigned long long)));
                for (auto& e : p->ev_fb) REACH_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
        }
    }
    p->uploaded = true;
    p->ran = false;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_set_output(reach_lgg09_plan* p, void* rho_device) {
    if (!p) return REACH_EINVAL;
    if (!rho_device) return set_error(p->ctx, REACH_EINVAL, "rho_device is NULL");
    p->rho = static_cast<double*>(rho_device);
    p->rho_own.release();
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_run(reach_lgg09_plan* p) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->uploaded) return set_error(ctx, REACH_ESTATE, "plan_run before plan_upload");
    NvtxRange nvtx_fn("reach_lgg09_plan_run");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = p->n, Kp = p->Kp, nbe = p->nbe, S = p->S, TM = p->TM, TN = p->TN, mu_pad = p->mu_pad;
    double* stack = p->stack.as<double>();
    auto block = [&](int b) { return stack + size_t(b) * nbe * Kp; };
    p->launches = 0;
    p->gemm_total = 0;
    p->gemm_sampled = 0;

    REACH_CUDA(ctx, cudaEventRecord(p->ev0, st));
    p->synced = false;
    p->steps_computed = p->nsteps;
    const InvChecks ichk{p->chk_ptr.as<int>(), p->chk_rows.as<int>(), p->chk_w.as<double>(), p->chk_bound.as<double>(), p->nchecks};
    // first step in [k0, k1) whose reach-set is certified disjoint from the invariant (atomicMin into first_bad)
    auto disjoint_scan = [&](long long k0, long long k1, cudaStream_t s) -> int {
        if (p->nchecks <= 0 || k1 <= k0) return REACH_OK;
        const int th = 128;
        first_disjoint_kernel<<<unsigned((k1 - k0 + th - 1) / th), th, 0, s>>>(p->rho, p->ndirs, k0, k1, ichk,
                                                                                p->first_bad.as<unsigned long long>());
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
        return REACH_OK;
    };
    if (p->nchecks > 0) REACH_CUDA(ctx, cudaMemsetAsync(p->first_bad.p, 0xFF, sizeof(unsigned long long), st));

    if (p->use_compact) {
        Lgg09Compact c{};
        c.prog = p->prog.dev;
        c.mu = p->mu;
        c.mu_g = p->cp_mu_g;
        c.nq = p->nq;
        c.ne = p->ne;
        c.nblk = p->cp_nblk;
        c.need_next = p->prog.need_next ? 1 : 0;
        c.ndirs = p->ndirs;
        c.nsteps = p->nsteps;
        c.sblk = p->cp_sblk;
        c.log2_sblk = p->cp_log2;
        c.M = p->cpM.as<double>();
        c.r0 = p->cpr0.as<double>();
        c.wl = p->cpwl.as<double>();
        c.wa = p->cpwa.as<double>();
        c.E = p->cpE.as<double>();
        c.WlE = p->cpWlE.as<double>();
        c.WaE = p->cpWaE.as<double>();
        c.slot_used = p->cp_slot_used;
        c.rstart = p->cprstart.as<double>();
        c.voff = p->cpvoff.as<double>();
        c.row_pos = p->row_pos.as<int>();
        c.row_neg = p->row_neg.as<int>();
        c.rho = p->rho;
        REACH_TRY(launch_compact(ctx, p->cp_G, p->cp_ne, c, p->prog.dev.qV >= 0, st, p->launches));
        REACH_TRY(disjoint_scan(0, p->nsteps, st));
        REACH_CUDA(ctx, cudaEventRecord(p->ev1, st));
        p->ran = true;
        return REACH_OK;
    }
    if (p->use_chain) {
        Lgg09Chain c{};
        c.prog = p->prog.dev;
        c.n = n;
        c.npad = p->ch_rpl * CHAIN_THREADS;
        c.maxnnz = p->ch_maxnnz;
        c.ell_val = p->ell_val.as<double>();
        c.ell_col = p->ell_col.as<int>();
        c.Wl = p->Wl.as<double>();
        c.Wa = p->Wa.as<double>();
        c.wstride = nbe;
        c.nq = p->nq;
        c.ne = p->ne;
        c.E = p->Edev.as<double>();
        c.L0 = p->L0.as<double>();
        c.ldl = Kp;
        c.mu = p->mu;
        c.nsteps = p->nsteps;
        c.need_next = p->prog.need_next ? 1 : 0;
        c.row_pos = p->row_pos.as<int>();
        c.row_neg = p->row_neg.as<int>();
        c.rho = p->rho;
        c.ndirs = p->ndirs;
        DevBuf dbgbuf;
        if (getenv("REACH_CHAIN_DEBUG")) { REACH_TRY(dbgbuf.alloc(ctx, 64)); c.dbg = dbgbuf.as<long long>(); }
        REACH_TRY(launch_chain(ctx, p->ch_rpl, p->ch_ne, c, p->ch_smem, st));
        if (c.dbg) {
            long long h[4];
            REACH_CUDA(ctx, cudaStreamSynchronize(st));
            REACH_CUDA(ctx, cudaMemcpy(h, c.dbg, sizeof(h), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[chain dbg] cycles per step: spmv %.0f, reduce %.0f, eval %.0f, prefix/out %.0f\n",
                    double(h[0]) / p->nsteps, double(h[1]) / p->nsteps, double(h[2]) / p->nsteps, double(h[3]) / p->nsteps);
        }
        ++p->launches;
        REACH_TRY(disjoint_scan(0, p->nsteps, st));
        REACH_CUDA(ctx, cudaEventRecord(p->ev1, st));
        p->ran = true;
        return REACH_OK;
    }

    // state reset so a plan can be re-run
    REACH_CUDA(ctx, cudaMemcpyAsync(p->Lbuf[0].p, p->L0.p, size_t(mu_pad) * Kp * sizeof(double), cudaMemcpyDeviceToDevice, st));
    REACH_CUDA(ctx, cudaMemsetAsync(p->s_pos.p, 0, p->s_pos.bytes, st));
    REACH_CUDA(ctx, cudaMemsetAsync(p->s_neg.p, 0, p->s_neg.bytes, st));

    // ---- row stack: block_b = [I;E] (Phi')^b, b = 0..S.  block_0 was uploaded. ----
    nvtxRangePushA("reach_lgg09: row stack (powers of Phi')");
    const bool stack_prebuilt = p->stack_ready;          // built by reach_lgg09_plan_stack_step calls since the last run
    p->stack_ready = false;
    p->stack_steps_done = 0;
    if (stack_prebuilt)                                  // the gathered row chunks may have spilled into the slack rows
        REACH_CUDA(ctx, cudaMemsetAsync(block(S + 1), 0, size_t(256) * Kp * sizeof(double), st));
    if (!stack_prebuilt) {   // Q_1 = Phi row-major: Q[k][l] = Phi(k,l) = phi_cm[k + n l]
        dim3 grid((n + 31) / 32, (n + 31) / 32), blk(32, 8);
        transpose_kernel<<<grid, blk, 0, st>>>(p->phi_cm.as<double>(), n, p->Qbuf[0].as<double>(), Kp, n);
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
    }
    // stack products use the warp-specialised 128x128 kernel when the blocks are tile-aligned and big
    const bool big = (p->packed || nbe % 128 == 0) && n >= 512;
    auto plain_gemm = [&](const double* X, int mrows, const double* Y, int yrows, double* out, int m_valid) -> int {
        const int t = big ? 128 : 64;
        GemmOperands g{X, Y, Kp, Kp, Kp, round_up(mrows, t) / t, round_up(yrows, t) / t};
        StoreEpi epi{out, Kp, m_valid, n};
        REACH_CUDA(ctx, (launch_cfg(big ? 0 : 1, g, epi, st)));
        ++p->launches;
        return REACH_OK;
    };
    // block_1 = block_0 * P_1   (Y[k][l] = P_1[l][k] = Phi(k,l) = Q_1)
    if (!stack_prebuilt) REACH_TRY(plain_gemm(block(0), nbe, p->Qbuf[0].as<double>(), n, block(1), nbe));
    int qcur = 0;
    for (int s = 1; s < S && !stack_prebuilt; s *= 2) {
        const int cnt = std::min(s, S - s);              // blocks s+1 .. s+cnt
        // Q_{2s} = Q_s Q_s : X = Q_s rows, Y[b][l] = Q_s[l][b] = P_s[b][l] = first n rows of block_s
        if (2 * s < S)
            REACH_TRY(plain_gemm(p->Qbuf[qcur].as<double>(), n, block(s), n, p->Qbuf[1 - qcur].as<double>(), n));
        // blocks s+1..s+cnt = blocks 1..cnt * P_s : Y = Q_s
        REACH_TRY(plain_gemm(block(1), cnt * nbe, p->Qbuf[qcur].as<double>(), n, block(s + 1), cnt * nbe));
        qcur ^= 1;
    }

    // digit planes (oz_gemm.cuh) or residue images (crt_gemm.cuh) of FP64 rows; the GEMM over them
    auto oz_slice = [&](const double* X, int rows_valid, int rows_padded, int ksplit, int tiles_total, int8_t* Q, int* e, int tile0,
                        bool is_y) {
        if (p->oz_crt)
            launch_crt_slice(X, Kp, rows_valid, rows_padded, n, p->crt, is_y, p->oz_nkb, tiles_total, Q, e, tile0, ksplit, st);
        else
            oz_slice_kernel<OZ_TM><<<dim3(rows_padded / 8, ksplit), 256, 0, st>>>(X, Kp, rows_valid, n, p->oz_T, p->oz_nkb, tiles_total, Q,
                                                                               e, tile0);
    };
    auto oz_launch = [&](const OzOperands& g, const auto& e) -> cudaError_t {
        if (p->oz_crt) {
            CrtOperands c{g.Xq, g.Yq, g.eX, g.eY, g.nkb, g.tiles_m, g.tiles_m_total, g.tile_n0, g.tiles_n, g.tiles_n_total,
                          p->crt_scr.as<uint32_t>(), p->crt};
            static const int dbg_flags = getenv("REACH_CRT_FLAGS") ? atoi(getenv("REACH_CRT_FLAGS")) : 0;   // timing experiments
            c.flags = dbg_flags;
            static const bool want_tim = getenv("REACH_CRT_TIMING") != nullptr;
            if (want_tim) {      // cycle counters of the warp roles of this launch, printed after a blocking copy (debug only)
                static unsigned long long* tim = nullptr;
                if (!tim) cudaMalloc(&tim, size_t(256) * 16 * 8);
                cudaMemsetAsync(tim, 0, size_t(256) * 16 * 8, st);
                c.tim = tim;
                cudaError_t le = launch_crt_gemm(c, e, st);
                std::vector<unsigned long long> h(size_t(256) * 16);
                cudaMemcpyAsync(h.data(), tim, h.size() * 8, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                double a[16] = {0};
                int nb = 0;
                for (int b = 0; b < 256; b += 2) {       // leaders only
                    if (!h[size_t(b) * 16 + 12]) continue;
                    ++nb;
                    for (int k = 0; k < 16; ++k) a[k] += double(h[size_t(b) * 16 + k]);
                }
                for (int k = 0; k < 16; ++k) a[k] /= std::max(1, nb) * 1e3;
                fprintf(stderr, "[crt timing, kcycles avg over %d leader CTAs] mma: wait tmem_empty %.0f, wait full %.0f, wait pfull %.0f, total %.0f | "
                        "drain(t0): wait scr_free %.0f, wait tmem_full %.0f, work %.0f | recon(t256): wait slice %.0f, sums %.0f, accept %.0f, finish %.0f, total %.0f\n",
                        nb, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[8], a[9], a[10], a[11], a[12]);
                return le;
            }
            return launch_crt_gemm(c, e, st);
        }
        return launch_oz_gemm(p->oz_T, g, e, st);
    };
    if (p->use_oz) {   // digit planes of the whole row stack (once per run)
        oz_slice(stack, (S + 1) * nbe, p->oz_tiles_n * OZ_TN, 1, p->oz_tiles_n, p->Yq.as<int8_t>(), p->eY.as<int>(), 0, true);
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
    }
    const int J = p->use_oz ? p->oz_J : 1;
    const int mu_ld = p->mu_ld;
    const int mt = mu_pad / OZ_TM;                      // row tiles of one pass's direction rows (digit-plane path)
    if (J > 1) {
        // carry stack: C_j = (Phi')^(jS) = first n rows of block jS, j = 1..J.  C_1 is a copy of block S, C_{j+1} = C_j x P_S
        // with Y = Q_S = the transpose of C_1 (the same NT product as the stack build).
        const size_t cblk = size_t(p->oz_ctiles) * OZ_TN * Kp;          // doubles per carry block
        double* C = p->cstack.as<double>();
        REACH_CUDA(ctx, cudaMemsetAsync(C, 0, p->cstack.bytes, st));
        REACH_CUDA(ctx, cudaMemcpy2DAsync(C, size_t(Kp) * 8, block(S), size_t(Kp) * 8, size_t(Kp) * 8, n, cudaMemcpyDeviceToDevice, st));
        double* QS = p->Qbuf[0].as<double>();
        {
            dim3 grid((n + 31) / 32, (n + 31) / 32), blk(32, 8);
            REACH_CUDA(ctx, cudaMemsetAsync(QS, 0, p->Qbuf[0].bytes, st));
            transpose_kernel<<<grid, blk, 0, st>>>(C, Kp, QS, Kp, n);
            REACH_CUDA(ctx, cudaGetLastError());
            ++p->launches;
        }
        for (int j = 1; j < J; ++j) REACH_TRY(plain_gemm(C + (j - 1) * cblk, n, QS, n, C + j * cblk, n));
        oz_slice(C, J * p->oz_ctiles * OZ_TN, J * p->oz_ctiles * OZ_TN, 1, J * p->oz_ctiles, p->Yq_c.as<int8_t>(), p->eY_c.as<int>(), 0, true);
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
    }
    nvtxRangePop();
    NvtxRange nvtx_main("reach_lgg09: main loop (GEMM passes + combine)");
    OzLgg09Epi<4> oepi{};
    oepi.ldl = Kp;
    oepi.n = n;
    oepi.nbe = nbe;
    oepi.Wl = p->Wl.as<double>();
    oepi.Wa = p->Wa.as<double>();
    oepi.nq = p->nq;
    oepi.featpart = p->featpart.as<double>();
    oepi.mu_pad = mu_ld;
    // one pass of the digit-plane GEMM: L (fp64) -> digit planes, then blocks b_lo..S of the stack
    auto oz_pass = [&](const double* Lcur, double* Lnext, int b_lo, int b_hi, int store_block, bool sample) -> int {
        // a few hundred chains are only mu_pad/8 row groups: split the k range so that ~2 CTAs per SM are in flight
        const int ksplit = std::max(1, std::min(8, (2 * ctx->sm_count) / std::max(1, mu_pad / 8)));
        oz_slice(Lcur, p->mu, mu_pad, ksplit, J * mt, p->Xq.as<int8_t>(), p->eX.as<int>(), 0, false);
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
        const int tile_lo = int((size_t(b_lo) * nbe) / OZ_TN);
        const int tile_hi = int((size_t(b_hi + 1) * nbe + OZ_TN - 1) / OZ_TN);
        OzOperands g{p->Xq.as<int8_t>(), p->Yq.as<int8_t>(), p->eX.as<int>(), p->eY.as<int>(), p->oz_nkb,
                     mt, J * mt, tile_lo, tile_hi - tile_lo, p->oz_tiles_n};
        oepi.Lnext = Lnext;
        oepi.store_block = store_block;
        if (sample) REACH_CUDA(ctx, cudaEventRecord(p->gemm_ev[2 * p->gemm_sampled], st));
        REACH_CUDA(ctx, (oz_launch(g, oepi)));
        if (sample) {
            REACH_CUDA(ctx, cudaEventRecord(p->gemm_ev[2 * p->gemm_sampled + 1], st));
            ++p->gemm_sampled;
        }
        ++p->launches;
        return REACH_OK;
    };

    // A group of Jg passes t .. t+Jg-1 from the direction rows L_t (see reach_lgg09_plan::oz_J): carry GEMM -> L_{t+1..t+Jg},
    // digit planes of [L_t; ...; L_{t+Jg-1}], ONE output GEMM over blocks b_lo..S for all Jg passes (features only: nothing
    // is carried out of it), and L_{t+Jg} becomes the plan's current direction rows.
    auto oz_group = [&](double* Lcur, int Jg, int b_lo, bool sample) -> int {
        const int ksplit = std::max(1, std::min(8, (2 * ctx->sm_count) / std::max(1, mu_pad / 8)));
        oz_slice(Lcur, p->mu, mu_pad, ksplit, J * mt, p->Xq.as<int8_t>(), p->eX.as<int>(), 0, false);
        REACH_CUDA(ctx, cudaGetLastError());
        OzOperands gc{p->Xq.as<int8_t>(), p->Yq_c.as<int8_t>(), p->eX.as<int>(), p->eY_c.as<int>(), p->oz_nkb,
                      mt, J * mt, 0, Jg * p->oz_ctiles, J * p->oz_ctiles};
        OzCarryEpi cepi{p->Lst.as<double>(), Kp, n, p->oz_ctiles, mu_pad};
        REACH_CUDA(ctx, (oz_launch(gc, cepi)));
        if (Jg > 1) {
            const int rows = (Jg - 1) * mu_pad;
            const int ks2 = std::max(1, std::min(8, (2 * ctx->sm_count) / std::max(1, rows / 8)));
            oz_slice(p->Lst.as<double>(), rows, rows, ks2, J * mt, p->Xq.as<int8_t>(), p->eX.as<int>(), mt, false);
            REACH_CUDA(ctx, cudaGetLastError());
            ++p->launches;
        }
        // the next group's (or pass's) direction rows
        REACH_CUDA(ctx, cudaMemcpyAsync(Lcur, p->Lst.as<double>() + size_t(Jg - 1) * mu_pad * Kp, size_t(mu_pad) * Kp * 8,
                                        cudaMemcpyDeviceToDevice, st));
        const int tile_lo = int((size_t(b_lo) * nbe) / OZ_TN);
        const int tile_hi = int((size_t(S + 1) * nbe + OZ_TN - 1) / OZ_TN);
        OzOperands g{p->Xq.as<int8_t>(), p->Yq.as<int8_t>(), p->eX.as<int>(), p->eY.as<int>(), p->oz_nkb,
                     Jg * mt, J * mt, tile_lo, tile_hi - tile_lo, p->oz_tiles_n};
        oepi.Lnext = nullptr;
        oepi.store_block = -1;
        if (sample) REACH_CUDA(ctx, cudaEventRecord(p->gemm_ev[2 * p->gemm_sampled], st));
        REACH_CUDA(ctx, (oz_launch(g, oepi)));
        if (sample) {
            REACH_CUDA(ctx, cudaEventRecord(p->gemm_ev[2 * p->gemm_sampled + 1], st));
            ++p->gemm_sampled;
        }
        p->launches += 3;
        return REACH_OK;
    };

    // ---- main loop over super-steps ----
    const int64_t Kmax = p->nsteps - 1 + (p->prog.need_next ? 1 : 0);
    const int64_t T = (Kmax + S - 1) / S;
    Lgg09Epi epi{};
    epi.ldl = Kp;
    epi.n = n;
    epi.nbe = nbe;
    epi.Wl = p->Wl.as<double>();
    epi.Wa = p->Wa.as<double>();
    epi.tile_mask = p->tile_mask.as<uint32_t>();
    epi.nq = p->nq;
    epi.featpart = p->featpart.as<double>();
    epi.mu_pad = mu_ld;

    Lgg09PackedEpi pepi{};
    pepi.ldl = Kp;
    pepi.n = n;
    pepi.nbe = nbe;
    pepi.Wl = p->Wl.as<double>();
    pepi.Wa = p->Wa.as<double>();
    pepi.nq = p->nq;
    pepi.featpart = p->featpart.as<double>();
    pepi.mu_pad = mu_ld;

    Lgg09Combine cmb{};
    cmb.prog = p->prog.dev;
    cmb.featpart = p->featpart.as<double>();
    cmb.tile_mask = p->tile_mask.as<uint32_t>();
    cmb.tpb = p->tpb;
    cmb.nq = p->nq;
    cmb.mu_pad = mu_ld;
    cmb.mu = p->mu;
    cmb.nsteps = p->nsteps;
    cmb.s_pos = p->s_pos.as<double>();
    cmb.s_neg = p->s_neg.as<double>();
    cmb.row_pos = p->row_pos.as<int>();
    cmb.row_neg = p->row_neg.as<int>();
    cmb.rho = p->rho;
    cmb.ndirs = p->ndirs;
    const int cthreads = 64, cblocks = (p->mu + cthreads - 1) / cthreads;
    double* ftot = p->ftot.as<double>();
    int carry_cur = 0;
    bool oz_now = p->use_oz;      // false after the guard has sent the run back to the DMMA kernel
    auto combine = [&](const Lgg09Combine& c, cudaStream_t st) -> int {
        const int nblk = c.b_hi - c.b_lo + 1;
        if (nblk > 0) {
            dim3 grid(cblocks, nblk);
            if (p->packed) lgg09_tilesum_packed_kernel<<<grid, cthreads, 0, st>>>(c, c.featpart, ftot, nbe, TN,
                                                                                  oz_now ? 0 : c.b_lo);
            else lgg09_tilesum_kernel<<<grid, cthreads, 0, st>>>(c, ftot);
            REACH_CUDA(ctx, cudaGetLastError());
            ++p->launches;
        }
        const int npair = c.has_carry ? nblk : (nblk > 0 ? nblk - 1 : 0);
        const int nslots = npair + (c.tail ? 1 : 0);
        if (nslots <= 0) {
            // nothing to emit (cannot happen for the passes issued below, kept for safety)
            return REACH_OK;
        }
        const double* cin = p->carry[carry_cur].as<double>();
        double* cout = p->carry[1 - carry_cur].as<double>();
        double* dv = p->dv.as<double>();
        if (c.prog.qV >= 0) {
            // running input sums s_k of the steps this pass emits (first step: kfirst), before the evaluation
            const long long kfirst = c.has_carry ? c.k0 + c.b_lo - 1 : (nblk > 0 ? c.k0 + c.b_lo : c.k0);
            lgg09_prefix_kernel<<<(p->mu + PREFIX_CB - 1) / PREFIX_CB, PREFIX_THREADS, 0, st>>>(c, ftot, cin, dv, nslots, kfirst);
            REACH_CUDA(ctx, cudaGetLastError());
            ++p->launches;
        }
        dim3 grid(cblocks, nslots);
        if (c.nq <= 4) lgg09_eval_kernel<4><<<grid, cthreads, 0, st>>>(c, ftot, cin, cout, dv);
        else if (c.nq <= 8) lgg09_eval_kernel<8><<<grid, cthreads, 0, st>>>(c, ftot, cin, cout, dv);
        else lgg09_eval_kernel<LGG09_MAX_FEATS><<<grid, cthreads, 0, st>>>(c, ftot, cin, cout, dv);
        REACH_CUDA(ctx, cudaGetLastError());
        ++p->launches;
        carry_cur ^= 1;
        return REACH_OK;
    };

    // sample at most 512 main-loop GEMM launches with event pairs
    const int64_t sample_every = std::max<int64_t>(1, (T + 511) / 512);
    const size_t want_ev = size_t(2) * size_t((T + sample_every - 1) / sample_every + 3);
    while (p->gemm_ev.size() < want_ev) {
        cudaEvent_t e;
        REACH_CUDA(ctx, cudaEventCreate(&e));
        p->gemm_ev.push_back(e);
    }

    int cur = 0;
    if (T == 0) {
        // NSTEPS == 1 and nothing is evaluated at r_1: only the features of r_0 are needed
        GemmOperands g{p->Lbuf[cur].as<double>(), block(0), Kp, Kp, Kp, mu_pad / TM, (nbe + TN - 1) / TN};
        epi.Lnext = p->Lbuf[1 - cur].as<double>();
        epi.block0 = 0;
        epi.store_block = -1;
        if (p->use_oz) {
            REACH_TRY(oz_pass(p->Lbuf[cur].as<double>(), epi.Lnext, 0, 0, -1, false));
            --p->launches;
        } else if (p->packed) {
            pepi.Lnext = epi.Lnext; pepi.slice_block0 = 0; pepi.store_block = -1;
            REACH_CUDA(ctx, (launch_cfg(p->cfg, g, pepi, st)));
        } else {
            REACH_CUDA(ctx, (launch_cfg(p->cfg, g, epi, st)));
        }
        ++p->launches;
        cmb.b_lo = 0; cmb.b_hi = 0; cmb.has_carry = 0; cmb.tail = 1; cmb.k0 = 0;
        REACH_TRY(combine(cmb, st));
    }
    p->oz_guard_max = 0.0;
    p->oz_drift_max = 0.0;
    p->oz_fallback_pass = -1;
    p->oz_replay_from = -1;
    p->oz_replayed = 0;
    p->oz_checks = 0;
    constexpr double kOzGuardTol = 1e-12;      // one pass, INT8 against DMMA from the same direction rows
    constexpr double kOzDriftTol = 1e-11;      // accumulated since pass 0, INT8 chain against the FP64 sentinel chain
    // first check with crt_lo moduli: above this the run switches to crt_hi (C5: 1.0e-13 with 13 moduli, 6e-15 with 14)
    const double crt_escalate_tol = getenv("REACH_CRT_ESCALATE_TOL") ? atof(getenv("REACH_CRT_ESCALATE_TOL")) : 2.5e-13;
    // test hook: the first check at a pass >= REACH_OZ_TRIP_AT is treated as failed (exercises the replay path)
    const char* trip_env = getenv("REACH_OZ_TRIP_AT");
    int64_t trip_at = trip_env ? atoll(trip_env) : -1;
    // Pass t's combine kernels only read what its GEMM wrote (per-tile partials, double-buffered by parity) and their
    // own carried state, so they run on the side stream beside the GEMM of pass t+1; guard-checked passes stay serial.
    const bool overlap = T > 2 && !getenv("REACH_NO_OVERLAP");
    double* const fp_base = p->featpart.as<double>();
    cudaStream_t st2 = p->st2, st3 = p->st3;
    bool side_busy[2] = {false, false};
    auto join_side = [&]() -> int {
        for (int i = 0; i < 2; ++i)
            if (side_busy[i]) {
                REACH_CUDA(ctx, cudaStreamWaitEvent(st, p->ev_comb[i], 0));
                side_busy[i] = false;
            }
        return REACH_OK;
    };
    // Early exit (reach_inhomog.jl:198-215: `_isdisjoint(X, set(F[k])) && break`): every pass scans the steps it emitted
    // and copies the "first disjoint step" word to pinned host memory; the host reads the copy of pass t-2 before it
    // launches pass t, so it stays two passes ahead of the device (no pipeline stall) and stops launching as soon as a
    // step has been certified disjoint.
    bool fb_pending[2] = {false, false};
    bool stopped = false;
    if (p->nchecks > 0) p->fb_host[0] = p->fb_host[1] = ~0ull;
    auto scan_and_poll = [&](long long k0, long long k1, cudaStream_t s, int slot) -> int {
        if (p->nchecks <= 0) return REACH_OK;
        REACH_TRY(disjoint_scan(k0, std::min<long long>(k1, p->nsteps), s));
        REACH_CUDA(ctx, cudaMemcpyAsync(&p->fb_host[slot], p->first_bad.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        REACH_CUDA(ctx, cudaEventRecord(p->ev_fb[slot], s));
        fb_pending[slot] = true;
        return REACH_OK;
    };
    auto combine_pass = [&](const Lgg09Combine& c, bool side, int par) -> int {
        cudaStream_t cs = side ? st2 : st;
        if (side) {
            REACH_CUDA(ctx, cudaEventRecord(p->ev_gemm[par], st));
            REACH_CUDA(ctx, cudaStreamWaitEvent(st2, p->ev_gemm[par], 0));
        }
        REACH_TRY(combine(c, cs));
        REACH_TRY(scan_and_poll(c.k0, c.k0 + S, cs, par));
        if (side) {
            REACH_CUDA(ctx, cudaEventRecord(p->ev_comb[par], st2));
            side_busy[par] = true;
        }
        return REACH_OK;
    };
    // ---- guard state of the digit-plane path ----
    const int nsen = p->nsen, sen_stride = p->sen_stride;
    bool sen_on = oz_now && p->oz_guard && T > 0 && !getenv("REACH_OZ_NO_SENTINEL");
    int sen_cur = 0;
    int64_t t_snap = 0;                        // the state at the start of pass t_snap is the last verified one
    if (sen_on) {
        // sentinel rows = chains 0, stride, 2 stride, ... of L_0; the sentinel stream starts once the stack is built
        REACH_CUDA(ctx, cudaMemsetAsync(p->Lsh[0].p, 0, p->Lsh[0].bytes, st));
        REACH_CUDA(ctx, cudaMemcpy2DAsync(p->Lsh[0].p, size_t(Kp) * 8, p->L0.p, size_t(sen_stride) * Kp * 8, size_t(Kp) * 8, nsen,
                                          cudaMemcpyDeviceToDevice, st));
        REACH_CUDA(ctx, cudaEventRecord(p->ev_sen[0], st));
        REACH_CUDA(ctx, cudaStreamWaitEvent(st3, p->ev_sen[0], 0));
    }
    // features of blocks 1 and S (the first step of the pass and the block that becomes L_{k+S}) on the DMMA kernel from
    // the rows X; nothing of the run's state is touched (no store block, separate totals buffer)
    // With few chains a 128x128 launch is a fraction of one wave (250 chains: 32 CTAs of 0.3 ms each); 64x64 tiles put four
    // times as many CTAs to work on it.  Same arithmetic (FP64 DMMA), only the tile shape differs.
    auto dmma_features = [&](const double* X, int rows_pad, int nchains, double* totals) -> int {
        const bool small = int64_t(rows_pad / 128) * ((nbe + 127) / 128) * 2 <= 3 * int64_t(ctx->sm_count);
        const int tm = small ? 64 : TM, tn = small ? 64 : TN;
        for (int cb : {1, S}) {
            GemmOperands gc{X, block(cb), Kp, Kp, Kp, rows_pad / tm, (nbe + tn - 1) / tn};
            pepi.Lnext = p->Lbuf[1 - cur].as<double>(); pepi.slice_block0 = cb; pepi.store_block = -1;
            REACH_CUDA(ctx, (launch_cfg(small ? 1 : 0, gc, pepi, st)));
            Lgg09Combine cc = cmb;
            cc.b_lo = cb; cc.b_hi = cb; cc.mu = nchains;
            dim3 grid((nchains + cthreads - 1) / cthreads, 1);
            lgg09_tilesum_packed_kernel<<<grid, cthreads, 0, st>>>(cc, cc.featpart, totals, nbe, tn, cb);
            REACH_CUDA(ctx, cudaGetLastError());
            p->launches += 2;
            if (S == 1) break;
        }
        return REACH_OK;
    };
    // sentinel rows (<= 128) x an n x n power: 64x64 tiles (2 x n/64 CTAs instead of n/128)
    auto sentinel_step = [&](const double* Ypow, int sen_from) -> int {
        GemmOperands gs{p->Lsh[sen_from].as<double>(), Ypow, Kp, Kp, Kp, OZ_TM / 64, round_up(n, 64) / 64};
        StoreEpi sepi{p->Lsh[1 - sen_from].as<double>(), Kp, nsen, n};
        REACH_CUDA(ctx, (launch_cfg(1, gs, sepi, st3)));
        ++p->launches;
        return REACH_OK;
    };
    const int gemm_ev_pairs = int(p->gemm_ev.size() / 2);
    // guarded digit-plane path: passes 0, 1, every 64th and the last are also run on the DMMA kernel
    auto is_check = [&](int64_t tt) { return oz_now && p->oz_guard && (tt < 2 || (tt & 63) == 0 || tt == T - 1); };
    int64_t npl = 0;                               // main-loop GEMM launches so far (a pass, or a group of passes)
    for (int64_t t = 0; t < T; ++t, ++npl) {
        const int b_lo = t == 0 ? 0 : 1;
        const int par = int(npl & 1);
        if (fb_pending[par]) {                     // the invariant scan of pass t-2
            REACH_CUDA(ctx, cudaEventSynchronize(p->ev_fb[par]));
            fb_pending[par] = false;
            if (p->fb_host[par] != ~0ull) {
                stopped = true;
                p->steps_computed = std::min<int64_t>(p->nsteps, t * int64_t(S));
                break;
            }
        }
        {
            double* fp = fp_base + (overlap ? size_t(par) * p->fp_stride : 0);
            oepi.featpart = fp; epi.featpart = fp; pepi.featpart = fp; cmb.featpart = fp;
        }
        GemmOperands g{p->Lbuf[cur].as<double>(), block(b_lo), Kp, Kp, Kp, mu_pad / TM,
                       ((S + 1 - b_lo) * nbe + TN - 1) / TN};
        epi.Lnext = p->Lbuf[1 - cur].as<double>();
        epi.block0 = b_lo;
        epi.store_block = S;
        const bool sample = (t % sample_every) == 0 && p->gemm_sampled < gemm_ev_pairs;
        cmb.b_lo = b_lo; cmb.b_hi = S; cmb.has_carry = t > 0; cmb.tail = 0; cmb.k0 = t * S;
        const bool check = is_check(t);
        const bool side = overlap && !check;
        if (oz_now && J > 1 && !check) {
            // pass group: as many unchecked passes as fit (the checked ones stay single passes, so the guard logic below
            // sees exactly the state it always saw)
            int Jg = 1;
            while (Jg < J && t + Jg < T && !is_check(t + Jg)) ++Jg;
            if (Jg >= 2) {
                if (side && side_busy[par]) {      // the combines of launch npl-2 have read this parity's partials
                    REACH_CUDA(ctx, cudaStreamWaitEvent(st, p->ev_comb[par], 0));
                    side_busy[par] = false;
                }
                REACH_TRY(oz_group(p->Lbuf[cur].as<double>(), Jg, b_lo, sample));
                ++p->gemm_total;
                cudaStream_t cs = side ? st2 : st;
                if (side) {
                    REACH_CUDA(ctx, cudaEventRecord(p->ev_gemm[par], st));
                    REACH_CUDA(ctx, cudaStreamWaitEvent(st2, p->ev_gemm[par], 0));
                }
                for (int j = 0; j < Jg; ++j) {     // the passes of the group, in step order (running sums, carried features)
                    Lgg09Combine c = cmb;
                    c.featpart = cmb.featpart + size_t(j) * mu_pad;
                    c.b_lo = (t + j == 0) ? 0 : 1; c.b_hi = S; c.has_carry = (t + j > 0); c.tail = 0; c.k0 = (t + j) * int64_t(S);
                    REACH_TRY(combine(c, cs));
                }
                if (side) {
                    REACH_CUDA(ctx, cudaEventRecord(p->ev_comb[par], st2));
                    side_busy[par] = true;
                }
                if (sen_on) {                      // sentinel rows Jg passes on: x (Phi')^(Jg S), FP64, own stream
                    REACH_TRY(sentinel_step(p->cstack.as<double>() + size_t(Jg - 1) * p->oz_ctiles * OZ_TN * Kp, sen_cur));
                    sen_cur ^= 1;
                }
                t += Jg - 1;
                continue;
            }
        }
        if (check) {
            REACH_TRY(join_side());
        } else if (side && side_busy[par]) {       // the combine of pass t-2 has read this parity's partials
            REACH_CUDA(ctx, cudaStreamWaitEvent(st, p->ev_comb[par], 0));
            side_busy[par] = false;
        }
        if (check) {
            if (sen_on) {                          // the sentinel rows of pass t, propagated in FP64 since pass 0
                REACH_CUDA(ctx, cudaEventRecord(p->ev_sen[1], st3));
                REACH_CUDA(ctx, cudaStreamWaitEvent(st, p->ev_sen[1], 0));
                REACH_TRY(dmma_features(p->Lsh[sen_cur].as<double>(), OZ_TM, nsen, p->ftot_sen.as<double>()));
            }
            REACH_TRY(dmma_features(p->Lbuf[cur].as<double>(), mu_pad, p->mu, p->ftot_ref.as<double>()));
        }
        if (oz_now) {
            REACH_TRY(oz_pass(p->Lbuf[cur].as<double>(), epi.Lnext, b_lo, S, S, sample));
            ++p->gemm_total;
            REACH_TRY(combine_pass(cmb, side, par));
            if (check) {
                REACH_CUDA(ctx, cudaMemsetAsync(p->guard_val.p, 0, sizeof(unsigned long long), st));
                REACH_CUDA(ctx, cudaMemsetAsync(p->drift_val.p, 0, sizeof(unsigned long long), st));
                lgg09_guard_kernel<<<(p->mu * p->nq + 127) / 128, 128, 0, st>>>(ftot, p->ftot_ref.as<double>(), p->nq, mu_ld, p->mu,
                                                                               1, S, p->guard_val.as<unsigned long long>());
                REACH_CUDA(ctx, cudaGetLastError());
                ++p->launches;
                if (sen_on) {
                    lgg09_guard_kernel<<<(nsen * p->nq + 127) / 128, 128, 0, st>>>(ftot, p->ftot_sen.as<double>(), p->nq, mu_ld, nsen,
                                                                                  1, S, p->drift_val.as<unsigned long long>(), sen_stride);
                    REACH_CUDA(ctx, cudaGetLastError());
                    ++p->launches;
                    // the sentinel stream may overwrite its older row buffer only after the features were taken from it
                    REACH_CUDA(ctx, cudaEventRecord(p->ev_sen[0], st));
                    REACH_CUDA(ctx, cudaStreamWaitEvent(st3, p->ev_sen[0], 0));
                }
                unsigned long long bits[2] = {0, 0};
                REACH_CUDA(ctx, cudaMemcpyAsync(&bits[0], p->guard_val.p, sizeof(bits[0]), cudaMemcpyDeviceToHost, st));
                REACH_CUDA(ctx, cudaMemcpyAsync(&bits[1], p->drift_val.p, sizeof(bits[1]), cudaMemcpyDeviceToHost, st));
                REACH_CUDA(ctx, cudaStreamSynchronize(st));
                double worst, drift;
                std::memcpy(&worst, &bits[0], sizeof(worst));
                std::memcpy(&drift, &bits[1], sizeof(drift));
                ++p->oz_checks;
                if (worst > p->oz_guard_max) p->oz_guard_max = worst;
                if (drift > p->oz_drift_max) p->oz_drift_max = drift;
                bool ok = worst <= kOzGuardTol && drift <= kOzDriftTol;          // NaN fails
                if (trip_at >= 0 && t >= trip_at) { ok = false; trip_at = -1; }
                if (t == 0 && p->crt_hi > p->oz_crt && !(worst <= crt_escalate_tol)) {
                    // marginal (or failed) with crt_lo moduli: take crt_hi from here on.  Nothing has been emitted that is kept:
                    // new residue images of the stack, state back to the start of pass 0 (the sentinel chains have not moved yet)
                    p->crt_escalated_at = worst;
                    p->oz_crt = p->oz_T = p->crt_hi;
                    p->crt = crt_make_consts(p->crt_hi);
                    p->oz_guard_max = 0.0;
                    p->oz_drift_max = 0.0;
                    oz_slice(stack, (S + 1) * nbe, p->oz_tiles_n * OZ_TN, 1, p->oz_tiles_n, p->Yq.as<int8_t>(), p->eY.as<int>(), 0, true);
                    if (J > 1)
                        oz_slice(p->cstack.as<double>(), J * p->oz_ctiles * OZ_TN, J * p->oz_ctiles * OZ_TN, 1, J * p->oz_ctiles,
                                 p->Yq_c.as<int8_t>(), p->eY_c.as<int>(), 0, true);
                    REACH_CUDA(ctx, cudaGetLastError());
                    p->launches += 2;
                    REACH_CUDA(ctx, cudaMemcpyAsync(p->Lbuf[0].p, p->L0.p, size_t(mu_pad) * Kp * 8, cudaMemcpyDeviceToDevice, st));
                    REACH_CUDA(ctx, cudaMemsetAsync(p->s_pos.p, 0, p->s_pos.bytes, st));
                    REACH_CUDA(ctx, cudaMemsetAsync(p->s_neg.p, 0, p->s_neg.bytes, st));
                    cur = 0;
                    carry_cur = 0;
                    if (p->nchecks > 0) {
                        REACH_CUDA(ctx, cudaMemsetAsync(p->first_bad.p, 0xFF, sizeof(unsigned long long), st));
                        fb_pending[0] = fb_pending[1] = false;
                        p->fb_host[0] = p->fb_host[1] = ~0ull;
                    }
                    t = -1;                        // the loop increment makes it 0
                    continue;
                }
                if (ok) {
                    // verified: keep the state at the start of pass t + 1 (L_{t+1}, running sums, carried features)
                    REACH_CUDA(ctx, cudaMemcpyAsync(p->L_snap.p, p->Lbuf[1 - cur].p, size_t(mu_pad) * Kp * 8, cudaMemcpyDeviceToDevice, st));
                    REACH_CUDA(ctx, cudaMemcpyAsync(p->s_snap.p, p->s_pos.p, size_t(mu_pad) * 8, cudaMemcpyDeviceToDevice, st));
                    REACH_CUDA(ctx, cudaMemcpyAsync(p->s_snap.as<double>() + mu_pad, p->s_neg.p, size_t(mu_pad) * 8, cudaMemcpyDeviceToDevice, st));
                    REACH_CUDA(ctx, cudaMemcpyAsync(p->carry_snap.p, p->carry[carry_cur].p, p->carry_snap.bytes, cudaMemcpyDeviceToDevice, st));
                    t_snap = t + 1;
                } else {
                    // Not FP64-faithful on this problem (rows with a wide dynamic range), or drifting: everything since
                    // the last verified pass is unverified INT8 output.  Restore that state and run passes t_snap .. T-1
                    // on the FP64 tensor pipe (their support values overwrite the INT8 ones).
                    oz_now = false;
                    sen_on = false;
                    p->oz_fallback_pass = t;
                    p->oz_replay_from = t_snap;
                    p->oz_replayed = t - t_snap + 1;
                    if (t_snap == 0) {
                        REACH_CUDA(ctx, cudaMemcpyAsync(p->Lbuf[0].p, p->L0.p, size_t(mu_pad) * Kp * 8, cudaMemcpyDeviceToDevice, st));
                        REACH_CUDA(ctx, cudaMemsetAsync(p->s_pos.p, 0, p->s_pos.bytes, st));
                        REACH_CUDA(ctx, cudaMemsetAsync(p->s_neg.p, 0, p->s_neg.bytes, st));
                    } else {
                        REACH_CUDA(ctx, cudaMemcpyAsync(p->Lbuf[0].p, p->L_snap.p, size_t(mu_pad) * Kp * 8, cudaMemcpyDeviceToDevice, st));
                        REACH_CUDA(ctx, cudaMemcpyAsync(p->s_pos.p, p->s_snap.p, size_t(mu_pad) * 8, cudaMemcpyDeviceToDevice, st));
                        REACH_CUDA(ctx, cudaMemcpyAsync(p->s_neg.p, p->s_snap.as<double>() + mu_pad, size_t(mu_pad) * 8, cudaMemcpyDeviceToDevice, st));
                        REACH_CUDA(ctx, cudaMemcpyAsync(p->carry[0].p, p->carry_snap.p, p->carry_snap.bytes, cudaMemcpyDeviceToDevice, st));
                    }
                    cur = 0;
                    carry_cur = 0;
                    if (p->nchecks > 0) {          // the invariant scans of the discarded passes are void
                        REACH_CUDA(ctx, cudaMemsetAsync(p->first_bad.p, 0xFF, sizeof(unsigned long long), st));
                        REACH_TRY(disjoint_scan(0, t_snap * int64_t(S), st));
                        fb_pending[0] = fb_pending[1] = false;
                        p->fb_host[0] = p->fb_host[1] = ~0ull;
                    }
                    t = t_snap - 1;                // the loop increment makes it t_snap
                    continue;
                }
            }
            if (sen_on) {                          // sentinel rows of pass t + 1 = rows of pass t x block S, FP64, own stream
                REACH_TRY(sentinel_step(block(S), sen_cur));
                sen_cur ^= 1;
            }
            cur ^= 1;
            continue;
        }
        const int dcfg = p->use_oz ? 0 : p->cfg;           // after a fallback: the warp-specialised DMMA kernel
        if (sample) REACH_CUDA(ctx, cudaEventRecord(p->gemm_ev[2 * p->gemm_sampled], st));
        if (p->packed) {
            pepi.Lnext = epi.Lnext; pepi.slice_block0 = b_lo; pepi.store_block = S;
            REACH_CUDA(ctx, (launch_cfg(dcfg, g, pepi, st)));
        } else {
            REACH_CUDA(ctx, (launch_cfg(dcfg, g, epi, st)));
        }
        if (sample) {
            REACH_CUDA(ctx, cudaEventRecord(p->gemm_ev[2 * p->gemm_sampled + 1], st));
            ++p->gemm_sampled;
        }
        ++p->launches;
        ++p->gemm_total;
        REACH_TRY(combine_pass(cmb, side, par));
        cur ^= 1;
    }
    if (p->use_oz && p->oz_guard && T > 0) {       // the sentinel stream reads the stack and the plan's buffers
        REACH_CUDA(ctx, cudaEventRecord(p->ev_sen[1], st3));
        REACH_CUDA(ctx, cudaStreamWaitEvent(st, p->ev_sen[1], 0));
    }
    REACH_TRY(join_side());
    if (!stopped && T > 0 && T * S < p->nsteps) {
        // one step left whose support values only need the carried features of r_{T S}
        cmb.b_lo = 1; cmb.b_hi = 0; cmb.has_carry = 1; cmb.tail = 1; cmb.k0 = T * S;
        REACH_TRY(combine(cmb, st));
        REACH_TRY(disjoint_scan(T * int64_t(S), p->nsteps, st));
    }

    if (p->nchecks > 0 && T == 0) REACH_TRY(disjoint_scan(0, p->nsteps, st));
    REACH_CUDA(ctx, cudaEventRecord(p->ev1, st));
    p->ran = true;
    return REACH_OK;
}

// ---- the row stack in steps, driven from outside -----------------------------------------------------------------
// Step 0 is block 1 = block_0 x P_1; step j >= 1 is the doubling step s = 2^(j-1): blocks s+1 .. s+min(s, S-s) = (blocks
// 1..) x P_s, after Q_{2s} = Q_s Q_s (an n x n x n product every caller repeats).  A caller computes rows [row_lo, row_hi)
// of the step's `*step_rows` output rows, which start at stack row `*dest_row`; the ranks of a sharded run take disjoint row
// ranges and all-gather them in place (reach_lgg09_plan_stack_info gives the device buffer), so each builds 1/N of the
// stack.  The products and their order are those of reach_lgg09_plan_run: the stack is bit-identical.
extern "C" int reach_lgg09_plan_stack_info(reach_lgg09_plan* p, void** stack, int64_t* rows_per_block, int64_t* ld, int32_t* S) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->uploaded) return set_error(ctx, REACH_ESTATE, "plan_stack_info before plan_upload");
    if (p->use_compact || p->use_chain)
        return set_error(ctx, REACH_EUNSUPPORTED, "this plan runs a chain kernel: it has no row stack");
    if (stack) *stack = p->stack.p;
    if (rows_per_block) *rows_per_block = p->nbe;
    if (ld) *ld = p->Kp;
    if (S) *S = p->S;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_stack_step(reach_lgg09_plan* p, int32_t step, int64_t row_lo, int64_t row_hi, int64_t* step_rows,
                                           int64_t* dest_row) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->uploaded) return set_error(ctx, REACH_ESTATE, "plan_stack_step before plan_upload");
    if (p->use_compact || p->use_chain)
        return set_error(ctx, REACH_EUNSUPPORTED, "this plan runs a chain kernel: it has no row stack");
    if (step < 0 || step > 30) return set_error(ctx, REACH_EINVAL, "stack step %d out of range", step);
    const int n = p->n, Kp = p->Kp, nbe = p->nbe, S = p->S;
    int64_t rows = 0, dest = 0;
    int s = 0;
    if (step == 0) {
        rows = nbe;
        dest = nbe;
    } else {
        s = 1 << (step - 1);
        if (s < S) {
            rows = int64_t(std::min(s, S - s)) * nbe;
            dest = int64_t(s + 1) * nbe;
        }
    }
    if (step_rows) *step_rows = rows;
    if (dest_row) *dest_row = dest;
    if (row_lo < 0 || rows == 0) return REACH_OK;                      // query only / past the last step
    if (row_lo > row_hi || row_hi > rows)
        return set_error(ctx, REACH_EINVAL, "rows [%lld, %lld) outside the step's %lld rows", (long long)row_lo, (long long)row_hi,
                         (long long)rows);
    // in order; a step may be called several times with different row ranges (its n x n x n product is done once)
    const bool first_call = step == p->stack_steps_done;
    if (!first_call && step != p->stack_steps_done - 1)
        return set_error(ctx, REACH_ESTATE, "stack steps must be taken in order (expected %d, got %d)", p->stack_steps_done, step);
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double* stack = p->stack.as<double>();
    const bool big = (p->packed || nbe % 128 == 0) && n >= 512;
    auto plain_gemm = [&](const double* X, int mrows, const double* Y, int yrows, double* out, int m_valid) -> int {
        const int t = big ? 128 : 64;
        GemmOperands g{X, Y, Kp, Kp, Kp, round_up(mrows, t) / t, round_up(yrows, t) / t};
        StoreEpi epi{out, Kp, m_valid, n};
        REACH_CUDA(ctx, (launch_cfg(big ? 0 : 1, g, epi, st)));
        return REACH_OK;
    };
    const int nrows = int(row_hi - row_lo);
    if (step == 0) {
        if (first_call) {
            dim3 grid((n + 31) / 32, (n + 31) / 32), blk(32, 8);
            transpose_kernel<<<grid, blk, 0, st>>>(p->phi_cm.as<double>(), n, p->Qbuf[0].as<double>(), Kp, n);
            REACH_CUDA(ctx, cudaGetLastError());
        }
        if (nrows > 0)
            REACH_TRY(plain_gemm(stack + size_t(row_lo) * Kp, nrows, p->Qbuf[0].as<double>(), n, stack + size_t(dest + row_lo) * Kp, nrows));
    } else {
        const int qcur = (step - 1) & 1;                 // Q_s ping-pongs between the two buffers from step 1 on
        if (first_call && 2 * s < S)      // needs the whole block s: the caller has gathered the previous step
            REACH_TRY(plain_gemm(p->Qbuf[qcur].as<double>(), n, stack + size_t(s) * nbe * Kp, n, p->Qbuf[1 - qcur].as<double>(), n));
        if (nrows > 0)
            REACH_TRY(plain_gemm(stack + size_t(nbe + row_lo) * Kp, nrows, p->Qbuf[qcur].as<double>(), n,
                                 stack + size_t(dest + row_lo) * Kp, nrows));
    }
    p->stack_steps_done = step + 1;
    return REACH_OK;
}

// every step has been taken (and gathered): the next reach_lgg09_plan_run uses the stack as it is
extern "C" int reach_lgg09_plan_stack_ready(reach_lgg09_plan* p) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    int need = 1;
    for (int s = 1; s < p->S; s *= 2) ++need;
    if (p->stack_steps_done != need)
        return set_error(ctx, REACH_ESTATE, "%d of %d stack steps taken", p->stack_steps_done, need);
    p->stack_ready = true;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_sync(reach_lgg09_plan* p) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->ran) return set_error(ctx, REACH_ESTATE, "plan_sync before plan_run");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    REACH_CUDA(ctx, cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    p->last_ms = ms;
    double acc = 0.0;
    for (int i = 0; i < p->gemm_sampled; ++i) {
        float t = 0.f;
        REACH_CUDA(ctx, cudaEventElapsedTime(&t, p->gemm_ev[2 * i], p->gemm_ev[2 * i + 1]));
        acc += t;
    }
    p->gemm_ms = p->gemm_sampled ? acc / p->gemm_sampled : 0.0;   // mean per launch
    p->nsteps_done = p->steps_computed;
    if (p->nchecks > 0) {
        unsigned long long fb = 0;
        REACH_CUDA(ctx, cudaMemcpy(&fb, p->first_bad.p, sizeof(fb), cudaMemcpyDeviceToHost));
        if (fb != ~0ull && int64_t(fb) < p->nsteps_done) p->nsteps_done = int64_t(fb);
    }
    p->synced = true;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_fetch(reach_lgg09_plan* p, int64_t step0, int64_t nsteps, double* rho_host) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->ran) return set_error(ctx, REACH_ESTATE, "plan_fetch before plan_run");
    if (!rho_host) return set_error(ctx, REACH_EINVAL, "rho_host is NULL");
    if (step0 < 0 || nsteps < 0 || step0 + nsteps > p->nsteps)
        return set_error(ctx, REACH_EINVAL, "fetch range [%lld, %lld) outside [0, %lld)", (long long)step0,
                         (long long)(step0 + nsteps), (long long)p->nsteps);
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    REACH_CUDA(ctx, cudaMemcpyAsync(rho_host, p->rho + size_t(step0) * p->ndirs,
                                    size_t(nsteps) * p->ndirs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_nsteps_done(reach_lgg09_plan* p, int64_t* out) {
    if (!p || !out) return REACH_EINVAL;
    if (!p->ran) return set_error(p->ctx, REACH_ESTATE, "nsteps_done before plan_run");
    if (!p->synced) REACH_TRY(reach_lgg09_plan_sync(p));
    *out = p->nsteps_done;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_max_over_steps(reach_lgg09_plan* p, double* max_host, int64_t* argmax_host) {
    if (!p) return REACH_EINVAL;
    reach_ctx* ctx = p->ctx;
    if (!p->ran) return set_error(ctx, REACH_ESTATE, "max_over_steps before plan_run");
    if (!max_host) return set_error(ctx, REACH_EINVAL, "max_host is NULL");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!p->synced) REACH_TRY(reach_lgg09_plan_sync(p));          // nsteps_done is only known after the run has finished
    DevBuf vmax, amax;
    REACH_TRY(vmax.alloc(ctx, sizeof(double) * p->ndirs, false));
    REACH_TRY(amax.alloc(ctx, sizeof(long long) * p->ndirs, false));
    const int th = 128;
    max_over_steps_kernel<<<(p->ndirs + th - 1) / th, th, 0, ctx->stream>>>(
        p->rho, p->ndirs, p->nsteps_done, vmax.as<double>(), amax.as<long long>());   // empty flowpipe: -inf, argmax -1
    REACH_CUDA(ctx, cudaGetLastError());
    REACH_CUDA(ctx, cudaMemcpyAsync(max_host, vmax.p, sizeof(double) * p->ndirs, cudaMemcpyDeviceToHost, ctx->stream));
    if (argmax_host) {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64");
        REACH_CUDA(ctx, cudaMemcpyAsync(argmax_host, amax.p, sizeof(int64_t) * p->ndirs, cudaMemcpyDeviceToHost, ctx->stream));
    }
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REACH_OK;
}

extern "C" void* reach_lgg09_plan_device_rho(reach_lgg09_plan* p) { return p ? p->rho : nullptr; }

extern "C" int reach_lgg09_plan_stats(reach_lgg09_plan* p, double* ms, int64_t* launches, int32_t* time_block,
                                      int32_t* tile_m, int32_t* tile_n, int32_t* nchains, double* gemm_ms,
                                      int64_t* gemm_launches) {
    if (!p) return REACH_EINVAL;
    if (ms) *ms = p->last_ms;
    if (launches) *launches = p->launches;
    if (time_block) *time_block = p->S;
    if (tile_m) *tile_m = (p->use_chain || p->use_compact) ? 0 : p->TM;
    if (tile_n) *tile_n = (p->use_chain || p->use_compact) ? 0 : p->TN;
    if (nchains) *nchains = p->mu;
    if (gemm_ms) *gemm_ms = p->gemm_ms;
    if (gemm_launches) *gemm_launches = p->gemm_total;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_path(reach_lgg09_plan* p, int32_t* path, int32_t* digit_planes, double* guard_max,
                                     int64_t* guard_checks, int64_t* fallback_pass) {
    if (!p) return REACH_EINVAL;
    if (path) *path = p->use_compact ? 4 : (p->use_chain ? 2 : (p->use_oz ? 3 : 1));
    if (digit_planes) *digit_planes = p->use_oz ? p->oz_T : 0;
    if (guard_max) *guard_max = p->oz_guard_max;
    if (guard_checks) *guard_checks = p->oz_checks;
    if (fallback_pass) *fallback_pass = p->oz_fallback_pass;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_invariant_info(reach_lgg09_plan* p, int32_t* faces, int32_t* covered, int32_t* exact,
                                               int64_t* steps_computed) {
    if (!p) return REACH_EINVAL;
    if (faces) *faces = p->inv_faces;
    if (covered) *covered = p->inv_covered;
    if (exact) *exact = p->inv_exact;
    if (steps_computed) *steps_computed = p->ran ? p->steps_computed : 0;
    return REACH_OK;
}

extern "C" int reach_lgg09_plan_guard(reach_lgg09_plan* p, double* drift_max, int64_t* replay_from, int64_t* replayed_passes) {
    if (!p) return REACH_EINVAL;
    if (drift_max) *drift_max = p->oz_drift_max;
    if (replay_from) *replay_from = p->oz_replay_from;
    if (replayed_passes) *replayed_passes = p->oz_replayed;
    return REACH_OK;
}

extern "C" int reach_lgg09(reach_ctx* ctx, int n, int ndirs, int64_t nsteps, const double* Phi, const double* dirs,
                           const reach_setexpr* Omega0, const reach_setexpr* V, const reach_invariant* inv,
                           const reach_lgg09_opts* opts, double* rho_host, int64_t* nsteps_done) {
    const bool timing = getenv("REACH_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t0 = now();
    reach_lgg09_plan* p = nullptr;
    REACH_TRY(reach_lgg09_plan_create(ctx, n, ndirs, nsteps, opts, &p));
    int rc = reach_lgg09_plan_upload(p, Phi, dirs, Omega0, V, inv);
    auto t1 = now();
    if (rc == REACH_OK) rc = reach_lgg09_plan_run(p);
    if (rc == REACH_OK) rc = reach_lgg09_plan_sync(p);
    auto t2 = now();
    if (rc == REACH_OK && rho_host && p->nsteps_done > 0) rc = reach_lgg09_plan_fetch(p, 0, p->nsteps_done, rho_host);
    if (rc == REACH_OK && nsteps_done) *nsteps_done = p->nsteps_done;
    auto t3 = now();
    reach_lgg09_plan_destroy(p);
    auto t4 = now();
    if (timing) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[reach_lgg09] create+upload %.1f ms, run+sync %.1f ms, fetch %.1f ms, destroy %.1f ms\n",
                ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
    }
    return rc;
}

// BOX algorithm (Algorithms/BOX/reach_homog.jl:8-47, reach_inhomog.jl:8-60): c_k = Phi^(k-1) c_1 + W_k.c,
// r_k = |Phi^(k-1)| r_1 + W_k.r with W_{k+1} = W_k + (Phi^(k-1) U.c, |Phi^(k-1)| U.r).  Row i of Phi^k is the chain
// (Phi')^k e_i, so this is the LGG09 recurrence over the box template with the (lin, abs) pair written out instead
// of abs +- lin.  recursive != 0 (reach_homog.jl:50-69): H_k = box(Phi H_{k-1}), i.e. c_k = Phi c_{k-1} and
// r_k = |Phi| r_{k-1} (homogeneous only, as in the reference): the same recurrence with singleton sets, once with
// Phi and once with |Phi|.
extern "C" int reach_box(reach_ctx* ctx, int n, int64_t nsteps, const double* Phi, const double* c0, const double* r0,
                         const double* cU, const double* rU, int recursive, const reach_lgg09_opts* opts,
                         double* c_host, double* r_host) {
    if (!ctx) return REACH_EINVAL;
    if (n < 1 || nsteps < 1) return set_error(ctx, REACH_EINVAL, "reach_box: n and NSTEPS must be >= 1");
    if (!Phi || !c0 || !r0) return set_error(ctx, REACH_EINVAL, "reach_box: Phi, c0 and r0 must not be NULL");
    if ((cU == nullptr) != (rU == nullptr)) return set_error(ctx, REACH_EINVAL, "reach_box: cU and rU must both be given or both be NULL");
    for (int i = 0; i < n; ++i)
        if (r0[i] < 0 || (rU && rU[i] < 0)) return set_error(ctx, REACH_EINVAL, "reach_box: negative radius");
    if (recursive && cU)     // the reference has no such method either (reach_inhomog.jl only defines recursive::Val{false})
        return set_error(ctx, REACH_EUNSUPPORTED, "reach_box: the recursive variant is defined for homogeneous systems only");
    reach_lgg09_opts o;
    if (opts) o = *opts; else reach_lgg09_default_opts(&o);
    o.reserved[1] = 1;                                   // (lin, abs) outputs
    o.share_negated = 1;
    std::vector<double> dirs(size_t(n) * 2 * n, 0.0);    // [I, -I]: chain i = e_i, outputs i (centre) and n + i (radius)
    for (int i = 0; i < n; ++i) { dirs[size_t(i) * n + i] = 1.0; dirs[size_t(n + i) * n + i] = -1.0; }
    std::vector<double> out(size_t(2) * n * nsteps);
    auto run = [&](const double* M, const double* oc, const double* orad, const double* vc, const double* vr) -> int {
        reach_term t0{REACH_TERM_BOX, 0, 0, 0, nullptr, oc, orad, nullptr};
        reach_term tv{REACH_TERM_BOX, 0, 0, 0, nullptr, vc, vr, nullptr};
        const int32_t one = 1;
        reach_setexpr O{1, &one, &t0}, V{1, &one, &tv};
        return reach_lgg09(ctx, n, 2 * n, nsteps, M, dirs.data(), &O, (vc || vr) ? &V : nullptr, nullptr, &o, out.data(), nullptr);
    };
    auto scatter = [&](double* dst, int row0) {
        if (!dst) return;
        for (int64_t k = 0; k < nsteps; ++k)
            std::memcpy(dst + size_t(k) * n, out.data() + size_t(k) * 2 * n + row0, sizeof(double) * n);
    };
    if (!recursive) {
        REACH_TRY(run(Phi, c0, r0, cU, rU));
        scatter(c_host, 0);
        scatter(r_host, n);
        return REACH_OK;
    }
    REACH_TRY(run(Phi, c0, nullptr, cU, nullptr));       // centres: singleton sets, Phi
    scatter(c_host, 0);
    std::vector<double> absPhi(size_t(n) * n);
    for (size_t i = 0; i < absPhi.size(); ++i) absPhi[i] = std::fabs(Phi[i]);
    REACH_TRY(run(absPhi.data(), r0, nullptr, rU, nullptr));   // radii: |Phi|^k r_0 + sum |Phi|^i r_U (all terms >= 0)
    scatter(r_host, 0);
    return REACH_OK;
}

// FP64 throughput probe of the DMMA kernel on an m x n x k problem (128x128 tiles).
extern "C" int reach_dgemm_probe(reach_ctx* ctx, int m, int n, int k, int iters, double* tflops) {
    if (!ctx || !tflops) return REACH_EINVAL;
    if (m < 1 || n < 1 || k < 1 || iters < 1) return set_error(ctx, REACH_EINVAL, "probe sizes must be positive");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int mp = round_up(m, 128), np = round_up(n, 128), kp = round_up(k, BK);
    DevBuf X, Y, C;
    REACH_TRY(X.alloc(ctx, size_t(mp) * kp * 8));
    REACH_TRY(Y.alloc(ctx, size_t(np) * kp * 8));
    REACH_TRY(C.alloc(ctx, size_t(mp) * np * 8));
    std::vector<double> h(size_t(std::max(mp, np)) * kp);
    for (size_t i = 0; i < h.size(); ++i) h[i] = double((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    REACH_CUDA(ctx, cudaMemcpy(X.p, h.data(), size_t(mp) * kp * 8, cudaMemcpyHostToDevice));
    REACH_CUDA(ctx, cudaMemcpy(Y.p, h.data(), size_t(np) * kp * 8, cudaMemcpyHostToDevice));
    GemmOperands g{X.as<double>(), Y.as<double>(), kp, kp, kp, mp / 128, np / 128};
    StoreEpi epi{C.as<double>(), np, mp, np};
    cudaEvent_t e0, e1;
    REACH_CUDA(ctx, cudaEventCreate(&e0));
    REACH_CUDA(ctx, cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) REACH_CUDA(ctx, (launch_cfg(0, g, epi, ctx->stream)));
    REACH_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < iters; ++i) REACH_CUDA(ctx, (launch_cfg(0, g, epi, ctx->stream)));
    REACH_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    REACH_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = 2.0 * m * double(n) * k * iters / (ms * 1e-3) / 1e12;
    return REACH_OK;
}

#include "lgg09_batch.inl"
