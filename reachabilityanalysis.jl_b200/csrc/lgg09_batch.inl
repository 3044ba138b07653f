// lgg09_batch.inl -- LGG09 over a vector of initial sets (included at the end of lgg09.cu).
//
// The reference's split caller `_solve_distributed` (src/Continuous/solve.jl:84-117) re-discretises and re-runs
// `post` once per initial set.  Phi, the template and the input set V are the same for every split; only Omega0
// differs.  The direction chains r_k = (Phi')^k l and the running input sum s_k therefore do not depend on the
// split at all, and the per-split part of reach_inhomog_dir_LGG09! (reach_inhomog.jl:49-64),
//     rho_i[j,k] = max_a ( F_{i,a,0}(r_k) + F_{i,a,1}(r_{k+1}) ) + s_k ,     F(r) = a.r + b.|r| + wl.(E r) + wa.|E r| ,
// is one big contraction: feature rows [a | b | a' | b'] (one per split and alternative) against the columns
// [y_k ; |y_k| ; y_{k+1} ; |y_{k+1}|], y = [I;E] r, of every (step, direction).  Two FP64 tensor-core GEMMs do it:
//   (A) the chains: L_k x ([I;E](Phi')^b, b = 0..S)'  with RREpi, which writes y and |y| of every step straight into
//       the column operand RR of (B) (both halves), and the block-S tile as L_{k+S};
//   (B) X (splits*alternatives x Kc) x RR' with SplitEpi: max over the alternatives by quad-row shuffles, + s_k, one
//       coalesced store into rho[split][step][dir] -- each split's ndirs x NSTEPS matrix exactly as the reference's.
// s_k comes from an ordinary device plan with Omega0 = ZeroSet (its rho IS s_k: reach_inhomog.jl:66-81).
// RR is produced in time chunks so that it stays within a fixed device budget.

namespace {

struct RREpi {
    double* RR;          // [chunk steps * ndirs (+pad)][Kc]
    double* Lnext;       // [ndirs_pad][ldl]
    int Kc, nw, n, halves, S, nbe, ndirs, ldl;
    long long k0, c0, ch;   // first step of this pass; chunk [c0, c0 + ch)

    template <int MT, int NT, int WN, class TC>
    __device__ __forceinline__ void run(double (&acc)[MT][NT][2], const TC& tc, double*) const {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = tc.tn * TC::TN + tc.wn0 + j * 8 + 2 * tc.t4;     // even; nbe is a multiple of 8
            const int b = col / nbe, i = col - b * nbe;
            if (b > S || i >= nw) continue;
            const long long step = k0 + b;
            const bool first = b < S && step >= c0 && step < c0 + ch;
            const bool second = halves == 2 && b >= 1 && step - 1 >= c0 && step - 1 < c0 + ch;
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) {
                const int m = tc.tm * TC::TM + tc.wm0 + mi * 8 + tc.g4;
                if (m >= ndirs) continue;
                const double v0 = acc[mi][j][0], v1 = acc[mi][j][1];
                const bool two = i + 1 < nw;
                if (b == S && i < n) {
                    Lnext[size_t(m) * ldl + i] = v0;
                    if (i + 1 < n) Lnext[size_t(m) * ldl + i + 1] = v1;
                }
                if (first) {
                    double* row = RR + (size_t(step - c0) * ndirs + m) * Kc;
                    row[i] = v0;
                    row[nw + i] = fabs(v0);
                    if (two) { row[i + 1] = v1; row[nw + i + 1] = fabs(v1); }
                }
                if (second) {
                    double* row = RR + (size_t(step - 1 - c0) * ndirs + m) * Kc + 2 * nw;
                    row[i] = v0;
                    row[nw + i] = fabs(v0);
                    if (two) { row[i + 1] = v1; row[nw + i + 1] = fabs(v1); }
                }
            }
        }
    }
};

struct SplitEpi {
    double* out;            // [nsplits][nsteps*ndirs]
    const double* s;        // [nsteps*ndirs] running input sums (NULL: homogeneous)
    long long split_stride; // nsteps * ndirs
    long long col0, ncols;  // columns of this chunk: global index col0 + c, c < ncols
    int nsplits, na;        // alternatives per split (1, 2, 4 or 8 rows)

    template <int MT, int NT, int WN, class TC>
    __device__ __forceinline__ void run(double (&acc)[MT][NT][2], const TC& tc, double*) const {
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            const int m = tc.tm * TC::TM + tc.wm0 + mi * 8 + tc.g4;           // row = split * na + alternative
            const int split = m / na;
            const bool writer = (tc.g4 & (na - 1)) == 0 && split < nsplits;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                double v0 = acc[mi][j][0], v1 = acc[mi][j][1];
                if (na >= 2) { v0 = fmax(v0, __shfl_xor_sync(0xffffffffu, v0, 4)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, 4)); }
                if (na >= 4) { v0 = fmax(v0, __shfl_xor_sync(0xffffffffu, v0, 8)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, 8)); }
                if (na >= 8) { v0 = fmax(v0, __shfl_xor_sync(0xffffffffu, v0, 16)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, 16)); }
                const long long c = (long long)tc.tn * TC::TN + tc.wn0 + j * 8 + 2 * tc.t4;
                if (!writer || c >= ncols) continue;
                const long long gi = col0 + c;
                double* o = out + size_t(split) * split_stride + gi;
                if (s) { v0 += s[gi]; if (c + 1 < ncols) v1 += s[gi + 1]; }
                if (c + 1 < ncols && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                    *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
                } else {
                    o[0] = v0;
                    if (c + 1 < ncols) o[1] = v1;
                }
            }
        }
    }
};

template <class Epi>
cudaError_t launch_batch_gemm(int cfg, const GemmOperands& g, const Epi& epi, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_nt_dgemm_ws<2, 4, 8, 4, 5>(g, epi, st);   // 128 x 128
        case 1: return launch_nt_dgemm<2, 2, 4, 4, 4>(g, epi, st);      // 64 x 64
        case 2: return launch_nt_dgemm<1, 4, 2, 4, 4>(g, epi, st);      // 16 x 128
    }
    return cudaErrorInvalidValue;
}
constexpr int kBatchTM[3] = {128, 64, 16}, kBatchTN[3] = {128, 64, 128};

}  // namespace

struct reach_lgg09_batch {
    reach_ctx* ctx = nullptr;
    int n = 0, ndirs = 0, nsplits = 0;
    int64_t nsteps = 0;
    int Kp = 0, ne = 0, nw = 0, nbe = 0, S = 1, halves = 1, na = 1, Kc = 0;
    int cfgA = 1, cfgB = 0, cfgP = 1;
    int ndirs_pad = 0, xrows_pad = 0;
    int64_t chunk = 0;
    DevBuf phi_rm, stack, L0, L[2], X, RR, rho;
    reach_lgg09_plan* splan = nullptr;      // Omega0 = ZeroSet, same V: its rho is s_k
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    bool ran = false;
    ~reach_lgg09_batch() {
        if (splan) reach_lgg09_plan_destroy(splan);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

extern "C" int reach_lgg09_batch_create(reach_ctx* ctx, int n, int ndirs, int64_t nsteps, int nsplits, const double* Phi,
                                        const double* dirs, const reach_setexpr* Omega0s, const reach_setexpr* V,
                                        reach_lgg09_batch** out) {
    if (!ctx) return REACH_EINVAL;
    if (!out) return set_error(ctx, REACH_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 1 || ndirs < 1 || nsteps < 1 || nsplits < 1)
        return set_error(ctx, REACH_EINVAL, "n, ndirs, nsteps and nsplits must be >= 1 (got %d, %d, %lld, %d)", n, ndirs,
                         (long long)nsteps, nsplits);
    if (!Phi || !dirs || !Omega0s) return set_error(ctx, REACH_EINVAL, "Phi, dirs and Omega0s must not be NULL");
    for (size_t i = 0; i < size_t(n) * n; ++i)
        if (!std::isfinite(Phi[i])) return set_error(ctx, REACH_EINVAL, "Phi contains a non-finite entry");
    for (size_t i = 0; i < size_t(n) * ndirs; ++i)
        if (!std::isfinite(dirs[i])) return set_error(ctx, REACH_EINVAL, "direction %zu is not finite", i / n);
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    // ---- features of every (split, alternative): F0 at r_k, F1 at r_{k+1}; the mapped rows E are shared ----
    Program P;
    P.n = n;
    struct AltFeat { Feature f[2]; };
    std::vector<std::vector<AltFeat>> alts(nsplits);
    int na = 1;
    bool need_next = false;
    for (int sidx = 0; sidx < nsplits; ++sidx) {
        const reach_setexpr& O = Omega0s[sidx];
        if (O.nalt < 1) return set_error(ctx, REACH_EINVAL, "Omega0 of split %d must have at least one alternative", sidx);
        if (O.nalt > 8) return set_error(ctx, REACH_EUNSUPPORTED, "Omega0 of split %d has %d alternatives (max 8)", sidx, O.nalt);
        na = std::max(na, int(O.nalt));
        const reach_term* t = O.terms;
        alts[sidx].resize(O.nalt);
        for (int a = 0; a < O.nalt; ++a) {
            AltFeat& af = alts[sidx][a];
            for (auto& x : af.f) { x.a.assign(n, 0.0); x.b.assign(n, 0.0); }
            const int len = O.alt_len ? O.alt_len[a] : 0;
            if (len < 0) return set_error(ctx, REACH_EINVAL, "negative alt_len");
            for (int i = 0; i < len; ++i, ++t) {
                if (t->which != 0 && t->which != 1) return set_error(ctx, REACH_EINVAL, "term: which must be 0 or 1");
                REACH_TRY(add_term(ctx, P, af.f[t->which], *t));
            }
            if (!af.f[1].empty()) need_next = true;
        }
    }
    while (na & (na - 1)) ++na;                       // 1, 2, 4, 8 rows per split

    reach_lgg09_batch* b = new (std::nothrow) reach_lgg09_batch;
    if (!b) return set_error(ctx, REACH_ENOMEM, "out of host memory");
    std::unique_ptr<reach_lgg09_batch> guard(b);
    b->ctx = ctx;
    b->n = n;
    b->ndirs = ndirs;
    b->nsteps = nsteps;
    b->nsplits = nsplits;
    b->Kp = round_up(n, BK);
    b->ne = int(P.erows.size());
    b->nw = n + b->ne;
    b->nbe = round_up(b->nw, 8);
    b->halves = need_next ? 2 : 1;
    b->na = na;
    b->Kc = round_up(b->halves * 2 * b->nw, BK);
    const int nbe = b->nbe, Kp = b->Kp, Kc = b->Kc, nw = b->nw;
    b->S = std::max(1, std::min(64, 2048 / nbe));
    b->S = int(std::min<int64_t>(b->S, nsteps));
    const int S = b->S;
    const bool bigA = n >= 512 && ndirs >= 128;
    b->cfgA = bigA ? 0 : 1;
    b->cfgP = 1;
    const int xrows = nsplits * na;
    b->cfgB = xrows >= 96 ? 0 : (xrows > 16 ? 1 : 2);
    b->ndirs_pad = round_up(ndirs, kBatchTM[b->cfgA]);
    b->xrows_pad = round_up(xrows, kBatchTM[b->cfgB]);
    // time chunk of the column operand: <= ~2 GB, a multiple of S
    {
        const double per_step = double(ndirs) * Kc * 8.0;
        double budget = 2.0e9;
        if (const char* env = getenv("REACH_BATCH_CHUNK_BYTES")) budget = std::max(1.0, atof(env));   // tests: force many chunks
        int64_t c = int64_t(budget / per_step);
        c = std::max<int64_t>(S, c / S * S);
        b->chunk = std::min<int64_t>(c, round_up64(nsteps, S));
    }
    REACH_CUDA(ctx, cudaEventCreate(&b->ev0));
    REACH_CUDA(ctx, cudaEventCreate(&b->ev1));

    // ---- device buffers ----
    const size_t stack_rows = size_t(S + 1) * nbe + 256;
    REACH_TRY(b->phi_rm.alloc(ctx, size_t(round_up(n, 64)) * Kp * 8));
    REACH_TRY(b->stack.alloc(ctx, stack_rows * Kp * 8));
    REACH_TRY(b->L0.alloc(ctx, size_t(b->ndirs_pad) * Kp * 8));
    REACH_TRY(b->L[0].alloc(ctx, size_t(b->ndirs_pad) * Kp * 8));
    REACH_TRY(b->L[1].alloc(ctx, size_t(b->ndirs_pad) * Kp * 8));
    REACH_TRY(b->X.alloc(ctx, size_t(b->xrows_pad) * Kc * 8));
    REACH_TRY(b->RR.alloc(ctx, (size_t(b->chunk) * ndirs + 256) * Kc * 8));
    REACH_TRY(b->rho.alloc(ctx, size_t(nsplits) * nsteps * ndirs * 8, false));

    // ---- host staging ----
    {
        std::vector<double> h(size_t(round_up(n, 64)) * Kp, 0.0);
        for (int l = 0; l < n; ++l)
            for (int m = 0; m < n; ++m) h[size_t(l) * Kp + m] = Phi[size_t(l) + size_t(n) * m];     // row-major Phi
        REACH_CUDA(ctx, cudaMemcpyAsync(b->phi_rm.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
        std::vector<double> blk0(size_t(nbe) * Kp, 0.0);                                            // [I ; E]
        for (int i = 0; i < n; ++i) blk0[size_t(i) * Kp + i] = 1.0;
        for (int e = 0; e < b->ne; ++e) std::memcpy(&blk0[size_t(n + e) * Kp], P.erows[e].data(), sizeof(double) * n);
        REACH_CUDA(ctx, cudaMemcpyAsync(b->stack.p, blk0.data(), blk0.size() * 8, cudaMemcpyHostToDevice, st));
        std::vector<double> l0(size_t(b->ndirs_pad) * Kp, 0.0);
        for (int j = 0; j < ndirs; ++j) std::memcpy(&l0[size_t(j) * Kp], dirs + size_t(j) * n, sizeof(double) * n);
        REACH_CUDA(ctx, cudaMemcpyAsync(b->L0.p, l0.data(), l0.size() * 8, cudaMemcpyHostToDevice, st));
        std::vector<double> x(size_t(b->xrows_pad) * Kc, 0.0);
        for (int sidx = 0; sidx < nsplits; ++sidx)
            for (int a = 0; a < na; ++a) {
                // a split with fewer alternatives repeats its first one (the max is unchanged)
                const AltFeat& af = alts[sidx][a < int(alts[sidx].size()) ? a : 0];
                double* row = &x[(size_t(sidx) * na + a) * Kc];
                for (int w = 0; w < b->halves; ++w) {
                    const Feature& f = af.f[w];
                    double* lin = row + size_t(w) * 2 * nw;
                    double* ab = lin + nw;
                    for (int i = 0; i < n; ++i) { lin[i] = f.a[i]; ab[i] = f.b[i]; }
                    for (auto& kv : f.rows) { lin[n + kv.first] = kv.second.first; ab[n + kv.first] = kv.second.second; }
                }
            }
        REACH_CUDA(ctx, cudaMemcpyAsync(b->X.p, x.data(), x.size() * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(ctx, cudaStreamSynchronize(st));
    }

    // ---- s_k: an ordinary plan with Omega0 = ZeroSet (reach_inhomog.jl:66-81) ----
    if (V) {
        reach_lgg09_opts o;
        reach_lgg09_default_opts(&o);
        REACH_TRY(reach_lgg09_plan_create(ctx, n, ndirs, nsteps, &o, &b->splan));
        const int32_t zero_len = 0;
        reach_setexpr zero{1, &zero_len, nullptr};
        REACH_TRY(reach_lgg09_plan_upload(b->splan, Phi, dirs, &zero, V, nullptr));
    }
    *out = guard.release();
    return REACH_OK;
}

extern "C" int reach_lgg09_batch_run(reach_lgg09_batch* b) {
    if (!b) return REACH_EINVAL;
    reach_ctx* ctx = b->ctx;
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = b->n, ndirs = b->ndirs, Kp = b->Kp, nbe = b->nbe, S = b->S, Kc = b->Kc;
    b->launches = 0;
    REACH_CUDA(ctx, cudaEventRecord(b->ev0, st));
    const double* s_dev = nullptr;
    if (b->splan) {
        REACH_TRY(reach_lgg09_plan_run(b->splan));
        b->launches += b->splan->launches;
        s_dev = b->splan->rho;
    }
    double* stack = b->stack.as<double>();
    // row stack [I;E](Phi')^b, b = 1..S, by sequential right-multiplication (the reference's association)
    for (int blk = 0; blk < S; ++blk) {
        const int t = kBatchTM[b->cfgP];
        GemmOperands g{stack + size_t(blk) * nbe * Kp, b->phi_rm.as<double>(), Kp, Kp, Kp, round_up(nbe, t) / t,
                       round_up(n, kBatchTN[b->cfgP]) / kBatchTN[b->cfgP]};
        StoreEpi epi{stack + size_t(blk + 1) * nbe * Kp, Kp, nbe, n};
        REACH_CUDA(ctx, (launch_batch_gemm(b->cfgP, g, epi, st)));
        ++b->launches;
    }
    REACH_CUDA(ctx, cudaMemcpyAsync(b->L[0].p, b->L0.p, b->L0.bytes, cudaMemcpyDeviceToDevice, st));
    int cur = 0;
    const int tnA = kBatchTN[b->cfgA], tmA = kBatchTM[b->cfgA], tmB = kBatchTM[b->cfgB], tnB = kBatchTN[b->cfgB];
    for (int64_t c0 = 0; c0 < b->nsteps; c0 += b->chunk) {
        const int64_t ch = std::min<int64_t>(b->chunk, b->nsteps - c0);
        for (int64_t k0 = c0; k0 < c0 + ch; k0 += S) {
            GemmOperands g{b->L[cur].as<double>(), stack, Kp, Kp, Kp, b->ndirs_pad / tmA,
                           int((size_t(S + 1) * nbe + tnA - 1) / tnA)};
            RREpi epi{b->RR.as<double>(), b->L[1 - cur].as<double>(), Kc, b->nw, n, b->halves, S, nbe, ndirs, Kp, k0, c0, ch};
            REACH_CUDA(ctx, (launch_batch_gemm(b->cfgA, g, epi, st)));
            ++b->launches;
            cur ^= 1;
        }
        const int64_t ncols = ch * ndirs;
        GemmOperands g{b->X.as<double>(), b->RR.as<double>(), Kc, Kc, Kc, b->xrows_pad / tmB, int((ncols + tnB - 1) / tnB)};
        SplitEpi epi{b->rho.as<double>(), s_dev, (long long)(b->nsteps) * ndirs, c0 * ndirs, ncols, b->nsplits, b->na};
        REACH_CUDA(ctx, (launch_batch_gemm(b->cfgB, g, epi, st)));
        ++b->launches;
    }
    REACH_CUDA(ctx, cudaEventRecord(b->ev1, st));
    b->ran = true;
    return REACH_OK;
}

extern "C" int reach_lgg09_batch_sync(reach_lgg09_batch* b) {
    if (!b) return REACH_EINVAL;
    REACH_CUDA(b->ctx, cudaSetDevice(b->ctx->device));
    REACH_CUDA(b->ctx, cudaStreamSynchronize(b->ctx->stream));
    return REACH_OK;
}

extern "C" int reach_lgg09_batch_fetch(reach_lgg09_batch* b, int split, int64_t step0, int64_t nsteps, double* rho_host) {
    if (!b) return REACH_EINVAL;
    reach_ctx* ctx = b->ctx;
    if (!b->ran) return set_error(ctx, REACH_ESTATE, "batch_fetch before batch_run");
    if (split < 0 || split >= b->nsplits || step0 < 0 || nsteps < 0 || step0 + nsteps > b->nsteps || !rho_host)
        return set_error(ctx, REACH_EINVAL, "batch_fetch: bad split / step range");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    REACH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double* src = b->rho.as<double>() + (size_t(split) * b->nsteps + step0) * b->ndirs;
    REACH_CUDA(ctx, cudaMemcpy(rho_host, src, size_t(nsteps) * b->ndirs * 8, cudaMemcpyDeviceToHost));
    return REACH_OK;
}

extern "C" int reach_lgg09_batch_max_over_steps(reach_lgg09_batch* b, int split, double* max_host) {
    if (!b) return REACH_EINVAL;
    reach_ctx* ctx = b->ctx;
    if (!b->ran) return set_error(ctx, REACH_ESTATE, "batch_max_over_steps before batch_run");
    if (split < -1 || split >= b->nsplits || !max_host) return set_error(ctx, REACH_EINVAL, "batch_max_over_steps: bad split");
    REACH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf v;
    REACH_TRY(v.alloc(ctx, size_t(b->ndirs) * 8, false));
    // split = -1: rho(d, MixedFlowpipe) = max over every reach-set of every split (MixedFlowpipe.jl:66-68);
    // [split][step][dir] is one ndirs x (nsplits*nsteps) column-major matrix
    const double* src = b->rho.as<double>() + (split < 0 ? 0 : size_t(split) * b->nsteps * b->ndirs);
    const long long cols = split < 0 ? (long long)b->nsplits * b->nsteps : b->nsteps;
    max_over_steps_kernel<<<(b->ndirs + 127) / 128, 128, 0, st>>>(src, b->ndirs, cols, v.as<double>(), nullptr);
    REACH_CUDA(ctx, cudaGetLastError());
    REACH_CUDA(ctx, cudaMemcpyAsync(max_host, v.p, size_t(b->ndirs) * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(ctx, cudaStreamSynchronize(st));
    return REACH_OK;
}

extern "C" int reach_lgg09_batch_stats(reach_lgg09_batch* b, double* ms, int64_t* launches, int32_t* time_block,
                                       int32_t* alternatives, int32_t* feature_cols, int64_t* chunk_steps) {
    if (!b) return REACH_EINVAL;
    if (ms) {
        *ms = 0.0;
        if (b->ran) {
            REACH_CUDA(b->ctx, cudaEventSynchronize(b->ev1));
            float f = 0.f;
            REACH_CUDA(b->ctx, cudaEventElapsedTime(&f, b->ev0, b->ev1));
            *ms = f;
        }
    }
    if (launches) *launches = b->launches;
    if (time_block) *time_block = b->S;
    if (alternatives) *alternatives = b->na;
    if (feature_cols) *feature_cols = b->Kc;
    if (chunk_steps) *chunk_steps = b->chunk;
    return REACH_OK;
}

extern "C" void* reach_lgg09_batch_device_rho(reach_lgg09_batch* b) { return b ? b->rho.p : nullptr; }

extern "C" void reach_lgg09_batch_destroy(reach_lgg09_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    delete b;
}
