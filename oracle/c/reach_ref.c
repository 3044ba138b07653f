/* reach_ref.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Never linked into, loaded by or called from the product
 * (libreach_cuda.so, the reachabilityanalysis.jl_b200 package): only tests/, __graft_entry__.smoke() and the CPU legs
 * of bench.py use it, as the checker and as the timed CPU arm.
 *
 * Plain C + OpenMP restatement of the reference's CPU loops for the hot path.  The reference is pure Julia and Julia is
 * not installed, so this is a restatement ("port"), not the reference itself; it is pinned against the NumPy oracle
 * (the oracle .py modules), which reproduces the reference's own known-answer tests (tests/test_oracle_kat.py).
 *
 *   ref_lgg09 shape 1 ("B1"): per-direction chain  r_{i+1} = Phi' r_i, directions spread over the cores
 *                             src/Algorithms/LGG09/reach_inhomog.jl:35-47 (threaded dispatch), :49-64 (loop),
 *                             reach_homog.jl:35-47,49-62; optional sparse Phi' as LGG09/post.jl:25-27 (`sparse=true`)
 *             shape 2 ("B2"): one GEMM per step over the whole direction block + a threaded support epilogue
 *                             reach_inhomog.jl:174-218 / 220-234, reach_homog.jl:382-438 (the invariant variants)
 *   support function rules  : LazySets (third party, Project.toml:34): Hyperrectangle sum d_i (c_i +- r_i) by the sign
 *                             of d_i, Zonotope d.c + sum |g_j.d|, LinearMap rho(M'd, X), MinkowskiSum = sum,
 *                             ConvexHull = max -- given here as the flattened program of include/reach.h
 *   ref_glgm06 ("B3")       : src/Algorithms/GLGM06/reach_inhomog.jl:6-43 per initial set, sets spread over the cores
 *                             (src/Continuous/solve.jl:107-117), linear_map_minkowski_sum Continuous/setops.jl:7-25,
 *                             reduce_order(., r, GIR05()) (LazySets)
 *
 * Two deliberate strengthenings over the literal reference, so that the CPU arm is a competent baseline rather than a
 * straw man (both switchable): (1) `literal == 0` evaluates terms under LinearMap(Phi, .) at the already computed
 * r_{i+1} instead of a second Phi' r product inside rho; (2) `share != 0` lets directions l and -l share one chain
 * (noted at reach_homog.jl:108-110).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/reach.h" /* struct layouts of the flattened support program only */

typedef void (*ref_dgemm_fn)(int order, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                             const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                             int64_t ldc);

/* ---- support function of one flattened term at d (length n), sign sg = +-1 applied to d ---- */
static double term_rho(const reach_term* t, int n, const double* d, double sg, double* scratch) {
    const double* v = d;
    int dim = n;
    double s = 0.0;
    if (t->q > 0) { /* rho(M'd, X) */
        for (int j = 0; j < t->q; ++j) {
            const double* col = t->M + (size_t)j * n;
            double acc = 0.0;
            for (int i = 0; i < n; ++i) acc += col[i] * d[i];
            scratch[j] = acc;
        }
        v = scratch;
        dim = t->q;
    }
    if (t->kind == REACH_TERM_BOX) {
        if (t->c && t->r) {
            for (int i = 0; i < dim; ++i) { const double x = sg * v[i]; s += x * (t->c[i] + copysign(t->r[i], x)); }
        } else if (t->c) {
            for (int i = 0; i < dim; ++i) s += sg * v[i] * t->c[i];
        } else if (t->r) {
            for (int i = 0; i < dim; ++i) s += fabs(v[i]) * t->r[i];
        }
        return s;
    }
    if (t->c)
        for (int i = 0; i < dim; ++i) s += sg * v[i] * t->c[i];
    for (int g = 0; g < t->p; ++g) {
        const double* col = t->G + (size_t)g * dim;
        double acc = 0.0;
        for (int i = 0; i < dim; ++i) acc += col[i] * v[i];
        s += fabs(acc);
    }
    return s;
}

static int max_q(const reach_setexpr* e) {
    int q = 1;
    if (!e) return q;
    int nt = 0;
    for (int a = 0; a < e->nalt; ++a) nt += e->alt_len ? e->alt_len[a] : 0;
    for (int i = 0; i < nt; ++i)
        if (e->terms[i].q > q) q = e->terms[i].q;
    return q;
}

/* rho(sg*r; set): max over alternatives of the sum of the terms; which == 1 terms are evaluated at rn = Phi' r */
static double set_rho(const reach_setexpr* e, int n, const double* r, const double* rn, double sg, double* scratch) {
    const reach_term* t = e->terms;
    double best = 0.0;
    for (int a = 0; a < e->nalt; ++a) {
        double s = 0.0;
        const int len = e->alt_len ? e->alt_len[a] : 0;
        for (int i = 0; i < len; ++i, ++t) s += term_rho(t, n, t->which ? rn : r, sg, scratch);
        if (a == 0 || s > best) best = s;
    }
    return best;
}

static int needs_next(const reach_setexpr* e) {
    int nt = 0;
    for (int a = 0; a < e->nalt; ++a) nt += e->alt_len ? e->alt_len[a] : 0;
    for (int i = 0; i < nt; ++i)
        if (e->terms[i].which) return 1;
    return 0;
}

static int is_zero_set(const reach_setexpr* e) {
    for (int a = 0; a < e->nalt; ++a)
        if (e->alt_len && e->alt_len[a]) return 0;
    return 1;
}

typedef struct { /* Phi' in CSR form (= Phi column-major in CSC form) */
    int64_t* ptr;
    int* idx;
    double* val;
} csr_t;

static void matvec_T(int n, const double* Phi, const csr_t* sp, const double* r, double* out) {
    if (sp) {
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int64_t z = sp->ptr[i]; z < sp->ptr[i + 1]; ++z) acc += sp->val[z] * r[sp->idx[z]];
            out[i] = acc;
        }
        return;
    }
    for (int i = 0; i < n; ++i) { /* (Phi' r)_i = column i of Phi . r */
        const double* col = Phi + (size_t)i * n;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int l = 0;
        for (; l + 3 < n; l += 4) { a0 += col[l] * r[l]; a1 += col[l + 1] * r[l + 1]; a2 += col[l + 2] * r[l + 2]; a3 += col[l + 3] * r[l + 3]; }
        for (; l < n; ++l) a0 += col[l] * r[l];
        out[i] = (a0 + a1) + (a2 + a3);
    }
}

static void naive_gemm_T(int n, int m, const double* Phi, const double* R, double* Rn) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < m; ++j) matvec_T(n, Phi, NULL, R + (size_t)j * n, Rn + (size_t)j * n);
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* rho: ndirs x nsteps column-major, rho[j + ndirs*k] = rho_l[j+1, k+1].  V == NULL: homogeneous.
 * shape: 1 = B1, 2 = B2.  sparse: Phi' held in CSR form (B1 only).  literal: see the header (B1 only).
 * share: directions j and j + ndirs/2 must be exact negations (box-like templates); one chain serves both.
 * Returns 0, or -1 on a bad argument / allocation failure. */
int ref_lgg09(int shape, int n, int ndirs, int64_t nsteps, const double* Phi, int sparse, const double* dirs,
              const reach_setexpr* O, const reach_setexpr* V, int literal, int share, double* rho, int nthreads,
              ref_dgemm_fn gemm) {
    if (n < 1 || ndirs < 1 || nsteps < 1 || !Phi || !dirs || !O || !rho) return -1;
    if (shape != 1 && shape != 2) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int nn = needs_next(O);
    const int zero0 = is_zero_set(O);
    int nchains = ndirs;
    if (share) {
        if (ndirs % 2) return -1;
        nchains = ndirs / 2;
        for (int j = 0; j < nchains; ++j)
            for (int i = 0; i < n; ++i)
                if (dirs[(size_t)j * n + i] != -dirs[(size_t)(j + nchains) * n + i]) return -1;
    }
    int mq = max_q(O);
    if (V && max_q(V) > mq) mq = max_q(V);

    csr_t sp = {0, 0, 0};
    if (sparse) {
        int64_t nnz = 0;
        for (size_t i = 0; i < (size_t)n * n; ++i) nnz += Phi[i] != 0.0;
        sp.ptr = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
        sp.idx = (int*)malloc(sizeof(int) * (nnz ? nnz : 1));
        sp.val = (double*)malloc(sizeof(double) * (nnz ? nnz : 1));
        if (!sp.ptr || !sp.idx || !sp.val) return -1;
        int64_t z = 0;
        for (int i = 0; i < n; ++i) {
            sp.ptr[i] = z;
            for (int l = 0; l < n; ++l)
                if (Phi[(size_t)i * n + l] != 0.0) { sp.idx[z] = l; sp.val[z] = Phi[(size_t)i * n + l]; ++z; }
        }
        sp.ptr[n] = z;
    }

    int rc = 0;
    if (shape == 1) {
#pragma omp parallel
        {
            double* r = (double*)malloc(sizeof(double) * n);
            double* rn = (double*)malloc(sizeof(double) * n);
            double* tmp = (double*)malloc(sizeof(double) * n);
            double* scratch = (double*)malloc(sizeof(double) * mq);
#pragma omp for schedule(dynamic, 1)
            for (int j = 0; j < nchains; ++j) {
                if (!r || !rn || !tmp || !scratch) { rc = -1; continue; }
                memcpy(r, dirs + (size_t)j * n, sizeof(double) * n);
                double sp_ = 0.0, sn_ = 0.0; /* running input sums of +r and -r */
                for (int64_t k = 0; k < nsteps; ++k) {
                    /* r_{k+1} = Phi' r_k (reach_inhomog.jl:60); computed first so that `which == 1` terms can use it */
                    matvec_T(n, Phi, sparse ? &sp : NULL, r, rn);
                    const double* rnext = rn;
                    if (literal && nn) { /* LazySets' LinearMap rule: its own Phi' d product inside rho */
                        matvec_T(n, Phi, sparse ? &sp : NULL, r, tmp);
                        rnext = tmp;
                    }
                    rho[j + (size_t)ndirs * k] = (zero0 ? 0.0 : set_rho(O, n, r, rnext, 1.0, scratch)) + sp_;      /* :56 */
                    if (V) sp_ += set_rho(V, n, r, rnext, 1.0, scratch);                                           /* :57 */
                    if (share) {
                        rho[j + nchains + (size_t)ndirs * k] = (zero0 ? 0.0 : set_rho(O, n, r, rnext, -1.0, scratch)) + sn_;
                        if (V) sn_ += set_rho(V, n, r, rnext, -1.0, scratch);
                    }
                    double* t = r; r = rn; rn = t;
                }
            }
            free(r); free(rn); free(tmp); free(scratch);
        }
    } else {
        double* R = (double*)malloc(sizeof(double) * (size_t)n * nchains);
        double* Rn = (double*)malloc(sizeof(double) * (size_t)n * nchains);
        double* s = (double*)calloc((size_t)ndirs, sizeof(double));
        const int nth = ref_num_threads();
        double* scratch_all = (double*)malloc(sizeof(double) * (size_t)mq * nth);
        if (!R || !Rn || !s || !scratch_all) { free(R); free(Rn); free(s); free(scratch_all); return -1; }
        memcpy(R, dirs, sizeof(double) * (size_t)n * nchains);
        for (int64_t k = 0; k < nsteps; ++k) {
            /* mul!(r_{i+1}, Phi', r_i) over the whole block (reach_inhomog.jl:206) */
            if (gemm) gemm(102 /*ColMajor*/, 112 /*Trans*/, 111 /*NoTrans*/, n, nchains, n, 1.0, Phi, n, R, n, 0.0, Rn, n);
            else naive_gemm_T(n, nchains, Phi, R, Rn);
#pragma omp parallel for schedule(static)
            for (int j = 0; j < nchains; ++j) {
#ifdef _OPENMP
                double* scratch = scratch_all + (size_t)mq * omp_get_thread_num();
#else
                double* scratch = scratch_all;
#endif
                const double* r = R + (size_t)j * n;
                const double* rn = Rn + (size_t)j * n;
                rho[j + (size_t)ndirs * k] = (zero0 ? 0.0 : set_rho(O, n, r, rn, 1.0, scratch)) + s[j];
                if (V) s[j] += set_rho(V, n, r, rn, 1.0, scratch);
                if (share) {
                    const int jn = j + nchains;
                    rho[jn + (size_t)ndirs * k] = (zero0 ? 0.0 : set_rho(O, n, r, rn, -1.0, scratch)) + s[jn];
                    if (V) s[jn] += set_rho(V, n, r, rn, -1.0, scratch);
                }
            }
            double* t = R; R = Rn; Rn = t;
        }
        free(R); free(Rn); free(s); free(scratch_all);
    }
    free(sp.ptr); free(sp.idx); free(sp.val);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------------------
 * GLGM06
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { double w; int idx; } wkey;

static int wkey_cmp(const void* a, const void* b) { /* (weight, index) ascending == stable sortperm */
    const wkey* x = (const wkey*)a;
    const wkey* y = (const wkey*)b;
    if (x->w < y->w) return -1;
    if (x->w > y->w) return 1;
    return x->idx - y->idx;
}

/* reduce_order(Zonotope(., G), r, GIR05()): G n x p in, Gout n x pout (capacity >= max(p, r n)).  sel (capacity p) gets
 * the kept source columns in output order; returns pout.  keys: scratch of p entries. */
static int gir05(int n, int p, int r, const double* G, double* Gout, int* sel, int* nsel, wkey* keys) {
    if ((int64_t)r * n >= p) {
        if (Gout != G) memcpy(Gout, G, sizeof(double) * (size_t)n * p);
        if (nsel) *nsel = -1; /* unchanged */
        return p;
    }
    if (r == 1) {
        double* diag = (double*)calloc((size_t)n, sizeof(double));
        for (int j = 0; j < p; ++j)
            for (int i = 0; i < n; ++i) diag[i] += fabs(G[(size_t)j * n + i]);
        memset(Gout, 0, sizeof(double) * (size_t)n * n);
        for (int i = 0; i < n; ++i) Gout[(size_t)i * n + i] = diag[i];
        free(diag);
        if (nsel) *nsel = 0;
        return n;
    }
    for (int j = 0; j < p; ++j) {
        double s1 = 0.0, mx = 0.0;
        for (int i = 0; i < n; ++i) { const double a = fabs(G[(size_t)j * n + i]); s1 += a; if (a > mx) mx = a; }
        keys[j].w = s1 - mx;
        keys[j].idx = j;
    }
    qsort(keys, (size_t)p, sizeof(wkey), wkey_cmp);
    const int m = p - n * (r - 1);
    double* diag = (double*)calloc((size_t)n, sizeof(double));
    for (int t = 0; t < m; ++t) { /* box the m smallest, summed in sorted order */
        const double* g = G + (size_t)keys[t].idx * n;
        for (int i = 0; i < n; ++i) diag[i] += fabs(g[i]);
    }
    const int kept = p - m;
    /* Gout may alias G only if the caller passes a different buffer; we require Gout != G here */
    for (int t = 0; t < kept; ++t) {
        memcpy(Gout + (size_t)t * n, G + (size_t)keys[m + t].idx * n, sizeof(double) * n);
        if (sel) sel[t] = keys[m + t].idx;
    }
    memset(Gout + (size_t)kept * n, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i) Gout[(size_t)(kept + i) * n + i] = diag[i];
    free(diag);
    if (nsel) *nsel = kept;
    return kept + n;
}

int ref_reduce_order_gir05(int n, int p, int r, const double* G, double* Gout, int* sel, int* nsel) {
    wkey* keys = (wkey*)malloc(sizeof(wkey) * (size_t)(p > 0 ? p : 1));
    if (!keys) return -1;
    const int out = gir05(n, p, r, G, Gout, sel, nsel, keys);
    free(keys);
    return out;
}

static void matmul_cm(int n, int p, const double* P, const double* G, double* out) { /* out (n x p) = P (n x n) G */
    for (int j = 0; j < p; ++j) {
        double* o = out + (size_t)j * n;
        for (int i = 0; i < n; ++i) o[i] = 0.0;
        for (int l = 0; l < n; ++l) {
            const double g = G[(size_t)j * n + l];
            const double* pc = P + (size_t)l * n;
            for (int i = 0; i < n; ++i) o[i] += pc[i] * g;
        }
    }
}

static double zono_rho(int n, int p, const double* d, const double* c, const double* G) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += d[i] * c[i];
    for (int j = 0; j < p; ++j) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += G[(size_t)j * n + i] * d[i];
        s += fabs(acc);
    }
    return s;
}

/* Inhomogeneous GLGM06 over nsplits initial zonotopes (c0 n x nsplits, G0 n x p0 x nsplits, already reduced as
 * post.jl:26 does) sharing Phi and U = (cU, GU n x pU).  Step 0 is Omega0 itself (reach_inhomog.jl:20).
 *   d, rho_out    : support value rho(d, Z_{split,step}) of every zonotope, nsplits x nsteps (split fastest); may be NULL
 *   nsamp, samp_* : zonotopes to return in full: (split, step) pairs -> samp_c (n each), samp_G (n x gcap each),
 *                   samp_ngens, samp_sel (gcap each: source column in [P G0 | W] numbering, -1 for the diagonal;
 *                   all -2 when the zonotope was not reduced)
 *   shared_chain  : 0 = every split recomputes P_k and W_k, as one `post` call per split does (solve.jl:107-117);
 *                   1 = computed once and shared (a competent CPU implementation)
 * Returns 0 or -1. */
int ref_glgm06(int n, int64_t nsteps, int nsplits, int max_order, const double* Phi, const double* c0, const double* G0,
               int p0, const double* cU, const double* GU, int pU, const double* d, double* rho_out, int nsamp,
               const int* samp_split, const int64_t* samp_step, double* samp_c, double* samp_G, int gcap,
               int* samp_ngens, int* samp_sel, int shared_chain, int nthreads) {
    if (n < 1 || nsteps < 1 || nsplits < 1 || max_order < 1 || p0 < 0 || pU < 0 || !Phi || !c0 || !cU) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int capW = max_order * n + pU + n; /* generators of W never exceed max(pU, r n) (+ slack) */
    const int capR = p0 + capW;
    const size_t nn = (size_t)n * n;
    /* shared chain storage: P_k (k = 1..nsteps-1) and W_k as they enter step k */
    double *Pall = NULL, *Wc = NULL, *WG = NULL;
    int* Wp = NULL;
    if (shared_chain) {
        Pall = (double*)malloc(sizeof(double) * nn * (size_t)nsteps);
        Wc = (double*)malloc(sizeof(double) * (size_t)n * nsteps);
        WG = (double*)malloc(sizeof(double) * (size_t)n * capW * nsteps);
        Wp = (int*)malloc(sizeof(int) * (size_t)nsteps);
        if (!Pall || !Wc || !WG || !Wp) { free(Pall); free(Wc); free(WG); free(Wp); return -1; }
        double* P = (double*)malloc(sizeof(double) * nn);
        double* Pn = (double*)malloc(sizeof(double) * nn);
        double* cw = (double*)malloc(sizeof(double) * n);
        double* Gw = (double*)malloc(sizeof(double) * (size_t)n * capW);
        double* Gt = (double*)malloc(sizeof(double) * (size_t)n * (capW + pU));
        wkey* keys = (wkey*)malloc(sizeof(wkey) * (size_t)(capW + pU));
        memcpy(P, Phi, sizeof(double) * nn);
        memcpy(cw, cU, sizeof(double) * n);
        if (pU) memcpy(Gw, GU, sizeof(double) * (size_t)n * pU);
        int pw = pU;
        for (int64_t k = 1; k < nsteps; ++k) {
            memcpy(Pall + nn * k, P, sizeof(double) * nn);
            memcpy(Wc + (size_t)n * k, cw, sizeof(double) * n);
            memcpy(WG + (size_t)n * capW * k, Gw, sizeof(double) * (size_t)n * pw);
            Wp[k] = pw;
            /* W <- reduce(P U (+) W)  (:35-36) */
            matmul_cm(n, pU, P, GU, Gt);
            memcpy(Gt + (size_t)n * pU, Gw, sizeof(double) * (size_t)n * pw);
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc += P[(size_t)l * n + i] * cU[l];
                cw[i] = acc + cw[i];
            }
            pw = gir05(n, pU + pw, max_order, Gt, Gw, NULL, NULL, keys);
            matmul_cm(n, n, P, Phi, Pn); /* P <- P Phi (:38) */
            double* t = P; P = Pn; Pn = t;
        }
        free(P); free(Pn); free(cw); free(Gw); free(Gt); free(keys);
    }
    int rc = 0;
#pragma omp parallel
    {
        double* P = (double*)malloc(sizeof(double) * nn);
        double* Pn = (double*)malloc(sizeof(double) * nn);
        double* cw = (double*)malloc(sizeof(double) * n);
        double* Gw = (double*)malloc(sizeof(double) * (size_t)n * capW);
        double* Gt = (double*)malloc(sizeof(double) * (size_t)n * capR);
        double* Gr = (double*)malloc(sizeof(double) * (size_t)n * capR);
        double* cr = (double*)malloc(sizeof(double) * n);
        int* sel = (int*)malloc(sizeof(int) * (size_t)capR);
        wkey* keys = (wkey*)malloc(sizeof(wkey) * (size_t)capR);
#pragma omp for schedule(dynamic, 1)
        for (int s = 0; s < nsplits; ++s) {
            if (!P || !Pn || !cw || !Gw || !Gt || !Gr || !cr || !sel || !keys) { rc = -1; continue; }
            const double* cs = c0 + (size_t)n * s;
            const double* Gs = G0 + (size_t)n * p0 * s;
            int pw = pU;
            if (!shared_chain) {
                memcpy(P, Phi, sizeof(double) * nn);
                memcpy(cw, cU, sizeof(double) * n);
                if (pU) memcpy(Gw, GU, sizeof(double) * (size_t)n * pU);
            }
            for (int64_t k = 0; k < nsteps; ++k) {
                const double *cz, *Gz;
                int pz, nsel = -2;
                if (k == 0) {
                    cz = cs; Gz = Gs; pz = p0;
                } else {
                    const double* Pk = shared_chain ? Pall + nn * k : P;
                    const double* cwk = shared_chain ? Wc + (size_t)n * k : cw;
                    const double* Gwk = shared_chain ? WG + (size_t)n * capW * k : Gw;
                    const int pwk = shared_chain ? Wp[k] : pw;
                    /* R = reduce(P Omega0 (+) W)  (:29-30): c = P c0 + cW, G = [P G0 | GW] */
                    matmul_cm(n, p0, Pk, Gs, Gt);
                    memcpy(Gt + (size_t)n * p0, Gwk, sizeof(double) * (size_t)n * pwk);
                    for (int i = 0; i < n; ++i) {
                        double acc = 0.0;
                        for (int l = 0; l < n; ++l) acc += Pk[(size_t)l * n + i] * cs[l];
                        cr[i] = acc + cwk[i];
                    }
                    pz = gir05(n, p0 + pwk, max_order, Gt, Gr, sel, &nsel, keys);
                    cz = cr; Gz = Gr;
                    if (!shared_chain) {
                        matmul_cm(n, pU, P, GU, Gt);
                        memcpy(Gt + (size_t)n * pU, Gw, sizeof(double) * (size_t)n * pw);
                        for (int i = 0; i < n; ++i) {
                            double acc = 0.0;
                            for (int l = 0; l < n; ++l) acc += P[(size_t)l * n + i] * cU[l];
                            cw[i] = acc + cw[i];
                        }
                        pw = gir05(n, pU + pw, max_order, Gt, Gw, NULL, NULL, keys);
                        matmul_cm(n, n, P, Phi, Pn);
                        double* t = P; P = Pn; Pn = t;
                    }
                }
                if (d && rho_out) rho_out[s + (size_t)nsplits * k] = zono_rho(n, pz, d, cz, Gz);
                for (int q = 0; q < nsamp; ++q) {
                    if (samp_split[q] != s || samp_step[q] != k) continue;
                    if (pz > gcap) { rc = -1; continue; }
                    memcpy(samp_c + (size_t)n * q, cz, sizeof(double) * n);
                    memcpy(samp_G + (size_t)n * gcap * q, Gz, sizeof(double) * (size_t)n * pz);
                    samp_ngens[q] = pz;
                    for (int t = 0; t < gcap; ++t) samp_sel[(size_t)gcap * q + t] = -2;
                    if (nsel >= 0) {
                        for (int t = 0; t < nsel; ++t) samp_sel[(size_t)gcap * q + t] = sel[t];
                        for (int t = nsel; t < pz; ++t) samp_sel[(size_t)gcap * q + t] = -1;
                    }
                }
            }
        }
        free(P); free(Pn); free(cw); free(Gw); free(Gt); free(Gr); free(cr); free(sel); free(keys);
    }
    free(Pall); free(Wc); free(WG); free(Wp);
    return rc;
}
