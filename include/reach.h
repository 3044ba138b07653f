/* reach.h -- C ABI of libreach_cuda.so
 *
 * B200-native (sm_100a) implementation of the linear set-propagation hot path of
 * ReachabilityAnalysis.jl: the LGG09 template support-function recurrence and the GLGM06
 * zonotope flowpipe.  The reference has no FFI for this path (it is pure Julia); the seam this
 * library replaces is the inside of the four loops below, which become `ccall`s:
 *
 *   reach_lgg09 / reach_lgg09_plan_*  <-  reach_homog_LGG09!   src/Algorithms/LGG09/reach_homog.jl:6-33,49-62
 *                                         reach_inhomog_LGG09! src/Algorithms/LGG09/reach_inhomog.jl:5-33,49-81
 *                                         (invariant variants reach_homog.jl:382-424, reach_inhomog.jl:174-218)
 *   reach_glgm06 / reach_glgm06_plan_* <- reach_homog_GLGM06!   src/Algorithms/GLGM06/reach_homog.jl:33-73
 *                                         reach_inhomog_GLGM06! src/Algorithms/GLGM06/reach_inhomog.jl:6-43
 *                                         reduce_order(., GIR05) call sites GLGM06/post.jl:26, reach_inhomog.jl:30,36
 *                                         linear_map_minkowski_sum src/Continuous/setops.jl:7-25
 *   reach_*_max_over_steps / fetch     <- rho(d, fp)  src/Flowpipes/AbstractFlowpipe.jl:35-37,
 *                                         support_function_matrix src/Flowpipes/Flowpipe.jl:304-306
 *
 * Conventions
 *   - plain C: pointers, sizes, no C++/torch types.  All matrices are Float64, COLUMN-MAJOR
 *     (Julia layout) unless stated.
 *   - every function returns 0 on success or a negative REACH_E* code; the message is available
 *     from reach_last_error().  The library never aborts and has no CPU fallback: without a
 *     usable CUDA device reach_create fails.
 *   - host pointers are borrowed for the duration of the call only.
 *   - one reach_ctx per calling thread (or lock externally); a ctx owns one device + one stream;
 *     reach_multi_* drives several devices from one handle (below).
 *   - time stamps (dt, dt0), TemplateReachSet / ReachSet construction and the :sfmat dictionary
 *     stay on the caller's side.
 */
#ifndef REACH_H
#define REACH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REACH_OK            0
#define REACH_EINVAL      (-1)  /* bad argument (maps to ArgumentError)            */
#define REACH_ECUDA       (-2)  /* CUDA runtime / driver error                      */
#define REACH_ENOMEM      (-3)  /* host or device allocation failed                 */
#define REACH_EUNSUPPORTED (-4) /* valid request the device path does not implement */
#define REACH_ESTATE      (-5)  /* call order violated (e.g. fetch before run)      */

typedef struct reach_ctx reach_ctx;

int         reach_version(void);                          /* 10000*major + 100*minor + patch */
int         reach_device_count(void);                     /* <0 on error                      */
int         reach_create(reach_ctx** out, int device);
void        reach_destroy(reach_ctx* ctx);
const char* reach_last_error(const reach_ctx* ctx);       /* ctx may be NULL: last create error */
int         reach_device_name(const reach_ctx* ctx, char* buf, size_t len);
int         reach_sm_count(const reach_ctx* ctx);
/* stream the ctx launches on (a cudaStream_t), for callers that interleave their own work */
void*       reach_stream(const reach_ctx* ctx);
int         reach_synchronize(reach_ctx* ctx);

/* ---------------------------------------------------------------------------------------------
 * Flattened support-function program.
 *
 * The lazy set trees that `discretize` hands to `post` (ConvexHull / MinkowskiSum / LinearMap of
 * Hyperrectangle, BallInf, Interval, Singleton, ZeroSet, Zonotope -- ForwardModule.jl:137-164)
 * are flattened by the caller into
 *
 *      rho(r_k; Omega0) = max over alternatives a of  sum over terms t in a of  term_t
 *      rho(r_k; V)      =                             sum over terms t       of  term_t
 *
 * Each term is evaluated at d = r_k (which = 0) or d = r_{k+1} = Phi' r_k (which = 1, i.e. the
 * term sat under `LinearMap(Phi, .)`), optionally after a further linear map d' = M' d:
 *
 *      REACH_TERM_BOX  : sum_i d'_i c_i + |d'_i| r_i            (c, r may be NULL = 0)
 *      REACH_TERM_ZONO : d'.c + sum_j |g_j . d'|                (G is dim x p, column-major)
 *
 * An alternative with zero terms evaluates to 0 (Omega0 = ZeroSet, reach_inhomog.jl:66-81).
 * ------------------------------------------------------------------------------------------- */
#define REACH_TERM_BOX  0
#define REACH_TERM_ZONO 1

typedef struct {
    int32_t kind;        /* REACH_TERM_BOX | REACH_TERM_ZONO                                  */
    int32_t which;       /* 0: r_k, 1: r_{k+1}; must be 0 inside V                            */
    int32_t q;           /* 0: the set lives in R^n; >0: it lives in R^q and M is n x q       */
    int32_t p;           /* REACH_TERM_ZONO: number of generators                             */
    const double* M;     /* n x q column-major, or NULL when q == 0                            */
    const double* c;     /* centre, length (q ? q : n), or NULL                                */
    const double* r;     /* REACH_TERM_BOX radius, length (q ? q : n), or NULL                 */
    const double* G;     /* REACH_TERM_ZONO generators, (q ? q : n) x p column-major          */
} reach_term;

typedef struct {
    int32_t           nalt;     /* >= 1                                                        */
    const int32_t*    alt_len;  /* nalt entries; terms of alternative a follow those of a-1    */
    const reach_term* terms;    /* sum(alt_len) entries                                        */
} reach_setexpr;

/* Invariant for the early-terminating variants (the `X` argument of reach_*_LGG09!, reach_homog.jl:382-424,
 * reach_inhomog.jl:174-218: `_isdisjoint(X, set(F[k])) && break`):
 *   REACH_INV_NONE       Universe
 *   REACH_INV_HALFSPACE  {x : a.x <= b}                       a: normal (n), b: offset (1)
 *   REACH_INV_BOX        [lo, hi]                             a: lo (n), b: hi (n); +-inf allowed
 *   REACH_INV_POLYHEDRON {x : A x <= b}, m faces              a: A row-major m x n (face i = a + i*n), b: offsets (m)
 * The template reach-set is P_k = {x : d_j.x <= rho_j[k]}.  A face {a.x <= b} is disjoint from P_k iff
 * min_{P_k} a.x > b.  The device certifies this from the template's own support values by weak duality
 * (sum_j lambda_j rho_j[k] < -b for a fixed lambda >= 0 with sum_j lambda_j d_j = -a), with lambda chosen at upload:
 *   (1) -a is a positive multiple of a template direction          -> exact for that face;
 *   (2) the +-e_i rows of the template that a's non-zeros need    -> exact when the template is a box template
 *       (only +-e_i directions, all of them) or a has one non-zero, otherwise sufficient (the box hull of P_k);
 *   faces with neither are not tested.  If NO face can be tested, plan_upload fails with REACH_EUNSUPPORTED.
 * The result equals the reference's exact polyhedral test when reach_lgg09_plan_invariant_info reports exact = 1:
 * a single half-space with an exact certificate, or a box invariant with a box template.  In the other cases a step
 * is only reported disjoint when it really is (never a false stop), but the run may stop later than the reference.
 * The run stops launching device work once a step is certified disjoint (the triggering step is dropped, like
 * `resize!(F, k - 1)`). */
#define REACH_INV_NONE       0
#define REACH_INV_HALFSPACE  1
#define REACH_INV_BOX        2
#define REACH_INV_POLYHEDRON 3
typedef struct {
    int32_t kind;
    int32_t m;         /* POLYHEDRON: number of faces (ignored otherwise; sits in what was padding) */
    const double* a;   /* HALFSPACE: normal (n)      BOX: lo (n)      POLYHEDRON: A, row-major m x n */
    const double* b;   /* HALFSPACE: offset (1)      BOX: hi (n)      POLYHEDRON: offsets (m)        */
} reach_invariant;

typedef struct {
    int32_t time_block;   /* S: steps advanced per device pass via the power stack (Phi')^1..S.
                             0 = choose automatically; 1 = strictly sequential chain (the
                             reference's association, r_{k+1} = Phi' r_k)                     */
    int32_t share_negated;/* 1 (default when zeroed struct is NOT used: see reach_lgg09_default_opts):
                             directions l and -l share the chain (Phi')^k l
                             (noted at reach_homog.jl:108-110); results are bit-identical     */
    int32_t tile_m;       /* 0 = auto; else force the direction-tile size (testing)           */
    int32_t tile_n;       /* 0 = auto; else force the row-tile size (testing)                 */
    int32_t path;         /* 0 = auto; 1 = batched FP64 (DMMA) GEMM over the power stack; 2 = persistent
                             warp-per-chain kernel (Phi' in shared memory, ELL form: small or
                             sparse Phi, e.g. ISS -- the device analogue of `sparse=true`,
                             LGG09/post.jl:25-27; strictly sequential association); 3 = the same
                             batched GEMM on the INT8 tcgen05 pipe through error-free digit planes
                             (unguarded when forced; auto picks it for large dense problems and
                             guards it against the DMMA kernel, see reach_lgg09_plan_path)      */
    int32_t reserved[3];  /* reserved[0]: FP64 emulation of path 3.  0 = default: residues modulo pairwise coprime
                             moduli, one INT8 product per modulus and FP64 product (csrc/crt_gemm.cuh): 14 moduli
                             when path 3 is forced; on the guarded auto path 13, and the run restarts at pass 0
                             with 14 if the first guard check exceeds 2.5e-13 (REACH_CRT_NO_ADAPT=1: always 14);
                             10..16 = that many moduli (13 matches the digit planes' accuracy, 12 is
                             below it); 4..8 = that many 8-bit digit planes (csrc/oz_gemm.cuh, 7 = 56 bits,
                             28 INT8 products per FP64 product).  Anything else: REACH_EINVAL.
                             reserved[1]: 1 = box output: the two rows of a direction pair (l, -l) hold
                             (lin, abs) = (centre, radius) of the box hull instead of abs +- lin
                             (what reach_box uses; Omega0 must have one alternative)            */
} reach_lgg09_opts;

void reach_lgg09_default_opts(reach_lgg09_opts* o);

typedef struct reach_lgg09_plan reach_lgg09_plan;

/* One-shot, host buffers in, host buffer out: what `reach_homog_LGG09!` / `reach_inhomog_LGG09!`
 * become.  rho_host is ndirs x nsteps column-major: rho_host[j + ndirs*k] = rho_l[j+1, k+1].
 * V == NULL  <=>  homogeneous.  inv == NULL <=> Universe.  *nsteps_done (may be NULL) receives
 * the number of valid columns (< nsteps only if an invariant stopped the run early; the column
 * that triggered the stop is not counted, as `resize!(F, k-1)` drops it). */
int reach_lgg09(reach_ctx* ctx, int n, int ndirs, int64_t nsteps,
                const double* Phi,              /* n x n column-major, NOT transposed          */
                const double* dirs,             /* n x ndirs column-major; row order of output */
                const reach_setexpr* Omega0, const reach_setexpr* V,
                const reach_invariant* inv, const reach_lgg09_opts* opts,
                double* rho_host, int64_t* nsteps_done);

/* Plan API: device-resident inputs and outputs (device-resident Flowpipe storage). */
int  reach_lgg09_plan_create(reach_ctx* ctx, int n, int ndirs, int64_t nsteps,
                             const reach_lgg09_opts* opts, reach_lgg09_plan** out);
int  reach_lgg09_plan_upload(reach_lgg09_plan* p, const double* Phi, const double* dirs,
                             const reach_setexpr* Omega0, const reach_setexpr* V,
                             const reach_invariant* inv);
/* If set before run, rho is written to this caller-owned DEVICE buffer (ndirs x nsteps doubles,
 * column-major) instead of the plan's own; lets a caller hand in memory it will feed to NCCL. */
int  reach_lgg09_plan_set_output(reach_lgg09_plan* p, void* rho_device);
int  reach_lgg09_plan_run(reach_lgg09_plan* p);            /* asynchronous on the ctx stream  */
/* The row stack [I;E](Phi')^b, b = 0..S, of a GEMM-path plan in steps, for callers that split its rows over several GPUs
 * (one plan per GPU, same Phi and time block): step 0 is block 1, step j >= 1 the doubling step s = 2^(j-1).  With
 * row_lo < 0 the call only reports the step's output rows (*step_rows, 0 past the last step) and where they start in the
 * stack (*dest_row); otherwise it computes rows [row_lo, row_hi) of them (asynchronously on the ctx stream).  Every caller
 * takes every step in order, all-gathers the rows in place (plan_stack_info: device buffer, rows per block, leading
 * dimension in doubles; at most 256 rows past the step's rows may be overwritten) and calls plan_stack_ready; the next
 * plan_run then skips its own stack build.  Same products in the same order as plan_run: bit-identical results. */
int  reach_lgg09_plan_stack_info(reach_lgg09_plan* p, void** stack_device, int64_t* rows_per_block, int64_t* ld, int32_t* S);
int  reach_lgg09_plan_stack_step(reach_lgg09_plan* p, int32_t step, int64_t row_lo, int64_t row_hi, int64_t* step_rows,
                                 int64_t* dest_row);
int  reach_lgg09_plan_stack_ready(reach_lgg09_plan* p);
int  reach_lgg09_plan_sync(reach_lgg09_plan* p);           /* waits; reports device faults    */
int  reach_lgg09_plan_fetch(reach_lgg09_plan* p, int64_t step0, int64_t nsteps, double* rho_host);
int  reach_lgg09_plan_nsteps_done(reach_lgg09_plan* p, int64_t* out);
/* rho(d_j, flowpipe) = max over steps of row j (AbstractFlowpipe.jl:35-37), reduced on device */
int  reach_lgg09_plan_max_over_steps(reach_lgg09_plan* p, double* max_host /* ndirs */,
                                     int64_t* argmax_host /* ndirs, may be NULL */);
void* reach_lgg09_plan_device_rho(reach_lgg09_plan* p);    /* device pointer, ndirs x nsteps   */
/* CUDA-event time of the last run (ms), the number of kernels it launched, and the launch
 * configuration chosen; any out pointer may be NULL */
int  reach_lgg09_plan_stats(reach_lgg09_plan* p, double* ms, int64_t* launches,
                            int32_t* time_block, int32_t* tile_m, int32_t* tile_n,
                            int32_t* nchains, double* gemm_ms, int64_t* gemm_launches);
/* Which main-loop kernel the plan uses / used: *path = 1 FP64 tensor-core (DMMA) GEMM, 2 persistent chain kernel,
 * 3 INT8 tcgen05 GEMM with FP64 results: *digit_planes = 10..16 moduli of the residue path (csrc/crt_gemm.cuh) or
 * 4..8 digit planes of the error-free slicing path (csrc/oz_gemm.cuh).
 * When path 3 was chosen automatically it is guarded: *guard_checks sampled passes were recomputed on the DMMA kernel,
 * *guard_max is the largest relative feature difference seen, and *fallback_pass >= 0 is the pass whose check failed
 * (-1: never): the run was then redone on the DMMA kernel from the last verified pass on (reach_lgg09_plan_guard).
 * Any pointer may be NULL. */
int  reach_lgg09_plan_path(reach_lgg09_plan* p, int32_t* path, int32_t* digit_planes, double* guard_max,
                           int64_t* guard_checks, int64_t* fallback_pass);
/* Invariant coverage of the uploaded plan: *faces half-spaces make up the invariant, *covered of them have a device
 * certificate, *exact = 1 when the device test is equivalent to the reference's `_isdisjoint(X, set(F[k]))`.
 * *steps_computed (after a run): time steps whose support values were produced before the run stopped launching
 * (== nsteps without an early exit; nsteps_done <= steps_computed).  Any pointer may be NULL. */
int  reach_lgg09_plan_invariant_info(reach_lgg09_plan* p, int32_t* faces, int32_t* covered, int32_t* exact,
                                     int64_t* steps_computed);
/* More on the guard of path 3.  Checks run at passes 0, 1, every 64th and the last.  A check compares (a) the pass's
 * features on the INT8 kernel and on the DMMA kernel from the same direction rows (tolerance 1e-12) and (b) the INT8
 * path's features with those of up to 128 sentinel chains that were propagated on the FP64 tensor pipe since pass 0
 * (*drift_max, tolerance 1e-11: the drift accumulated over all passes so far).  After every passed check the state
 * (direction rows, running input sums, carried features) is kept; a failed check restores the last kept state and
 * recomputes every pass since then on the DMMA kernel: *replay_from is the first recomputed pass (-1: none),
 * *replayed_passes the number of INT8 passes whose results were discarded.  Any pointer may be NULL. */
int  reach_lgg09_plan_guard(reach_lgg09_plan* p, double* drift_max, int64_t* replay_from, int64_t* replayed_passes);
void reach_lgg09_plan_destroy(reach_lgg09_plan* p);

/* LGG09 over a vector of initial sets -- the split caller `_solve_distributed` (src/Continuous/solve.jl:84-117),
 * which in the reference re-runs `post` per initial set.  Phi, the template and V are shared; Omega0s[i] is the
 * flattened Omega0 of split i.  The direction chains and the running input sum are computed once, the per-split
 * part is one FP64 tensor-core contraction (csrc/lgg09_batch.inl).  Output: rho[split][step][dir], i.e. split i's
 * ndirs x nsteps column-major matrix (the reference's rho_l, LGG09/post.jl:37-49) at offset i*ndirs*nsteps.
 * No invariant (early termination is per split; use the plan API for that). */
typedef struct reach_lgg09_batch reach_lgg09_batch;
int  reach_lgg09_batch_create(reach_ctx* ctx, int n, int ndirs, int64_t nsteps, int nsplits,
                              const double* Phi, const double* dirs,
                              const reach_setexpr* Omega0s /* [nsplits] */, const reach_setexpr* V /* NULL: homogeneous */,
                              reach_lgg09_batch** out);
int  reach_lgg09_batch_run(reach_lgg09_batch* b);          /* asynchronous on the ctx stream */
int  reach_lgg09_batch_sync(reach_lgg09_batch* b);
int  reach_lgg09_batch_fetch(reach_lgg09_batch* b, int split, int64_t step0, int64_t nsteps, double* rho_host);
/* split >= 0: rho(d_j, flowpipe of that split); split = -1: rho(d_j, MixedFlowpipe) = max over all splits and steps
 * (Flowpipes/MixedFlowpipe.jl:66-68), reduced on the device */
int  reach_lgg09_batch_max_over_steps(reach_lgg09_batch* b, int split, double* max_host /* ndirs */);
int  reach_lgg09_batch_stats(reach_lgg09_batch* b, double* ms, int64_t* launches, int32_t* time_block,
                             int32_t* alternatives, int32_t* feature_cols, int64_t* chunk_steps);
void* reach_lgg09_batch_device_rho(reach_lgg09_batch* b);  /* device pointer, [nsplits][nsteps][ndirs] */
void reach_lgg09_batch_destroy(reach_lgg09_batch* b);

/* BOX (src/Algorithms/BOX/reach_homog.jl:8-47,50-69, reach_inhomog.jl:8-60): hyperrectangular flowpipe
 *   non-recursive: c_k = Phi^(k-1) c_1 + W_k.c,  r_k = |Phi^(k-1)| r_1 + W_k.r,  W_2 = U, W_{k+1} = W_k (+) box(Phi^(k-1) U)
 *   recursive    : H_k = box(Phi H_{k-1})          (homogeneous only, as in the reference)
 * on the LGG09 device path (row i of Phi^k is the direction chain of e_i).  c0, r0: centre and radius of Omega0
 * (already a hyperrectangle: BOX/post.jl:29); cU, rU: the input hyperrectangle or both NULL (homogeneous).
 * c_host, r_host: n x nsteps column-major (column k = step k+1 of the reference), either may be NULL.  Invariants
 * (`_isdisjoint(X, H_k) && break`) are checked by the caller on the returned boxes. */
int reach_box(reach_ctx* ctx, int n, int64_t nsteps, const double* Phi, const double* c0, const double* r0,
              const double* cU, const double* rU, int recursive, const reach_lgg09_opts* opts,
              double* c_host, double* r_host);

/* ---------------------------------------------------------------------------------------------
 * GLGM06: zonotope flowpipes for a batch of initial sets ("splits") that share Phi and U.
 *
 *   homogeneous   (pU < 0):  Z_k = Phi Z_{k-1}                       (reach_homog.jl:33-73)
 *   inhomogeneous (pU >= 0): R_k = reduce(P_{k-1} Omega0 (+) W_k),  W_{k+1} = reduce(P_{k-1} U (+) W_k),
 *                            P_k = P_{k-1} Phi                       (reach_inhomog.jl:6-43)
 *
 * with reduce = reduce_order(., max_order, GIR05()).  Step 1 is Omega0 itself, after the initial
 * reduce_order of post.jl:26 (done on the device as well).
 *
 * Device-resident storage: per (split, step) the reduced zonotope [kept generators in sorted
 * order | n x n diagonal] is stored as the centre (n), the split's own mapped generators
 * P_{k-1} G0 (n x p0), the shared input generators W_k (once per step, not per split), the
 * selection (indices into [own | shared], sorted order) and the diagonal (n).  fetch materialises
 * the dense n x p matrix of one (split, step) exactly as the reference would hold it.
 * ------------------------------------------------------------------------------------------- */
typedef struct reach_glgm06_plan reach_glgm06_plan;

int  reach_glgm06_plan_create(reach_ctx* ctx, int n, int64_t nsteps, int nsplits, int max_order,
                              int p0, int pU /* <0: homogeneous */, reach_glgm06_plan** out);
int  reach_glgm06_plan_upload(reach_glgm06_plan* p, const double* Phi,
                              const double* c0 /* n x nsplits */,
                              const double* G0 /* n x p0 x nsplits */,
                              const double* cU /* n, or NULL */, const double* GU /* n x pU, or NULL */);
int  reach_glgm06_plan_run(reach_glgm06_plan* p);
int  reach_glgm06_plan_sync(reach_glgm06_plan* p);
/* number of generators of the (split, step) zonotope; step is 0-based (step 0 = Omega0) */
int  reach_glgm06_plan_ngens(reach_glgm06_plan* p, int split, int64_t step, int* out);
/* materialise one zonotope: c (n), G (n x ngens column-major).  sel (may be NULL, capacity
 * ngens) receives for each output column the index of the source generator in [own | shared]
 * numbering of the un-reduced zonotope, or -1 for the diagonal columns. */
int  reach_glgm06_plan_fetch(reach_glgm06_plan* p, int split, int64_t step,
                             double* c_host, double* G_host, int32_t* sel);
/* rho(d, Z_{split,step}) for all splits and steps on the device: out is nsplits x nsteps
 * column-major (split fastest) */
int  reach_glgm06_plan_support(reach_glgm06_plan* p, const double* d /* n */, double* out_host);
int  reach_glgm06_plan_stats(reach_glgm06_plan* p, double* ms, int64_t* launches,
                             int64_t* bytes_stored);
void reach_glgm06_plan_destroy(reach_glgm06_plan* p);

/* One-shot form (single problem or batch), host buffers out.  G_host is laid out with a fixed
 * stride of gcap columns per zonotope: zonotope (split s, step k) starts at
 * G_host + ((k*nsplits + s) * gcap) * n, and ngens_host[k*nsplits + s] columns of it are valid.
 * gcap >= max(p0_reduced, max_order*n).  Any output pointer may be NULL. */
int reach_glgm06(reach_ctx* ctx, int n, int64_t nsteps, int nsplits, int max_order,
                 const double* Phi, const double* c0, const double* G0, int p0,
                 const double* cU, const double* GU, int pU,
                 double* c_host, double* G_host, int gcap, int32_t* ngens_host);

/* reduce_order(Zonotope(c, G), r, GIR05()) on the device for one zonotope (post.jl:26 and tests):
 * G n x p in, Gred n x pred out (capacity p columns), *pred the new count, sel as above. */
int reach_reduce_order_gir05(reach_ctx* ctx, int n, int p, int max_order, const double* G,
                             double* Gred, int* pred, int32_t* sel);

/* ---------------------------------------------------------------------------------------------
 * Several GPUs behind one handle (single process, one host thread per device inside the library).
 *
 * The path has no inter-step exchange.  LGG09: the template directions are sharded over the devices
 * (`Threads.@threads for j in eachindex(l)`, LGG09/reach_inhomog.jl:35-40, becomes "device r owns a contiguous block of
 * direction chains"; l and -l stay on one device so they keep sharing (Phi')^k l); the one exchange is the all-gather of
 * the per-device support-value blocks, done by the copy engines as strided 2-D copies that place each block's rows at
 * their position in the caller's ndirs x nsteps matrix -- into rho_host and/or into one full device-resident copy per
 * GPU (rho_dev[j], written through peer memory over NVLink).  GLGM06: the initial-set splits are sharded
 * (`_solve_distributed`, Continuous/solve.jl:84-117); the shared chain is recomputed per device; no exchange.
 * Results are bit-identical to the single-device calls.  No invariant argument: early termination is decided on
 * all directions of a step, which no single device holds.
 * ------------------------------------------------------------------------------------------- */
typedef struct reach_multi reach_multi;
int         reach_multi_create(reach_multi** out, const int* device_ids /* NULL: 0..ndev-1 */, int ndev);
void        reach_multi_destroy(reach_multi* m);
int         reach_multi_ndev(const reach_multi* m);
reach_ctx*  reach_multi_ctx(reach_multi* m, int i);               /* the i-th device's context (owned by m) */
const char* reach_multi_last_error(const reach_multi* m);         /* m may be NULL: last create error        */
/* rho_host: ndirs x nsteps column-major, rows in the caller's order, or NULL.  rho_dev: NULL, or ndev device pointers
 * (entry j on device j, ndirs x nsteps doubles, NULL entries skipped) that each receive the complete matrix.
 * ms_per_device (may be NULL, ndev entries): CUDA-event time of each device's run.  On the GEMM paths the devices build the
 * row stack together (row ranges of every doubling step, peer copies) when their plans agree on the time block, else each
 * builds its own; the results are bit-identical either way (REACH_MULTI_NO_SHARED_STACK=1 forces the per-device build). */
int reach_multi_lgg09(reach_multi* m, int n, int ndirs, int64_t nsteps, const double* Phi, const double* dirs,
                      const reach_setexpr* Omega0, const reach_setexpr* V, const reach_lgg09_opts* opts,
                      double* rho_host, void* const* rho_dev, double* ms_per_device);
/* Same arguments and output layout as reach_glgm06; device r runs a contiguous, balanced block of the splits. */
int reach_multi_glgm06(reach_multi* m, int n, int64_t nsteps, int nsplits, int max_order,
                       const double* Phi, const double* c0, const double* G0, int p0,
                       const double* cU, const double* GU, int pU,
                       double* c_host, double* G_host, int gcap, int32_t* ngens_host);

/* Device-side discretisation matrices (src/Discretization/Exponentiation.jl): Phi = exp(A*delta) (:168-171),
 * Phi1(A, delta) = sum_i delta^(i+1)/(i+1)! A^i (:263-291) and Phi2(A, delta) = sum_i delta^(i+2)/(i+2)! A^i
 * (:397-427), which the Forward and NoBloating models turn into (Phi, Omega0, V) (ForwardModule.jl:86-164,
 * NoBloatingModule.jl:72-105).  The reference exponentiates a 2n x 2n / 3n x 3n block matrix; here the same
 * scaling and squaring runs on the three n x n blocks with the FP64 tensor-core GEMM (csrc/expm.cu).
 * A and the outputs are n x n column-major host matrices; any output may be NULL; *squarings (may be NULL)
 * returns the scaling exponent s (delta/2^s is the Taylor step). */
int reach_discretize_phi(reach_ctx* ctx, int n, const double* A, double delta,
                         double* Phi, double* Phi1, double* Phi2, int32_t* squarings);

/* FP64 GEMM throughput probe with this library's own DMMA kernel on an m x n x k problem
 * (C[m x n] = X[m x k] * Y[n x k]'), returns achieved TFLOP/s; used by bench.py beside cuBLAS. */
int reach_dgemm_probe(reach_ctx* ctx, int m, int n, int k, int iters, double* tflops);

/* The same contraction on the INT8 tcgen05 tensor pipe through `slices` (4..8) error-free 8-bit digit planes
 * per operand (csrc/oz_gemm.cuh): C[m x n] (row-major) = X[m x k] * Y[n x k]' for row-major host matrices.
 * C_host may be NULL.  *tflops counts 2 m n k per launch (FP64-equivalent), *slice_ms the digit-plane
 * conversion of both operands.  Used by the tests (accuracy against NumPy) and by bench.py. */
int reach_ozgemm_probe(reach_ctx* ctx, int m, int n, int k, int slices, const double* X_host,
                       const double* Y_host, double* C_host, int iters, double* tflops, double* slice_ms);

/* The same probe through the residue path (csrc/crt_gemm.cuh): `moduli` (8..16) INT8 GEMMs per FP64 product, one per
 * pairwise coprime modulus <= 256, rebuilt with the Chinese remainder theorem; 2-CTA tcgen05 MMAs. */
int reach_crtgemm_probe(reach_ctx* ctx, int m, int n, int k, int moduli, const double* X_host,
                        const double* Y_host, double* C_host, int iters, double* tflops, double* slice_ms);

/* Host-only: the constants of the residue path for `moduli` (10..16) moduli -- the moduli, the split fractions
 * f_i = f1_i + f2_i of w_i / P (w_i the Chinese-remainder weights, f1_i a multiple of 2^-40), P rounded to FP64 and the
 * 2-norm budgets 2^alpha / 2^beta of the integer X / Y rows (2^(alpha+beta) <= P/4).  Arrays hold `moduli` entries; any
 * pointer may be NULL.  tests/test_crt_host.py pins them against exact integer arithmetic without a GPU. */
int reach_crt_constants(int moduli, int32_t* p, double* f1, double* f2, double* P, int32_t* alpha, int32_t* beta);

#ifdef __cplusplus
}
#endif
#endif /* REACH_H */
