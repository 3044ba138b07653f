// reach_multi.cu -- several GPUs behind ONE handle of the C ABI (include/reach.h, "reach_multi_*").
//
// The reference's caller is one Julia process (`Threads.@threads for j in eachindex(l)`, LGG09/reach_inhomog.jl:35-40;
// `_solve_distributed`, Continuous/solve.jl:84-117), so the multi-GPU form of the boundary is one process driving ndev
// devices, not one process per device (bench.py's torchrun layout keeps using one reach_ctx per rank).  The path has no
// inter-step exchange:
//   LGG09  : template directions are sharded (a direction and its negation stay on one device so they keep sharing the
//            chain (Phi')^k l); every device runs an ordinary plan on its directions; the only exchange is the
//            all-gather of the per-device support-value blocks.  It is done by the copy engines as strided 2-D
//            copies that place every block's rows directly at their position in the caller's ndirs x NSTEPS matrix
//            (host buffer, and/or one full device-resident copy per GPU written through peer memory over NVLink) --
//            an NCCL all-gather would need the same bytes plus a pack/permute kernel on either side.
//   GLGM06 : initial-set splits are sharded; the shared chain (P_k, W_k) is recomputed per device; no exchange.
// Each device is driven by its own host thread (the guarded INT8 path synchronises with the host between passes).
#include "reach_common.cuh"

#include <algorithm>
#include <cstring>
#include <map>
#include <thread>

using namespace reach;

struct reach_multi {
    std::vector<reach_ctx*> ctx;
    std::vector<int> dev;
    std::string err;
};

namespace {

int multi_error(reach_multi* m, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (m) m->err = buf; else g_create_error = buf;
    return code;
}

// Chains: l and -l (exact negation) share one; first-appearance order.  Returns the rows of every chain.
std::vector<std::vector<int>> chain_groups(const double* dirs, int n, int ndirs) {
    std::map<std::string, int> index;
    std::vector<std::vector<int>> groups;
    std::vector<double> key(n);
    for (int j = 0; j < ndirs; ++j) {
        const double* d = dirs + size_t(j) * n;
        double sign = 1.0;
        for (int i = 0; i < n; ++i)
            if (d[i] != 0.0) { sign = d[i] < 0 ? -1.0 : 1.0; break; }
        for (int i = 0; i < n; ++i) key[i] = sign * d[i] + 0.0;          // + 0.0: -0.0 and 0.0 compare equal as bytes
        std::string k(reinterpret_cast<const char*>(key.data()), sizeof(double) * n);
        auto it = index.find(k);
        if (it == index.end() || groups[it->second].size() >= 2) {
            index[k] = int(groups.size());
            groups.push_back({j});
        } else {
            groups[it->second].push_back(j);
        }
    }
    return groups;
}

struct RowRun { int global0, local0, len; };      // rows global0.. of the caller's matrix = local rows local0..

std::vector<RowRun> row_runs(const std::vector<int>& rows) {
    std::vector<RowRun> runs;
    for (int l = 0; l < int(rows.size()); ++l) {
        if (!runs.empty() && rows[l] == runs.back().global0 + runs.back().len) ++runs.back().len;
        else runs.push_back({rows[l], l, 1});
    }
    return runs;
}

}  // namespace

extern "C" int reach_multi_create(reach_multi** out, const int* device_ids, int ndev) {
    if (!out) return multi_error(nullptr, REACH_EINVAL, "out is NULL");
    *out = nullptr;
    if (ndev < 1 || ndev > 64) return multi_error(nullptr, REACH_EINVAL, "ndev must be in 1..64, got %d", ndev);
    reach_multi* m = new (std::nothrow) reach_multi;
    if (!m) return multi_error(nullptr, REACH_ENOMEM, "out of host memory");
    for (int i = 0; i < ndev; ++i) {
        const int d = device_ids ? device_ids[i] : i;
        reach_ctx* c = nullptr;
        const int rc = reach_create(&c, d);
        if (rc != REACH_OK) {
            for (reach_ctx* x : m->ctx) reach_destroy(x);
            delete m;
            return rc;                                       // message already in the thread-local create error
        }
        m->ctx.push_back(c);
        m->dev.push_back(d);
    }
    // peer access for the device-resident all-gather (NVLink / NVSwitch); harmless where it is not available
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < ndev; ++j) {
            if (m->dev[i] == m->dev[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[j]) == cudaSuccess && can) {
                cudaSetDevice(m->dev[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j], 0);
                if (e != cudaSuccess) cudaGetLastError();    // already enabled, or refused: copies fall back to staging
            }
        }
    *out = m;
    return REACH_OK;
}

extern "C" void reach_multi_destroy(reach_multi* m) {
    if (!m) return;
    for (reach_ctx* c : m->ctx) reach_destroy(c);
    delete m;
}

extern "C" int reach_multi_ndev(const reach_multi* m) { return m ? int(m->ctx.size()) : REACH_EINVAL; }

extern "C" reach_ctx* reach_multi_ctx(reach_multi* m, int i) {
    return (m && i >= 0 && i < int(m->ctx.size())) ? m->ctx[i] : nullptr;
}

extern "C" const char* reach_multi_last_error(const reach_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }

extern "C" int reach_multi_lgg09(reach_multi* m, int n, int ndirs, int64_t nsteps, const double* Phi, const double* dirs,
                                 const reach_setexpr* Omega0, const reach_setexpr* V, const reach_lgg09_opts* opts,
                                 double* rho_host, void* const* rho_dev, double* ms_per_device) {
    if (!m) return REACH_EINVAL;
    if (n < 1 || ndirs < 1 || nsteps < 1) return multi_error(m, REACH_EINVAL, "n, ndirs and NSTEPS must be >= 1");
    if (!Phi || !dirs || !Omega0) return multi_error(m, REACH_EINVAL, "Phi, dirs and Omega0 must not be NULL");
    const int ndev = int(m->ctx.size());
    const auto groups = chain_groups(dirs, n, ndirs);
    const int per = (int(groups.size()) + ndev - 1) / ndev;
    std::vector<std::vector<int>> rows(ndev);
    for (int r = 0; r < ndev; ++r)
        for (int g = r * per; g < std::min<int>((r + 1) * per, int(groups.size())); ++g)
            for (int j : groups[g]) rows[r].push_back(j);
    std::vector<int> rc(ndev, REACH_OK);
    std::vector<std::string> msg(ndev);
    std::vector<reach_lgg09_plan*> plans(ndev, nullptr);
    auto fail = [&](int r, int code) { rc[r] = code; msg[r] = reach_last_error(m->ctx[r]); };

    // phase 1: every device runs its directions
    auto run_device = [&](int r) {
        if (rows[r].empty()) return;
        reach_ctx* c = m->ctx[r];
        const int mloc = int(rows[r].size());
        std::vector<double> local(size_t(n) * mloc);
        for (int l = 0; l < mloc; ++l) std::memcpy(&local[size_t(l) * n], dirs + size_t(rows[r][l]) * n, sizeof(double) * n);
        int e = REACH_OK;
        if (!plans[r]) {                                     // (not created by the shared stack build below)
            e = reach_lgg09_plan_create(c, n, mloc, nsteps, opts, &plans[r]);
            if (e == REACH_OK) e = reach_lgg09_plan_upload(plans[r], Phi, local.data(), Omega0, V, nullptr);
        }
        if (e == REACH_OK) e = reach_lgg09_plan_run(plans[r]);
        if (e == REACH_OK) e = reach_lgg09_plan_sync(plans[r]);
        if (e == REACH_OK && ms_per_device)
            e = reach_lgg09_plan_stats(plans[r], &ms_per_device[r], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (e != REACH_OK) fail(r, e);
    };
    // phase 0 (GEMM paths, >= 2 devices with work): the row stack [I;E](Phi')^b is the same on every device, so the devices
    // build it together -- each computes a row range of every doubling step (reach_lgg09_plan_stack_step) and pushes it to
    // the others by peer copies -- instead of each repeating the whole build (the largest per-device fixed cost of a sharded
    // run).  Same products in the same order: bit-identical to the per-device build.  Any irregularity (a chain-kernel
    // plan, different time blocks) falls back to the per-device build inside plan_run.
    auto shared_stack = [&]() {
        int nwork = 0;
        for (int r = 0; r < ndev; ++r) nwork += rows[r].empty() ? 0 : 1;
        if (nwork < 2 || nwork != ndev || getenv("REACH_MULTI_NO_SHARED_STACK")) return;
        auto create_device = [&](int r) {
            reach_ctx* c = m->ctx[r];
            const int mloc = int(rows[r].size());
            std::vector<double> local(size_t(n) * mloc);
            for (int l = 0; l < mloc; ++l) std::memcpy(&local[size_t(l) * n], dirs + size_t(rows[r][l]) * n, sizeof(double) * n);
            int e = reach_lgg09_plan_create(c, n, mloc, nsteps, opts, &plans[r]);
            if (e == REACH_OK) e = reach_lgg09_plan_upload(plans[r], Phi, local.data(), Omega0, V, nullptr);
            if (e != REACH_OK) fail(r, e);
        };
        {
            std::vector<std::thread> th;
            for (int r = 0; r < ndev; ++r) th.emplace_back(create_device, r);
            for (auto& t : th) t.join();
        }
        if (!std::all_of(rc.begin(), rc.end(), [](int v) { return v == REACH_OK; })) return;
        std::vector<double*> stack(ndev, nullptr);
        int64_t nbe0 = 0, ld0 = 0;
        int32_t S0 = 0;
        for (int r = 0; r < ndev; ++r) {
            void* sp = nullptr;
            int64_t nbe = 0, ld = 0;
            int32_t S = 0;
            if (reach_lgg09_plan_stack_info(plans[r], &sp, &nbe, &ld, &S) != REACH_OK) return;      // chain kernel: nothing to share
            if (r == 0) { nbe0 = nbe; ld0 = ld; S0 = S; }
            else if (nbe != nbe0 || ld != ld0 || S != S0) return;                                    // different time blocks
            stack[r] = static_cast<double*>(sp);
        }
        for (int step = 0;; ++step) {
            int64_t srows = 0, dest = 0;
            if (reach_lgg09_plan_stack_step(plans[0], step, -1, -1, &srows, &dest) != REACH_OK) return;
            if (srows == 0) break;
            int64_t chunk = (srows + ndev - 1) / ndev;
            chunk = (chunk + 7) / 8 * 8;
            for (int r = 0; r < ndev; ++r) {
                const int64_t lo = std::min(srows, r * chunk), hi = std::min(srows, (r + 1) * chunk);
                const int e = reach_lgg09_plan_stack_step(plans[r], step, lo, hi, nullptr, nullptr);
                if (e != REACH_OK) { fail(r, e); return; }
            }
            for (int r = 0; r < ndev; ++r) reach_synchronize(m->ctx[r]);
            for (int r = 0; r < ndev; ++r) {                 // device r pushes its rows into every other device's stack
                const int64_t lo = std::min(srows, r * chunk), hi = std::min(srows, (r + 1) * chunk);
                if (hi <= lo) continue;
                cudaSetDevice(m->dev[r]);
                cudaStream_t st = static_cast<cudaStream_t>(reach_stream(m->ctx[r]));
                const size_t off = size_t(dest + lo) * size_t(ld0), bytes = size_t(hi - lo) * size_t(ld0) * sizeof(double);
                for (int j = 0; j < ndev; ++j)
                    if (j != r && cudaMemcpyPeerAsync(stack[j] + off, m->dev[j], stack[r] + off, m->dev[r], bytes, st) != cudaSuccess) {
                        rc[r] = REACH_ECUDA;
                        msg[r] = "peer copy of the row stack failed";
                        return;
                    }
            }
            for (int r = 0; r < ndev; ++r) reach_synchronize(m->ctx[r]);
        }
        for (int r = 0; r < ndev; ++r)
            if (reach_lgg09_plan_stack_ready(plans[r]) != REACH_OK) return;     // (plan_run then builds the stack itself)
    };
    shared_stack();
    bool ok = std::all_of(rc.begin(), rc.end(), [](int v) { return v == REACH_OK; });
    if (ok) {
        std::vector<std::thread> th;
        for (int r = 0; r < ndev; ++r) th.emplace_back(run_device, r);
        for (auto& t : th) t.join();
    }
    ok = std::all_of(rc.begin(), rc.end(), [](int v) { return v == REACH_OK; });

    // phase 2: the all-gather.  Device r pushes its rows into the host matrix and into every device's full copy.
    auto gather_device = [&](int r) {
        if (rows[r].empty() || !plans[r]) return;
        reach_ctx* c = m->ctx[r];
        if (cudaSetDevice(m->dev[r]) != cudaSuccess) { rc[r] = REACH_ECUDA; msg[r] = "cudaSetDevice failed"; return; }
        cudaStream_t st = static_cast<cudaStream_t>(reach_stream(c));
        const int mloc = int(rows[r].size());
        const double* src = static_cast<const double*>(reach_lgg09_plan_device_rho(plans[r]));
        cudaError_t e = cudaSuccess;
        for (const RowRun& run : row_runs(rows[r])) {
            const double* s = src + run.local0;
            if (rho_host && e == cudaSuccess)
                e = cudaMemcpy2DAsync(rho_host + run.global0, size_t(ndirs) * 8, s, size_t(mloc) * 8, size_t(run.len) * 8,
                                      size_t(nsteps), cudaMemcpyDeviceToHost, st);
            if (rho_dev)
                for (int j = 0; j < ndev && e == cudaSuccess; ++j)
                    if (rho_dev[j])
                        e = cudaMemcpy2DAsync(static_cast<double*>(rho_dev[j]) + run.global0, size_t(ndirs) * 8, s,
                                              size_t(mloc) * 8, size_t(run.len) * 8, size_t(nsteps), cudaMemcpyDefault, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc[r] = REACH_ECUDA; msg[r] = std::string("all-gather copy failed: ") + cudaGetErrorString(e); }
    };
    if (ok && (rho_host || rho_dev)) {
        std::vector<std::thread> th;
        for (int r = 0; r < ndev; ++r) th.emplace_back(gather_device, r);
        for (auto& t : th) t.join();
    }
    for (int r = 0; r < ndev; ++r)
        if (plans[r]) {
            cudaSetDevice(m->dev[r]);
            reach_lgg09_plan_destroy(plans[r]);
        }
    for (int r = 0; r < ndev; ++r)
        if (rc[r] != REACH_OK) return multi_error(m, rc[r], "device %d: %s", m->dev[r], msg[r].c_str());
    return REACH_OK;
}

extern "C" int reach_multi_glgm06(reach_multi* m, int n, int64_t nsteps, int nsplits, int max_order, const double* Phi,
                                  const double* c0, const double* G0, int p0, const double* cU, const double* GU, int pU,
                                  double* c_host, double* G_host, int gcap, int32_t* ngens_host) {
    if (!m) return REACH_EINVAL;
    if (n < 1 || nsteps < 1 || nsplits < 1) return multi_error(m, REACH_EINVAL, "n, NSTEPS and the number of splits must be >= 1");
    const int ndev = int(m->ctx.size());
    std::vector<int> rc(ndev, REACH_OK);
    std::vector<std::string> msg(ndev);
    auto run_device = [&](int r) {
        const int base = nsplits / ndev, extra = nsplits % ndev;
        const int lo = r * base + std::min(r, extra), cnt = base + (r < extra ? 1 : 0);
        if (cnt == 0) return;
        reach_ctx* c = m->ctx[r];
        const bool whole = ndev == 1;
        std::vector<double> cl, Gl;
        std::vector<int32_t> nl;
        if (!whole) {
            if (c_host) cl.resize(size_t(n) * cnt * nsteps);
            if (G_host) Gl.resize(size_t(n) * gcap * cnt * nsteps);
            if (ngens_host) nl.resize(size_t(cnt) * nsteps);
        }
        const int e = reach_glgm06(c, n, nsteps, cnt, max_order, Phi, c0 + size_t(lo) * n, G0 + size_t(lo) * n * p0, p0, cU, GU, pU,
                                   whole ? c_host : (c_host ? cl.data() : nullptr), whole ? G_host : (G_host ? Gl.data() : nullptr),
                                   gcap, whole ? ngens_host : (ngens_host ? nl.data() : nullptr));
        if (e != REACH_OK) { rc[r] = e; msg[r] = reach_last_error(c); return; }
        if (whole) return;
        for (int64_t k = 0; k < nsteps; ++k) {          // zonotope (split s, step k) sits at index k*nsplits + s
            if (c_host) std::memcpy(c_host + (size_t(k) * nsplits + lo) * n, &cl[size_t(k) * cnt * n], sizeof(double) * n * cnt);
            if (G_host) std::memcpy(G_host + (size_t(k) * nsplits + lo) * gcap * n, &Gl[size_t(k) * cnt * gcap * n],
                                    sizeof(double) * n * gcap * cnt);
            if (ngens_host) std::memcpy(ngens_host + size_t(k) * nsplits + lo, &nl[size_t(k) * cnt], sizeof(int32_t) * cnt);
        }
    };
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r) th.emplace_back(run_device, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < ndev; ++r)
        if (rc[r] != REACH_OK) return multi_error(m, rc[r], "device %d: %s", m->dev[r], msg[r].c_str());
    return REACH_OK;
}
