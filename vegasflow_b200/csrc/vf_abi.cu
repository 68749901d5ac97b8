// extern "C" surface declared in include/vegasflow_b200.h.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include <dlfcn.h>

#include "vf_aux.cuh"
#include "vf_event.cuh"

namespace vf {

static thread_local char g_error[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return VF_ERR_CUDA;
}
void count_launch(int n) { g_launches += n; }

// Event-kernel timing hook: when enabled every event-kernel launch is bracketed by a pair of
// CUDA events on the launching stream; vf_kernel_time_ms sums the elapsed times.
struct TimingState {
    bool on = false;
    std::vector<cudaEvent_t> ev[2];  // start/stop pairs; category 0 = event kernels, 1 = epilogues
    size_t used[2] = {0, 0};
};
static thread_local TimingState g_timing;
void timing_begin(cudaStream_t stream, int category) {
    if (!g_timing.on) return;
    auto& ev = g_timing.ev[category];
    size_t& used = g_timing.used[category];
    if (used + 2 > ev.size()) {
        for (int k = 0; k < 2; ++k) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            ev.push_back(e);
        }
    }
    cudaEventRecord(ev[used], stream);
}
void timing_end(cudaStream_t stream, int category) {
    if (!g_timing.on || g_timing.used[category] + 2 > g_timing.ev[category].size()) return;
    cudaEventRecord(g_timing.ev[category][g_timing.used[category] + 1], stream);
    g_timing.used[category] += 2;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static int make_limits(int n_dim, const double* xmin, const double* xdelta, Limits* lim) {
    memset(lim, 0, sizeof(*lim));
    if ((xmin == nullptr) != (xdelta == nullptr)) {
        set_error("xmin and xdelta must both be given or both be NULL");
        return VF_ERR_INVALID;
    }
    if (!xmin) return VF_OK;
    if (n_dim > kMaxDim) {
        set_error("integration limits support n_dim <= %d", kMaxDim);
        return VF_ERR_UNSUPPORTED;
    }
    lim->has = 1;
    double jac = xdelta[0];  // tf.reduce_prod(xdelta), monte_carlo.py:171, left to right
    for (int j = 1; j < n_dim; ++j) jac = jac * xdelta[j];
    lim->xdeltajac = jac;
    for (int j = 0; j < n_dim; ++j) {
        lim->xmin[j] = xmin[j];
        lim->xdelta[j] = xdelta[j];
    }
    return VF_OK;
}

// Host constants of the built-in integrands.
static void fill_consts(int integrand, int n_dim, IntegrandConsts* ic) {
    memset(ic, 0, sizeof(*ic));
    if (integrand == VF_INTEGRAND_SYMGAUSS) {
        // examples/simgauss_tf.py:26-31
        const double a = 0.1;
        const double n100 = 100.0 * n_dim;
        ic->p[0] = std::pow(1.0 / a / std::sqrt(M_PI), (double)n_dim);
        ic->p[1] = (n100 + 1) * n100 / 2.0;
    }
}

static size_t workspace_need(int n_dim) {
    return ws_doubles(n_dim) * sizeof(double);
}

// `mode` words carry the sampling mode in the low byte and stream options above it.
static inline int split_mode(int& mode) {
    const int rng_bits = (mode & VF_MODE_RNG32) ? 32 : 52;
    mode &= 0xff;
    return rng_bits;
}
static inline int check_rng_bits(int rng_bits) {
    if (rng_bits == 52 || rng_bits == 32) return VF_OK;
    set_error("rng_bits must be 52 or 32 (got %d)", rng_bits);
    return VF_ERR_INVALID;
}

// Device alias of the caller's page-locked result ring: the tail kernel of every iteration
// stores (res, sigma) straight into it (posted writes over PCIe/C2C) -- the per-iteration
// device->host read-back of the reference's logging (monte_carlo.py:699-710) without a memcpy
// node between the kernels and without a host synchronisation per iteration.
static int map_host_results(double* host_results, double** dev_alias) {
    *dev_alias = nullptr;
    if (!host_results) return VF_OK;
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, host_results, 0) != cudaSuccess || !d) {
        (void)cudaGetLastError();
        set_error("host_results must be page-locked, device-mapped host memory "
                  "(cudaHostAlloc / torch pin_memory)");
        return VF_ERR_INVALID;
    }
    *dev_alias = (double*)d;
    return VF_OK;
}

static int check_common(int n_dim, int64_t n) {
    if (n_dim < 1) {
        set_error("n_dim must be >= 1 (got %d)", n_dim);
        return VF_ERR_INVALID;
    }
    if (n < 0) {
        set_error("negative event count");
        return VF_ERR_INVALID;
    }
    return VF_OK;
}

// ---- user integrands: separately compiled modules (vf_register_user_integrand) -------------
struct UserIntegrandEntry {
    void* handle;
    int (*event)(const EventLaunch*);
    int (*digest)(const DigestLaunch*);
    int (*plus)(const PlusLaunch*);
    int (*supported)(int);
};
static std::vector<UserIntegrandEntry> g_user;
static std::mutex g_user_mutex;

// Returns a COPY (the table may grow under another thread's registration).
static bool user_entry(int integrand, UserIntegrandEntry* out) {
    std::lock_guard<std::mutex> lock(g_user_mutex);
    const int k = integrand - VF_INTEGRAND_USER_BASE;
    if (k < 0 || k >= (int)g_user.size()) return false;
    *out = g_user[k];
    return true;
}

static int unknown_integrand(int integrand) {
    set_error("unknown integrand id %d", integrand);
    return VF_ERR_INVALID;
}

static int do_launch_event(int integrand, const EventLaunch& L) {
    switch (integrand) {
        case VF_INTEGRAND_SYMGAUSS: return launch_event<SymGauss>(L);
        case VF_INTEGRAND_PRODUCT: return launch_event<Product>(L);
        case VF_INTEGRAND_DRELLYAN_LO: return launch_event<DrellYanLO>(L);
        case VF_INTEGRAND_SINGLETOP_LO: return launch_event<SingleTopLO>(L);
        default: break;
    }
    UserIntegrandEntry u;
    if (user_entry(integrand, &u)) return u.event(&L);
    return unknown_integrand(integrand);
}
static int do_launch_digest(int integrand, const DigestLaunch& L) {
    switch (integrand) {
        case VF_INTEGRAND_SYMGAUSS: return launch_digest<SymGauss>(L);
        case VF_INTEGRAND_PRODUCT: return launch_digest<Product>(L);
        case VF_INTEGRAND_DRELLYAN_LO: return launch_digest<DrellYanLO>(L);
        case VF_INTEGRAND_SINGLETOP_LO: return launch_digest<SingleTopLO>(L);
        default: break;
    }
    UserIntegrandEntry u;
    if (user_entry(integrand, &u)) return u.digest(&L);
    return unknown_integrand(integrand);
}
static int do_launch_plus(int integrand, const PlusLaunch& L) {
    switch (integrand) {
        case VF_INTEGRAND_SYMGAUSS: return launch_plus<SymGauss>(L);
        case VF_INTEGRAND_PRODUCT: return launch_plus<Product>(L);
        case VF_INTEGRAND_DRELLYAN_LO: return launch_plus<DrellYanLO>(L);
        case VF_INTEGRAND_SINGLETOP_LO: return launch_plus<SingleTopLO>(L);
        default: break;
    }
    UserIntegrandEntry u;
    if (user_entry(integrand, &u)) return u.plus(&L);
    return unknown_integrand(integrand);
}

}  // namespace vf

using namespace vf;

extern "C" {

int vf_version(void) { return VF_ABI_VERSION; }
const char* vf_last_error(void) { return g_error; }

int vf_integrand_id(const char* name) {
    if (!name) return VF_ERR_INVALID;
    if (!strcmp(name, "symgauss")) return VF_INTEGRAND_SYMGAUSS;
    if (!strcmp(name, "product")) return VF_INTEGRAND_PRODUCT;
    if (!strcmp(name, "drellyan_lo")) return VF_INTEGRAND_DRELLYAN_LO;
    if (!strcmp(name, "singletop_lo")) return VF_INTEGRAND_SINGLETOP_LO;
    set_error("unknown integrand '%s'", name);
    return VF_ERR_INVALID;
}

int vf_supported(int integrand, int n_dim) {
    switch (integrand) {
        case VF_INTEGRAND_SYMGAUSS: return supported_dim<SymGauss>(n_dim);
        case VF_INTEGRAND_PRODUCT: return supported_dim<Product>(n_dim);
        case VF_INTEGRAND_DRELLYAN_LO: return supported_dim<DrellYanLO>(n_dim);
        case VF_INTEGRAND_SINGLETOP_LO: return supported_dim<SingleTopLO>(n_dim);
        default: break;
    }
    UserIntegrandEntry u;
    if (user_entry(integrand, &u)) return u.supported(n_dim);
    return 0;
}

int vf_register_user_integrand(const char* module_path) {
    if (!module_path) {
        set_error("vf_register_user_integrand: null path");
        return VF_ERR_INVALID;
    }
    void* h = dlopen(module_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        set_error("cannot load integrand module %s: %s", module_path, dlerror());
        return VF_ERR_INVALID;
    }
    UserIntegrandEntry e;
    e.handle = h;
    e.event = (int (*)(const EventLaunch*))dlsym(h, "vfu_launch_event");
    e.digest = (int (*)(const DigestLaunch*))dlsym(h, "vfu_launch_digest");
    e.plus = (int (*)(const PlusLaunch*))dlsym(h, "vfu_launch_plus");
    e.supported = (int (*)(int))dlsym(h, "vfu_supported_dim");
    int (*abi)(void) = (int (*)(void))dlsym(h, "vfu_abi_version");
    if (!e.event || !e.digest || !e.plus || !e.supported || !abi) {
        set_error("%s does not export the vfu_* entry points", module_path);
        dlclose(h);
        return VF_ERR_INVALID;
    }
    if (abi() != VF_ABI_VERSION) {
        set_error("%s was built against ABI %d, library is %d", module_path, abi(), VF_ABI_VERSION);
        dlclose(h);
        return VF_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lock(g_user_mutex);
    g_user.push_back(e);
    return VF_INTEGRAND_USER_BASE + (int)g_user.size() - 1;
}

double vf_flops_per_event(int mode, int integrand, int n_dim, int plus) {
    // add/sub/mul/div = 1, each transcendental = 1.
    const double d = n_dim;
    double base = mode == VF_MODE_VEGAS ? 12.0 * d + 5.0 : 3.0 * d + 5.0;
    if (plus) base += d + 1.0;
    switch (integrand) {
        case VF_INTEGRAND_SYMGAUSS: return base + 4.0 * d + 4.0;
        case VF_INTEGRAND_PRODUCT: return base + (d - 1.0);
        // "count by instantiating the shared integrand header with an op-counting
        // scalar type on the host" -- vf_integrands.cuh compiled with a counting scalar
        // (tests/host_shim/count_flops_host.cpp; fma = 2, div / sqrt / transcendental = 1),
        // averaged over uniformly drawn events (single-top's count depends on the branch taken:
        // 244.7).  The reference's LITERAL chain -- zero-padded complex arithmetic, acos / sincos
        // round trips, staged quotients -- is 448 / 1354 operations (oracle/count_flops.py);
        // bench.py reports that figure beside this one.
        case VF_INTEGRAND_DRELLYAN_LO: return base + 128.0;
        case VF_INTEGRAND_SINGLETOP_LO: return base + 245.0;
        default: return base;
    }
}

size_t vf_workspace_bytes(int n_dim) { return n_dim < 1 ? 0 : workspace_need(n_dim); }

int vf_run_event(int mode, int integrand, int n_dim, uint64_t ev_begin, int64_t n_events,
                 double xjac, uint64_t seed, uint32_t iteration, int train,
                 const double* divisions, const double* xmin, const double* xdelta,
                 double* out_sums, double* out_hist, int accumulate, void* workspace,
                 size_t workspace_bytes, void* stream) {
    int rc = check_common(n_dim, n_events);
    if (rc) return rc;
    const int rng_bits = split_mode(mode);
    if (mode != VF_MODE_PLAIN && mode != VF_MODE_VEGAS) {
        set_error("unknown mode %d", mode);
        return VF_ERR_INVALID;
    }
    if (mode == VF_MODE_VEGAS && !divisions) {
        set_error("VEGAS mode needs a divisions grid");
        return VF_ERR_INVALID;
    }
    const bool with_hist = mode == VF_MODE_VEGAS && train;
    if (!out_sums || (with_hist && !out_hist) || !workspace) {
        set_error("null output/workspace pointer");
        return VF_ERR_INVALID;
    }
    if (workspace_bytes < workspace_need(n_dim)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, workspace_need(n_dim));
        return VF_ERR_WORKSPACE;
    }
    EventLaunch L;
    L.mode = mode;
    L.rng_bits = rng_bits;
    L.n_dim = n_dim;
    L.stream = (cudaStream_t)stream;
    int nblocks = 0;
    L.nblocks_out = &nblocks;
    rc = make_limits(n_dim, xmin, xdelta, &L.k.lim);
    if (rc) return rc;
    fill_consts(integrand, n_dim, &L.k.ic);
    L.k.divisions = divisions;
    L.k.partials = (double*)workspace;
    L.k.ev_begin = ev_begin;
    L.k.ev_end = ev_begin + (uint64_t)n_events;
    L.k.xjac = xjac;
    L.k.pk = make_philox_keys(seed);
    L.k.iteration = iteration;
    L.k.train = train;
    rc = do_launch_event(integrand, L);
    if (rc) return rc;
    return launch_finalize((double*)workspace, nblocks, n_dim, with_hist, out_sums, out_hist,
                           accumulate, L.stream);
}

int vf_run_iterations(int mode, int integrand, int n_dim, int64_t n_events, uint64_t seed,
                      uint32_t first_iteration, int n_iter, int train, double* divisions,
                      const double* xmin, const double* xdelta, double* packed, double* results,
                      double* host_results, void* workspace, size_t workspace_bytes,
                      void* stream) {
    int rc = check_common(n_dim, n_events);
    if (rc) return rc;
    const int rng_bits = split_mode(mode);
    if (mode != VF_MODE_PLAIN && mode != VF_MODE_VEGAS) {
        set_error("unknown mode %d", mode);
        return VF_ERR_INVALID;
    }
    if (n_iter < 0 || n_events < 2 || !packed || !results || !workspace ||
        (mode == VF_MODE_VEGAS && !divisions)) {
        set_error("vf_run_iterations: bad arguments");
        return VF_ERR_INVALID;
    }
    if (workspace_bytes < workspace_need(n_dim)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, workspace_need(n_dim));
        return VF_ERR_WORKSPACE;
    }
    const bool with_hist = mode == VF_MODE_VEGAS && train;
    EventLaunch L;
    L.mode = mode;
    L.rng_bits = rng_bits;
    L.n_dim = n_dim;
    L.stream = (cudaStream_t)stream;
    int nblocks = 0;
    L.nblocks_out = &nblocks;
    rc = make_limits(n_dim, xmin, xdelta, &L.k.lim);
    if (rc) return rc;
    fill_consts(integrand, n_dim, &L.k.ic);
    L.k.divisions = divisions;
    L.k.partials = (double*)workspace;
    L.k.ev_begin = 0;
    L.k.ev_end = (uint64_t)n_events;
    L.k.xjac = 1.0 / (double)n_events;  // monte_carlo.py:224-227
    L.k.pk = make_philox_keys(seed);
    L.k.train = train;
    double* out_hist = packed;
    double* out_sums = packed + (size_t)n_dim * kBins;
    double* host_alias = nullptr;
    rc = map_host_results(host_results, &host_alias);
    if (rc) return rc;
    for (int it = 0; it < n_iter; ++it) {
        L.k.iteration = first_iteration + (uint32_t)it;
        rc = do_launch_event(integrand, L);
        if (rc) return rc;
        rc = launch_finalize_epilogue((double*)workspace, nblocks, n_dim, with_hist, n_events,
                                      train, out_sums, out_hist, divisions, results + 2 * it,
                                      host_alias ? host_alias + 2 * it : nullptr, L.stream);
        if (rc) return rc;
    }
    return VF_OK;
}

size_t vf_exchange_bytes(int n_dim, int world, int64_t n_cubes) {
    return (n_dim < 1 || world < 1 || n_cubes < 0) ? 0 : exchange_bytes(n_dim, world, n_cubes);
}

int vf_run_iterations_sharded(int mode, int integrand, int n_dim, uint64_t ev_begin,
                              int64_t n_events_local, int64_t n_events_total, uint64_t seed,
                              uint32_t first_iteration, int n_iter, int train, double* divisions,
                              const double* xmin, const double* xdelta, double* packed,
                              double* results, double* host_results, void* workspace,
                              size_t workspace_bytes, int rank, int world,
                              const uint64_t* peer_buffers, uint64_t first_seq, void* stream) {
    double* result = results;
    const uint64_t seq = first_seq;
    int rc = check_common(n_dim, n_events_local);
    if (rc) return rc;
    const int rng_bits = split_mode(mode);
    if (mode != VF_MODE_PLAIN && mode != VF_MODE_VEGAS) {
        set_error("unknown mode %d", mode);
        return VF_ERR_INVALID;
    }
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || !peer_buffers || seq == 0 ||
        n_iter < 0 || n_events_total < 2 || !packed || !result || !workspace ||
        (mode == VF_MODE_VEGAS && !divisions)) {
        set_error("vf_run_iterations_sharded: bad arguments (world must be 1..%d)", kMaxWorld);
        return VF_ERR_INVALID;
    }
    if (workspace_bytes < workspace_need(n_dim)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, workspace_need(n_dim));
        return VF_ERR_WORKSPACE;
    }
    const bool with_hist = mode == VF_MODE_VEGAS && train;
    EventLaunch L;
    L.mode = mode;
    L.rng_bits = rng_bits;
    L.n_dim = n_dim;
    L.stream = (cudaStream_t)stream;
    int nblocks = 0;
    L.nblocks_out = &nblocks;
    rc = make_limits(n_dim, xmin, xdelta, &L.k.lim);
    if (rc) return rc;
    fill_consts(integrand, n_dim, &L.k.ic);
    L.k.divisions = divisions;
    L.k.partials = (double*)workspace;
    L.k.ev_begin = ev_begin;
    L.k.ev_end = ev_begin + (uint64_t)n_events_local;
    L.k.xjac = 1.0 / (double)n_events_total;  // monte_carlo.py:224-227
    L.k.pk = make_philox_keys(seed);
    L.k.train = train;
    PeerPtrs peers;
    for (int p = 0; p < kMaxWorld; ++p)
        peers.base[p] = p < world ? (unsigned long long*)(uintptr_t)peer_buffers[p] : nullptr;
    double* host_alias = nullptr;
    rc = map_host_results(host_results, &host_alias);
    if (rc) return rc;
    for (int it = 0; it < n_iter; ++it) {
        L.k.iteration = first_iteration + (uint32_t)it;
        rc = do_launch_event(integrand, L);
        if (rc) return rc;
        rc = launch_exchange_epilogue((double*)workspace, nblocks, n_dim, with_hist,
                                      n_events_total, train, packed + (size_t)n_dim * kBins, packed,
                                      divisions, result + 2 * it,
                                      host_alias ? host_alias + 2 * it : nullptr, rank, world,
                                      peers, seq + it, L.stream);
        if (rc) return rc;
    }
    return VF_OK;
}

int vf_refine_grid(int n_dim, const double* hist, double* divisions, void* stream) {
    if (n_dim < 1 || !hist || !divisions) {
        set_error("vf_refine_grid: bad arguments");
        return VF_ERR_INVALID;
    }
    return launch_refine(n_dim, hist, divisions, (cudaStream_t)stream);
}

int vf_iteration_epilogue(int n_dim, int64_t n_events, int train, const double* sums,
                          const double* hist, double* divisions, double* result, void* stream) {
    if (n_dim < 1 || !sums || !result || (train && (!hist || !divisions))) {
        set_error("vf_iteration_epilogue: bad arguments");
        return VF_ERR_INVALID;
    }
    return launch_epilogue(n_dim, n_events, train, sums, hist, divisions, result,
                           (cudaStream_t)stream);
}

int vf_digest_from_uniforms(int mode, int integrand, int n_dim, int64_t n, const double* rnds,
                            const double* divisions, double xjac, const double* xmin,
                            const double* xdelta, double* x, double* w, int32_t* ind, double* wf,
                            void* stream) {
    int rc = check_common(n_dim, n);
    if (rc) return rc;
    if (n == 0) return VF_OK;
    if (!rnds || (mode == VF_MODE_VEGAS && !divisions)) {
        set_error("vf_digest_from_uniforms: null input");
        return VF_ERR_INVALID;
    }
    DigestLaunch L;
    L.mode = mode;
    L.n_dim = n_dim;
    L.stream = (cudaStream_t)stream;
    rc = make_limits(n_dim, xmin, xdelta, &L.k.lim);
    if (rc) return rc;
    fill_consts(integrand, n_dim, &L.k.ic);
    L.k.rnds = rnds;
    L.k.divisions = divisions;
    L.k.x = x;
    L.k.w = w;
    L.k.ind = ind;
    L.k.wf = wf;
    L.k.n = n;
    L.k.xjac = xjac;
    return do_launch_digest(integrand, L);
}

int vf_uniforms(int n_dim, uint64_t ev_begin, int64_t n, uint64_t seed, uint32_t iteration,
                int rng_bits, double* rnds, void* stream) {
    int rc = check_common(n_dim, n);
    if (rc) return rc;
    rc = check_rng_bits(rng_bits);
    if (rc) return rc;
    return launch_uniforms(n_dim, ev_begin, n, seed, iteration, rng_bits, rnds,
                           (cudaStream_t)stream);
}

int vf_sample(int mode, int n_dim, uint64_t ev_begin, int64_t n, double xjac, uint64_t seed,
              uint32_t iteration, const double* divisions, const double* xmin,
              const double* xdelta, double* x, double* w, int32_t* ind, void* stream) {
    int rc = check_common(n_dim, n);
    if (rc) return rc;
    const int rng_bits = split_mode(mode);
    if (n_dim > kMaxDim) {
        set_error("vf_sample supports n_dim <= %d", kMaxDim);
        return VF_ERR_UNSUPPORTED;
    }
    if (!x || !w || (mode == VF_MODE_VEGAS && !divisions)) {
        set_error("vf_sample: null pointer");
        return VF_ERR_INVALID;
    }
    Limits lim;
    rc = make_limits(n_dim, xmin, xdelta, &lim);
    if (rc) return rc;
    return launch_sample(mode, n_dim, ev_begin, n, xjac, seed, iteration, rng_bits, divisions, lim,
                         x, w, ind, (cudaStream_t)stream);
}

int vf_accumulate(int n_dim, int64_t n, const double* w, const double* f, const int32_t* ind,
                  int train, double* out_sums, double* out_hist, int accumulate, void* workspace,
                  size_t workspace_bytes, void* stream) {
    int rc = check_common(n_dim, n);
    if (rc) return rc;
    if (n_dim > kMaxDim) {
        set_error("vf_accumulate supports n_dim <= %d", kMaxDim);
        return VF_ERR_UNSUPPORTED;
    }
    const bool with_hist = train && ind && out_hist;
    if (!w || !f || !out_sums || !workspace) {
        set_error("vf_accumulate: null pointer");
        return VF_ERR_INVALID;
    }
    if (workspace_bytes < workspace_need(n_dim)) {
        set_error("workspace too small");
        return VF_ERR_WORKSPACE;
    }
    int nblocks = 0;
    rc = launch_accumulate(n_dim, n, w, f, ind, with_hist, (double*)workspace, &nblocks,
                           (cudaStream_t)stream);
    if (rc) return rc;
    return launch_finalize((double*)workspace, nblocks, n_dim, with_hist, out_sums, out_hist,
                           accumulate, (cudaStream_t)stream);
}

int vfp_run_event(int integrand, int n_dim, int n_strat, int64_t n_cubes, int64_t n_events,
                  const int32_t* n_ev, const int64_t* ev_offset, double xjac, uint64_t seed,
                  uint32_t iteration, int rng_bits, int train, const double* divisions,
                  const double* xmin,
                  const double* xdelta, double* ress, double* ress2, double* out_hist,
                  int accumulate, void* workspace, size_t workspace_bytes, const double* rnds,
                  double* x, double* w, int32_t* ind, double* wf, void* stream) {
    int rc = check_common(n_dim, n_events);
    if (rc) return rc;
    rc = check_rng_bits(rng_bits);
    if (rc) return rc;
    if (!n_ev || !ev_offset || !divisions || !ress || !ress2 || !workspace ||
        (train && !out_hist) || n_strat < 1 || n_cubes < 1) {
        set_error("vfp_run_event: bad arguments");
        return VF_ERR_INVALID;
    }
    if (workspace_bytes < workspace_need(n_dim)) {
        set_error("workspace too small");
        return VF_ERR_WORKSPACE;
    }
    PlusLaunch L;
    L.n_dim = n_dim;
    L.rng_bits = rng_bits;
    L.stream = (cudaStream_t)stream;
    int nblocks = 0;
    L.nblocks_out = &nblocks;
    rc = make_limits(n_dim, xmin, xdelta, &L.k.lim);
    if (rc) return rc;
    fill_consts(integrand, n_dim, &L.k.ic);
    L.k.divisions = divisions;
    L.k.partials = (double*)workspace;
    L.k.n_ev = n_ev;
    L.k.ev_offset = ev_offset;
    L.k.ress = ress;
    L.k.ress2 = ress2;
    L.k.rnds = rnds;
    L.k.x = x;
    L.k.w = w;
    L.k.ind = ind;
    L.k.wf = wf;
    L.k.n_cubes = n_cubes;
    L.k.n_events = n_events;
    L.k.n_strat = n_strat;
    L.k.rank = 0;
    L.k.world = 1;
    L.k.xjac = xjac;
    L.k.pk = make_philox_keys(seed);
    L.k.iteration = iteration;
    L.k.train = train;
    rc = do_launch_plus(integrand, L);
    if (rc) return rc;
    if (!train) return VF_OK;
    // histogram only: the scalar block writes into a scratch pair past the partial records
    double* scratch_sums = (double*)workspace + ws_doubles(n_dim) - 2;
    return launch_finalize((double*)workspace, nblocks, n_dim, true, scratch_sums, out_hist,
                           accumulate, L.stream);
}

int vfp_iteration_epilogue(int64_t n_cubes, const double* ress, const double* ress2, int adaptive,
                           int min_neval_hcube, int64_t init_calls, int32_t* n_ev,
                           int64_t* ev_offset, double* arr_var, double* result,
                           int64_t* n_events_out, void* stream) {
    if (n_cubes < 1 || !ress || !ress2 || !n_ev || !arr_var || !result ||
        (adaptive && (!ev_offset || !n_events_out))) {
        set_error("vfp_iteration_epilogue: bad arguments");
        return VF_ERR_INVALID;
    }
    return launch_plus_epilogue(n_cubes, ress, ress2, adaptive, min_neval_hcube, init_calls, n_ev,
                                ev_offset, arr_var, result, n_events_out, (cudaStream_t)stream);
}

int vfp_run_iterations(int integrand, int n_dim, int n_strat, int64_t n_cubes, uint64_t seed,
                       uint32_t first_iteration, int n_iter, int rng_bits, int train, int adaptive,
                       int min_neval_hcube, int64_t init_calls, double* divisions,
                       const double* xmin, const double* xdelta, int32_t* n_ev,
                       int64_t* ev_offset, double* ress, double* ress2, double* arr_var,
                       double* out_hist, double* results, double* host_results, void* workspace,
                       size_t workspace_bytes, int rank, int world, const uint64_t* peer_buffers,
                       uint64_t first_seq, void* stream) {
    int rc = check_common(n_dim, 0);
    if (rc) return rc;
    rc = check_rng_bits(rng_bits);
    if (rc) return rc;
    if (n_iter < 0 || n_strat < 1 || n_cubes < 1 || !divisions || !n_ev || !ev_offset || !ress ||
        !ress2 || !arr_var || !results || !workspace || (train && !out_hist) || world < 1 ||
        world > kMaxWorld || rank < 0 || rank >= world ||
        (world > 1 && (!peer_buffers || first_seq == 0))) {
        set_error("vfp_run_iterations: bad arguments (world must be 1..%d)", kMaxWorld);
        return VF_ERR_INVALID;
    }
    if (workspace_bytes < workspace_need(n_dim)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, workspace_need(n_dim));
        return VF_ERR_WORKSPACE;
    }
    PlusLaunch L;
    L.n_dim = n_dim;
    L.rng_bits = rng_bits;
    L.stream = (cudaStream_t)stream;
    int nblocks = 0;
    L.nblocks_out = &nblocks;
    rc = make_limits(n_dim, xmin, xdelta, &L.k.lim);
    if (rc) return rc;
    fill_consts(integrand, n_dim, &L.k.ic);
    L.k.divisions = divisions;
    L.k.partials = (double*)workspace;
    L.k.n_ev = n_ev;
    L.k.ev_offset = ev_offset;
    L.k.ress = ress;
    L.k.ress2 = ress2;
    L.k.rnds = nullptr;
    L.k.x = L.k.w = L.k.wf = nullptr;
    L.k.ind = nullptr;
    L.k.n_cubes = n_cubes;
    L.k.n_events = -1;  // device-resident: ev_offset[n_cubes]
    L.k.n_strat = n_strat;
    L.k.rank = rank;
    L.k.world = world;
    L.k.xjac = 1.0 / (double)n_cubes;  // vflowplus.py:139
    L.k.pk = make_philox_keys(seed);
    L.k.train = train;
    PeerPtrs peers;
    for (int p = 0; p < kMaxWorld; ++p)
        peers.base[p] = (world > 1 && p < world) ? (unsigned long long*)(uintptr_t)peer_buffers[p]
                                                 : nullptr;
    double* host_alias = nullptr;
    rc = map_host_results(host_results, &host_alias);
    if (rc) return rc;
    for (int it = 0; it < n_iter; ++it) {
        L.k.iteration = first_iteration + (uint32_t)it;
        rc = do_launch_plus(integrand, L);
        if (rc) return rc;
        rc = launch_plus_iteration_tail((double*)workspace, nblocks, n_dim, train, out_hist,
                                        divisions, n_cubes, ress, ress2, adaptive, min_neval_hcube,
                                        init_calls, n_ev, ev_offset, arr_var, results + 3 * it,
                                        host_alias ? host_alias + 3 * it : nullptr, rank, world,
                                        peers, first_seq + it, L.stream);
        if (rc) return rc;
    }
    return VF_OK;
}

int vf_fp64_peak_probe(int iters, double* tflops) {
    if (iters < 1 || !tflops) {
        set_error("vf_fp64_peak_probe: bad arguments");
        return VF_ERR_INVALID;
    }
    return run_fp64_probe(iters, tflops);
}

int vf_sm_count(void) { return sm_count(); }

int vf_kernel_timing(int enable) {
    g_timing.on = enable != 0;
    g_timing.used[0] = g_timing.used[1] = 0;
    return VF_OK;
}

static int timing_sum(int category, double* total_ms, int* launches) {
    if (!total_ms || !launches) {
        set_error("vf_kernel_time_ms: null pointer");
        return VF_ERR_INVALID;
    }
    double tot = 0.0;
    const auto& ev = g_timing.ev[category];
    const size_t used = g_timing.used[category];
    for (size_t k = 0; k + 1 < used; k += 2) {
        VF_CUDA_CHECK(cudaEventSynchronize(ev[k + 1]));
        float ms = 0.f;
        VF_CUDA_CHECK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int)(used / 2);
    return VF_OK;
}

int vf_kernel_time_ms(double* total_ms, int* launches) { return timing_sum(0, total_ms, launches); }
int vf_epilogue_time_ms(double* total_ms, int* launches) { return timing_sum(1, total_ms, launches); }

int64_t vf_launch_count(int reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

}  // extern "C"
