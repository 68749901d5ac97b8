/*
 * vegasflow_b200 -- C ABI of the B200-native VEGAS hot path.
 *
 * Drop-in boundary for the per-iteration event loop of N3PDF/vegasflow v1.4.0.
 * The reference has no native plugin API; its seam is the Python abstract
 * method contract of MonteCarloFlow (src/vegasflow/monte_carlo.py:278-310):
 * `_run_event(integrand, ncalls) -> (res, res2, arr_res2)` called once per
 * chunk by `run_event` (monte_carlo.py:420-480) and `refine_grid(arr_res2)`
 * (src/vegasflow/vflow.py:349-362).  Each entry point below names the
 * reference lines it replaces.  Citations are relative to /root/reference.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the message is in
 *     vf_last_error() (thread-local).  No exceptions cross the ABI.
 *   - pointers marked [dev] are device pointers owned by the caller (torch
 *     allocations in the Python host layer); [host] are host pointers read
 *     synchronously during the call.  The library allocates no persistent
 *     device memory; scratch comes from the caller (vf_workspace_bytes).
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered
 *     and asynchronous.  Not re-entrant per workspace.
 *   - arithmetic is IEEE fp64 with the reference's operation order (no FMA
 *     contraction on the map/integrand path); bins BINS_MAX = 50
 *     (src/vegasflow/configflow.py:13).
 *   - grid layout: divisions[n_dim][51] row-major (src/vegasflow/vflow.py:239-242);
 *     histogram arr_res2[n_dim][50] row-major (vflow.py:387).
 *   - random stream: Philox4x32-10, key = seed, counter =
 *     (event_lo, event_hi, block, iteration); block b feeds dimensions 2b, 2b+1 with a 52-bit
 *     mantissa fill each (default) or 4b..4b+3 with 32 bits each (VF_MODE_RNG32 / rng_bits 32);
 *     r = TECH_CUT + u*(1-2*TECH_CUT) (monte_carlo.py:264-266 semantics).
 */
#ifndef VEGASFLOW_B200_H
#define VEGASFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VF_BINS_MAX 50
#define VF_ABI_VERSION 2

/* sampling modes */
#define VF_MODE_PLAIN 0 /* PlainFlow, src/vegasflow/plain.py:18-35 */
#define VF_MODE_VEGAS 1 /* VegasFlow, src/vegasflow/vflow.py:389-430 */
/* OR'ed into a `mode` argument: draw FOUR 32-bit-resolution uniforms per Philox block instead
 * of two 52-bit ones (half the integer-multiply work; resolution 2.3e-10 << TECH_CUT = 1e-8).
 * Off by default: the default stream fills all 52 mantissa bits like tf.random.uniform. */
#define VF_MODE_RNG32 0x100

/* built-in integrand ids (vf_integrand_id) */
#define VF_INTEGRAND_SYMGAUSS 0     /* examples/simgauss_tf.py:22-32 */
#define VF_INTEGRAND_PRODUCT 1      /* README.md:63-68 */
#define VF_INTEGRAND_DRELLYAN_LO 2  /* examples/drellyan_lo_tf.py:27-249, n_dim 4 */
#define VF_INTEGRAND_SINGLETOP_LO 3 /* examples/singletop_lo_tf.py:45-270, n_dim 3 */
#define VF_INTEGRAND_USER_BASE 16   /* ids returned by vf_register_user_integrand */

/* error codes */
#define VF_OK 0
#define VF_ERR_INVALID (-1)     /* bad argument */
#define VF_ERR_UNSUPPORTED (-2) /* (integrand, n_dim) combination not instantiated */
#define VF_ERR_CUDA (-3)        /* CUDA runtime error, see vf_last_error() */
#define VF_ERR_WORKSPACE (-4)   /* workspace too small */

int vf_version(void);
const char* vf_last_error(void);

/* Name -> id: "symgauss", "product", "drellyan_lo", "singletop_lo"; <0 if unknown. */
int vf_integrand_id(const char* name);
/*
 * Register a user integrand compiled into its own module: a shared library built from
 * vegasflow_b200/csrc/vf_user_integrand.cu.in around the user's
 *     __device__ double integrand(const double* x, int n_dim)
 * (see vegasflow_b200.integrands.cuda_integrand, which drives nvcc).  The fused kernels are
 * instantiated for that function inside the module, so it runs inline in the event kernel
 * exactly like the built-ins.  Native counterpart of the reference's "integrand in C / CUDA"
 * examples (examples/simgauss_cffi.py:25-62, examples/cuda/integrand.cpp:41-89), which also
 * compile user code at run time.  Returns the new integrand id (>= VF_INTEGRAND_USER_BASE).
 */
int vf_register_user_integrand(const char* module_path);
/* 1 if the fused kernel is instantiated for (integrand, n_dim), else 0. */
int vf_supported(int integrand, int n_dim);
/* Algorithmic fp64 flops per event of the fused iteration. */
double vf_flops_per_event(int mode, int integrand, int n_dim, int plus);

/* Scratch (bytes) for per-block scalar records and the global histogram accumulator.
 * The caller ZERO-INITIALISES the workspace once before its first use; the library leaves the
 * accumulator zeroed after every reduction, so the same workspace is reused across calls. */
size_t vf_workspace_bytes(int n_dim);

/*
 * Fused event kernel.  Replaces, for one chunk of events,
 *   MonteCarloFlow._generate_random_array  (monte_carlo.py:249-275)
 *   _generate_random_array / importance_sampling_digest (vflow.py:93-126, 39-83)
 *   the integrand call, tmp = w*f, tmp2, the two reduce_sums (vflow.py:412-421)
 *   _importance_sampling_array_filling / consume_array_into_indices
 *     (vflow.py:370-387, utils.py:17-44)
 *   and, over chunks, run_event + _accumulate (monte_carlo.py:420-480, 72-92).
 *
 * Evaluates global event indices [ev_begin, ev_begin + n_events).
 * out_sums[0] (+)= sum w*f, out_sums[1] (+)= sum (w*f)^2,
 * out_hist[j*50+b] (+)= sum (w*f)^2 over events whose dim-j bin is b
 * (only when mode == VEGAS and train != 0; may be NULL otherwise).
 * accumulate != 0 adds into the outputs (the reference's chunk semantics),
 * 0 overwrites them.
 * xjac is 1/N of the WHOLE iteration (monte_carlo.py:224-227), also for chunks.
 * xmin/xdelta: [host] arrays of n_dim doubles or both NULL (monte_carlo.py:159-175).
 */
int vf_run_event(int mode, int integrand, int n_dim, uint64_t ev_begin, int64_t n_events,
                 double xjac, uint64_t seed, uint32_t iteration, int train,
                 const double* divisions /*[dev] [n_dim][51], NULL for PLAIN*/,
                 const double* xmin /*[host]*/, const double* xdelta /*[host]*/,
                 double* out_sums /*[dev] [2]*/, double* out_hist /*[dev] [n_dim][50]*/,
                 int accumulate, void* workspace /*[dev]*/, size_t workspace_bytes, void* stream);

/*
 * Grid refinement.  Replaces VegasFlow.refine_grid + refine_grid_per_dimension
 * (vflow.py:349-362, 135-211): smoothing, damping with ALPHA = 1.5, rebinning.
 * divisions is updated in place.
 */
int vf_refine_grid(int n_dim, const double* hist /*[dev] [n_dim][50]*/,
                   double* divisions /*[dev] [n_dim][51]*/, void* stream);

/*
 * Per-iteration epilogue: (res, sigma) of VegasFlow._iteration_content
 * (vflow.py:437-438; identical algebra for PlainFlow, plain.py:37-43) written to
 * result[0..1], followed by vf_refine_grid when train != 0.
 * sums = (sum wf, sum (wf)^2) of the whole iteration.
 */
int vf_iteration_epilogue(int n_dim, int64_t n_events, int train, const double* sums /*[dev] [2]*/,
                          const double* hist /*[dev]*/, double* divisions /*[dev]*/,
                          double* result /*[dev] [2]*/, void* stream);

/*
 * Whole single-device integration loop: n_iter iterations of (fused event kernel over all
 * n_events, block reduction + sigma + grid refinement), enqueued back to back with no host
 * synchronisation.  Replaces the iteration loop of run_integration (monte_carlo.py:679-685)
 * over VegasFlow._iteration_content (vflow.py:432-442) / PlainFlow._run_iteration
 * (plain.py:37-43) together with run_event + _accumulate (monte_carlo.py:420-480, 72-92).
 * Iteration k uses Philox stream `first_iteration + k` and xjac = 1/n_events.
 * results[k] = (res_k, sigma_k); packed = [hist n_dim*50 | sum wf | sum (wf)^2] of the last
 * iteration; divisions is refined in place when train != 0.  The inverse-variance
 * combination (monte_carlo.py:713-732) stays with the caller.  When host_results is not NULL
 * (page-locked host memory) row k is also copied there asynchronously right after iteration k
 * by the iteration's tail kernel (stores through the device alias of the mapped host memory)
 * -- the reference's per-iteration read-back for logging (monte_carlo.py:699-710) without a
 * host synchronisation; the rows are valid once the stream has been synchronised.
 */
int vf_run_iterations(int mode, int integrand, int n_dim, int64_t n_events, uint64_t seed,
                      uint32_t first_iteration, int n_iter, int train,
                      double* divisions /*[dev] in/out, NULL for PLAIN*/,
                      const double* xmin /*[host]*/, const double* xdelta /*[host]*/,
                      double* packed /*[dev] [n_dim*50+2]*/, double* results /*[dev] [n_iter][2]*/,
                      double* host_results /*[host, pinned] [n_iter][2] or NULL*/,
                      void* workspace /*[dev]*/, size_t workspace_bytes, void* stream);

/*
 * n_iter iterations of one rank of a multi-GPU run, each fused with its collective: the event
 * kernel over this rank's events [ev_begin, ev_begin + n_events_local), then ONE kernel that reduces
 * the block partials, exchanges the [n_dim*50+2] sums with all peers by P2P stores over
 * NVLink (one-shot all-reduce on peer-mapped memory; every 16-byte store carries the value and
 * the exchange sequence number, so the data is its own arrival flag: no fence, no flag hop),
 * adds the `world` contributions in rank order, computes (res, sigma) and refines the
 * grid -- every rank ends with bit-identical divisions.  Replaces the joblib device pool and
 * host-side _accumulate of the reference (monte_carlo.py:143-157, 318-365, 454-480, 72-92).
 * peer_buffers: [host] `world` device addresses of every rank's exchange buffer of
 * vf_exchange_bytes(n_dim, world) bytes (zero-initialised, peer-mapped, e.g. torch symmetric
 * memory); iteration k uses exchange sequence number first_seq + k: the sequence must start
 * at 1 and advance by one per iteration, identically on every rank.  world <= 8.
 * results[k] = (res_k, sigma_k).
 */
size_t vf_exchange_bytes(int n_dim, int world, int64_t n_cubes /*0 unless VEGAS+*/);
int vf_run_iterations_sharded(int mode, int integrand, int n_dim, uint64_t ev_begin,
                              int64_t n_events_local, int64_t n_events_total, uint64_t seed,
                              uint32_t first_iteration, int n_iter, int train,
                              double* divisions /*[dev]*/, const double* xmin /*[host]*/,
                              const double* xdelta /*[host]*/, double* packed /*[dev] [n_dim*50+2]*/,
                              double* results /*[dev] [n_iter][2]*/,
                              double* host_results /*[host, pinned] or NULL*/,
                              void* workspace /*[dev]*/, size_t workspace_bytes, int rank, int world,
                              const uint64_t* peer_buffers /*[host] [world]*/, uint64_t first_seq,
                              void* stream);

/*
 * Parity entry: the same device code as vf_run_event, but fed external
 * uniforms at the reference's own RNG/algorithm seam
 * (_digest_random_generation, monte_carlo.py:268) and writing the per-event
 * quantities.  rnds [n][n_dim] row-major in [TECH_CUT, 1-TECH_CUT).
 * Any output pointer may be NULL.
 */
int vf_digest_from_uniforms(int mode, int integrand, int n_dim, int64_t n,
                            const double* rnds /*[dev]*/, const double* divisions /*[dev]*/,
                            double xjac, const double* xmin /*[host]*/,
                            const double* xdelta /*[host]*/, double* x /*[dev] [n][n_dim]*/,
                            double* w /*[dev] [n]*/, int32_t* ind /*[dev] [n][n_dim]*/,
                            double* wf /*[dev] [n]*/, void* stream);

/* The engine's uniforms for events [ev_begin, ev_begin+n): rnds [n][n_dim]; rng_bits 52 | 32.
 * Stream definition (52): Philox4x32-10, key = seed, counter = (event_lo, event_hi, dim/2,
 * iteration); words (2h, 2h+1) of the block fill the mantissa of m in [1,2); v = fma(m, 1-2T, 3T)
 * with T = TECH_CUT; r = 2 - v, exact, a multiple of 2^-52 in (T, 1-T].  The fused kernels
 * consume v directly (xn = fma(v, 50, -50) == 50*(1-r) bit for bit). */
int vf_uniforms(int n_dim, uint64_t ev_begin, int64_t n, uint64_t seed, uint32_t iteration,
                int rng_bits, double* rnds /*[dev]*/, void* stream);

/*
 * Unfused path for integrands that are not built in (the generic
 * `compile(callable)` of monte_carlo.py:489-636): sample writes x, w, ind for
 * the caller to evaluate f(x) itself; accumulate consumes f.
 * Replaces MonteCarloFlow._generate_random_array (monte_carlo.py:249-275) and
 * generate_random_array (monte_carlo.py:229-247).
 */
int vf_sample(int mode, int n_dim, uint64_t ev_begin, int64_t n, double xjac, uint64_t seed,
              uint32_t iteration, const double* divisions /*[dev]*/,
              const double* xmin /*[host]*/, const double* xdelta /*[host]*/,
              double* x /*[dev] [n][n_dim]*/, double* w /*[dev] [n]*/,
              int32_t* ind /*[dev] [n][n_dim], may be NULL*/, void* stream);

/* tmp = w*f, tmp2, sums and histogram (vflow.py:416-428) from caller-evaluated f.
 * ind may be NULL (no histogram).  Same output semantics as vf_run_event. */
int vf_accumulate(int n_dim, int64_t n, const double* w /*[dev]*/, const double* f /*[dev]*/,
                  const int32_t* ind /*[dev] [n][n_dim]*/, int train, double* out_sums /*[dev]*/,
                  double* out_hist /*[dev]*/, int accumulate, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * VEGAS+ (src/vegasflow/vflowplus.py).  Events are ordered by hypercube
 * (vflowplus.py:67); cube c has lexicographic coordinates, dim 0 most
 * significant (vflowplus.py:126-128).  ev_offset is the exclusive prefix sum
 * of n_ev (n_cubes+1 entries).
 * ---------------------------------------------------------------------- */

/*
 * Fused stratified event kernel.  Replaces generate_samples_in_hypercubes
 * (vflowplus.py:46-80) and VegasFlowPlus._run_event (vflowplus.py:187-220).
 * ress[c] += sum wf, ress2[c] += sum (wf)^2 over the events of cube c
 * (caller zeroes them); histogram as vf_run_event.  xjac = 1/n_cubes
 * (vflowplus.py:139).  n_events = ev_offset[n_cubes] (host copy).
 * If rnds != NULL the kernel consumes external uniforms [n_events][n_dim]
 * instead of Philox, and writes any non-NULL x/w/ind/wf (parity entry).
 */
int vfp_run_event(int integrand, int n_dim, int n_strat, int64_t n_cubes, int64_t n_events,
                  const int32_t* n_ev /*[dev]*/, const int64_t* ev_offset /*[dev]*/, double xjac,
                  uint64_t seed, uint32_t iteration, int rng_bits /*52 | 32*/, int train,
                  const double* divisions /*[dev]*/,
                  const double* xmin /*[host]*/, const double* xdelta /*[host]*/,
                  double* ress /*[dev] [n_cubes]*/, double* ress2 /*[dev] [n_cubes]*/,
                  double* out_hist /*[dev]*/, int accumulate, void* workspace,
                  size_t workspace_bytes, const double* rnds /*[dev] or NULL*/,
                  double* x /*[dev]*/, double* w /*[dev]*/, int32_t* ind /*[dev]*/,
                  double* wf /*[dev]*/, void* stream);

/*
 * Per-iteration VEGAS+ epilogue.  Replaces the tail of _run_event
 * (arr_var = ress2*n_ev - ress^2, vflowplus.py:216-217), _iteration_content
 * (res, sigma, vflowplus.py:230-233) and, when adaptive != 0,
 * redistribute_samples (vflowplus.py:153-163; arr_var clamped at 0 before the
 * power -- documented divergence).  Writes arr_var, result[0..1] = (res, sigma),
 * and when adaptive the new n_ev, ev_offset and *n_events_out (device int64).
 */
int vfp_iteration_epilogue(int64_t n_cubes, const double* ress /*[dev]*/,
                           const double* ress2 /*[dev]*/, int adaptive, int min_neval_hcube,
                           int64_t init_calls, int32_t* n_ev /*[dev] in/out*/,
                           int64_t* ev_offset /*[dev] out*/, double* arr_var /*[dev] out*/,
                           double* result /*[dev] [2]*/, int64_t* n_events_out /*[dev]*/,
                           void* stream);

/*
 * n_iter whole VEGAS+ iterations with the sample allocation RESIDENT ON THE DEVICE: per
 * iteration the stratified event kernel (event count read from ev_offset[n_cubes] on the
 * device) and ONE tail kernel doing histogram reduction + grid refinement (vflow.py:349-362),
 * arr_var, (res, sigma) (vflowplus.py:216-233), redistribute_samples + new offsets
 * (vflowplus.py:153-163) and the zeroing of the per-cube sums -- no host synchronisation and no
 * host copy of n_events between iterations.  Replaces VegasFlowPlus._iteration_content
 * (vflowplus.py:222-242) inside the loop of run_integration (monte_carlo.py:679-685), including
 * the `n_events` setter + recompile of the reference (monte_carlo.py:187-193).
 * results[k] = (res_k, sigma_k, n_events of iteration k+1); host_results as vf_run_iterations.
 * ress / ress2 must be zero on entry and are zero on exit; xjac = 1/n_cubes (vflowplus.py:139).
 * world > 1 (the reference is single-device, vflowplus.py:88-100): rank r evaluates
 * the events of a contiguous cube range balanced on the event prefix sum; the tail kernel
 * all-reduces the histogram and the partial (res, sigma^2) and all-gathers the per-cube
 * variances through the peer buffers (vf_exchange_bytes(n_dim, world, n_cubes) bytes each,
 * zero-initialised), then every rank redistributes redundantly: n_ev stays bit-identical.
 */
int vfp_run_iterations(int integrand, int n_dim, int n_strat, int64_t n_cubes, uint64_t seed,
                       uint32_t first_iteration, int n_iter, int rng_bits /*52 | 32*/, int train,
                       int adaptive, int min_neval_hcube, int64_t init_calls,
                       double* divisions /*[dev] in/out*/, const double* xmin /*[host]*/,
                       const double* xdelta /*[host]*/, int32_t* n_ev /*[dev] in/out*/,
                       int64_t* ev_offset /*[dev] in/out [n_cubes+1]*/,
                       double* ress /*[dev] [n_cubes]*/, double* ress2 /*[dev] [n_cubes]*/,
                       double* arr_var /*[dev] out [n_cubes]*/,
                       double* out_hist /*[dev] [n_dim][50]*/,
                       double* results /*[dev] [n_iter][3]*/,
                       double* host_results /*[host, pinned] [n_iter][3] or NULL*/,
                       void* workspace /*[dev]*/, size_t workspace_bytes, int rank, int world,
                       const uint64_t* peer_buffers /*[host] [world] or NULL*/, uint64_t first_seq,
                       void* stream);

/* ------------------------------------------------------------------------
 * Measurement helpers (no reference counterpart)
 * ---------------------------------------------------------------------- */

/* Dependent-free DFMA chains on every SM; returns achieved fp64 TFLOP/s in *tflops
 * (device-timed with CUDA events, synchronous). */
int vf_fp64_peak_probe(int iters, double* tflops /*[host]*/);
/* Bracket every event-kernel launch of the calling thread with CUDA events on its launching
 * stream (enable != 0 starts a fresh measurement, 0 stops); vf_kernel_time_ms synchronises on
 * the recorded events and returns their summed duration and count. */
int vf_kernel_timing(int enable);
int vf_kernel_time_ms(double* total_ms /*[host]*/, int* launches /*[host]*/);
/* Same for the reduce/exchange/refine kernels that follow each event kernel. */
int vf_epilogue_time_ms(double* total_ms /*[host]*/, int* launches /*[host]*/);
/* Number of SMs of the current device. */
int vf_sm_count(void);
/* Kernels launched by this library on the calling thread since the last reset. */
int64_t vf_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* VEGASFLOW_B200_H */
