// Fused VEGAS event kernels (sm_100a): Philox -> grid map -> integrand ->
// reductions + shared-memory histograms, one launch per chunk of events.
// Reference citations are file:line relative to /root/reference.
#pragma once
#include <atomic>
#include <type_traits>

#include "vf_common.cuh"
#include "vf_integrands.cuh"

namespace vf {

// Launch configuration per (dimension count, register class).  Shared memory per block:
//   table   NDIM*50*TC*16 B    (x_ini, Delta) pairs, TC lane-interleaved copies
//   pairs   NP*50*50*JC*8 B    one 50x50 histogram per PAIR of dimensions (2p, 2p+1), JC copies
//   singles (NDIM-2NP)*50*HC*8 B  one 50-bin histogram per remaining dimension, HC copies
// Interleaving by (lane % copies) keeps the LDS.128 table reads and the 64-bit updates of the
// single histograms bank-conflict free for any bin pattern (TC >= 8, HC >= 16).
//
// PAIR HISTOGRAMS.  The reference fills one histogram per dimension with (w f)^2 (vflow.py:
// 370-387); every update is a shared-memory read-modify-write loop (load, DADD, ATOMS.CAST.SPIN,
// branch: sm_100 has no native shared fp64 add), seven instructions per dimension per event.  A
// 50x50 histogram over the bin PAIR of two dimensions takes ONE update for both, and its row and
// column sums -- taken once per launch, in a fixed order -- are exactly the two per-dimension
// histograms (sums of non-negative terms: the different summation order is ~1e-16 relative).
// Half the updates: -4.1 % kernel time at d = 8, -5.8 % at d = 4, -7.4 % at d = 10
// (profiles/r2_k1_pairs.txt), although the 2500-cell histograms are hit with bank conflicts
// (random cells) and leave room for one or two copies only.  Used for light integrands with
// d <= 10, d = 12 and d = 14 (choose_smem); the remaining dimension of an odd n_dim, every
// dimension of the other shapes and of the heavy integrands keep per-dimension histograms.
//
// ONE block per SM for every shape; measured on B200 (profiles/r2_k1_variants.txt,
// r2_k1_threads.txt):
//   * 1024 threads with the 64-register budget beat 512 threads with 128 registers for every
//     light integrand (-10 % at d = 9 ... 16; -2 % against two 512-thread blocks at d <= 8) and,
//     since the kernels got leaner (62 registers, no spill even at d = 20), 768 threads at
//     d = 19, 20 as well (-6 %, -9 %); 896 threads (72 registers) win at d = 10, 12, 13, 16
//     (-5, -2.7, -1.5, -1.8 %) and are the unmeasured default beyond d = 20 (user integrands);
//   * per-dimension histograms: copies are what the shared-memory atomics need most, HC = 32 (one
//     copy per lane, no same-address collisions inside a warp) wherever it fits, else 16;
//   * heavy integrands keep 512 threads and the 128-register budget unless they name a block
//     size themselves (`kBlockThreads`).
constexpr size_t kSmemBudget = 227 * 1024 - 2048;  // opt-in limit minus static + reserved
struct SmemChoice {
    int np, jc, tc, hc;  // dimension pairs, copies of a pair histogram, table copies, single copies
};
constexpr size_t cfg_smem_bytes(int n_dim, SmemChoice c) {
    return (size_t)n_dim * kBins * 16 * c.tc + (size_t)c.np * kBins * kBins * 8 * c.jc +
           (size_t)(n_dim - 2 * c.np) * kBins * 8 * c.hc;
}
constexpr SmemChoice choose_smem(int n_dim, bool heavy) {
    // measured, d = 2 ... 20 (profiles/r2_k1_pairs.txt): pairs pay with two copies, or with one
    // copy where they free the shared memory for a 16-copy table (d = 10) or at least keep 8
    // (d = 12, 14); odd d gain < 1 %, d = 11 and 13 lose 1-5 %, and from d = 15 on the pairs would
    // squeeze the table to 4 copies (+16 ... 23 %), as does any partial pairing at d = 16 ... 20
    if (!heavy && n_dim >= 2 && (n_dim <= 10 || n_dim == 12 || n_dim == 14)) {
        const int np = n_dim / 2;
        // d = 4: an 8-copy table is as fast as 16 copies (0.934 vs 0.949 ms per 1e8 events) and
        // keeps the block at 106 KB: two blocks fit one SM, so the next iteration's event kernel
        // (programmatic dependent launch) can get resident beside the running one
        if (n_dim == 4) return SmemChoice{np, 2, 8, 32};
        const SmemChoice cands[] = {{np, 2, 16, 32}, {np, 2, 8, 32}, {np, 1, 16, 32}, {np, 1, 8, 32}};
        for (const SmemChoice& c : cands)
            if (cfg_smem_bytes(n_dim, c) <= kSmemBudget) return c;
    }
    // one histogram per dimension; the table gets what is left of the 227 KB (TC = 16 up to
    // d = 8, 8 up to d = 12 and d = 15 ... 18, else 4)
    const int hc = cfg_smem_bytes(n_dim, SmemChoice{0, 0, 4, 32}) <= kSmemBudget ? 32 : 16;
    const int tc = (n_dim <= 8 && cfg_smem_bytes(n_dim, SmemChoice{0, 0, 16, hc}) <= kSmemBudget)
                       ? 16
                       : (cfg_smem_bytes(n_dim, SmemChoice{0, 0, 8, hc}) <= kSmemBudget ? 8 : 4);
    return SmemChoice{0, 1, tc, hc};
}
template <int NDIM, bool HEAVY, int THREADS = 0>
struct CfgT {
    static constexpr int kThreads =
        THREADS ? THREADS
                : (HEAVY ? 512
                         : ((NDIM == 10 || NDIM == 12 || NDIM == 13 || NDIM == 16 || NDIM > 20) ? 896
                                                                                                : 1024));
    static constexpr SmemChoice kChoice = choose_smem(NDIM, HEAVY);
    static constexpr int NP = kChoice.np;        // dimensions 0 ... 2NP-1 are paired
    static constexpr int JC = kChoice.jc;
    static constexpr int TC = kChoice.tc;
    static constexpr int HC = kChoice.hc;
    static constexpr int kSingles = NDIM - 2 * NP;
    static constexpr int kTblEntries = NDIM * kBins * TC;
    static constexpr int kPairEntries = NP * kBins * kBins * JC;
    static constexpr int kHistEntries = kPairEntries + kSingles * kBins * HC;
    static constexpr size_t kSmemBytes = cfg_smem_bytes(NDIM, kChoice);
    static_assert(kSmemBytes <= kSmemBudget, "n_dim too large for the fused kernels");
};
// an integrand may fix its block size with `static constexpr int kBlockThreads`
template <class I, class = void>
struct BlockThreadsOf {
    static constexpr int value = 0;
};
template <class I>
struct BlockThreadsOf<I, std::void_t<decltype(I::kBlockThreads)>> {
    static constexpr int value = I::kBlockThreads;
};
template <class I, int NDIM>
using Cfg = CfgT<NDIM, I::kHeavy, BlockThreadsOf<I>::value>;
// The VEGAS+ kernel carries the cube coordinates and per-cube divisors on top of the event
// kernel's state (64 registers spill 56 B at d = 8, 120 B at d = 10): it takes a block size of its
// own, measured on uniform allocations of 5e7 events (profiles/r2_plus_variants.txt): 640 threads
// (84 ... 96 registers) -9 % at d = 4, -4 % at d = 8 and d = 12; 768 threads -2 % at d = 6; the
// event kernel's block size at d = 10 ... 16; 768 threads from d = 17 on, where 64 registers would
// spill hundreds of bytes.
constexpr int plus_threads(int n_dim) {
    return n_dim <= 4 ? 640
                      : (n_dim <= 7 ? 768 : (n_dim <= 9 || n_dim == 12 ? 640 : (n_dim >= 17 ? 768 : 0)));
}
template <class I, int NDIM>
using PlusCfg = CfgT<NDIM, I::kHeavy,
                     BlockThreadsOf<I>::value ? BlockThreadsOf<I>::value
                                              : (I::kHeavy ? 0 : plus_threads(NDIM))>;

struct EventKernelArgs {
    const double* divisions;  // [NDIM][51]
    double* partials;         // workspace: scalars[grid][2] | acc[NDIM*50]
    uint64_t ev_begin, ev_end;
    double xjac;
    uint32_t iteration;
    int train;
    PhiloxKeys pk;
    Limits lim;
    IntegrandConsts ic;
};

// Zero the histogram copies (independent of the previous kernel), wait for the previous kernel
// of the stream (it refines `divisions`), then stage the (x_ini, Delta) pairs.
template <class C>
__device__ __forceinline__ void stage_grid(const double* __restrict__ divisions, double2* tbl,
                                           double* hist, bool zero_hist) {
    if (zero_hist)
        for (int i = threadIdx.x; i < C::kHistEntries; i += blockDim.x) hist[i] = 0.0;
    pdl_wait();
    // one thread per (dimension, bin): a single round trip to L2 for the two edges, then the TC
    // copies are written from registers
    for (int jb = threadIdx.x; jb < C::kTblEntries / C::TC; jb += blockDim.x) {
        const int j = jb / kBins, b = jb - j * kBins;
        const double x_ini = divisions[j * kEdges + b];      // vflow.py:70
        const double x_fin = divisions[j * kEdges + b + 1];  // vflow.py:71
        const double2 e = make_double2(x_ini, __dsub_rn(x_fin, x_ini));  // vflow.py:73
#pragma unroll
        for (int c = 0; c < C::TC; ++c) tbl[jb * C::TC + c] = e;
    }
}

// monte_carlo.py:270-274: w *= xjac; x = xmin + x*xdelta; w *= prod(xdelta)
template <int NDIM>
__device__ __forceinline__ double apply_jacobians(double w, double (&x)[NDIM], double xjac,
                                                  const Limits& lim) {
    w = __dmul_rn(w, xjac);
    if (lim.has) {
#pragma unroll
        for (int j = 0; j < NDIM; ++j)
            x[j] = __dadd_rn(lim.xmin[j], __dmul_rn(x[j], lim.xdelta[j]));
        w = __dmul_rn(w, lim.xdeltajac);
    }
    return w;
}

// Block-level tail shared by the event kernels: the two scalars are reduced in a fixed order
// into this block's record; the histogram copies are reduced in-block and added to the global
// accumulator with one fp64 RED per (dimension, bin).
template <class C, int NDIM>
__device__ __forceinline__ void write_partials(double sum, double sum2, const double* hist,
                                               bool with_hist, double* workspace) {
    __shared__ double red[2][C::kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    sum = warp_sum(sum);
    sum2 = warp_sum(sum2);
    if (lane == 0) {
        red[0][warp] = sum;
        red[1][warp] = sum2;
    }
    __syncthreads();  // also orders the shared-memory histogram updates
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int w = 0; w < C::kThreads / 32; ++w) t += red[threadIdx.x][w];
        workspace[(size_t)blockIdx.x * 2 + threadIdx.x] = t;
    }
    if (with_hist) {
        double* acc = workspace + ws_acc_offset();
        // paired dimensions: row / column sums of the pair histograms.  Four lanes share one
        // (dimension, bin): lane q adds the cells k = q, q+4, ... of every copy, the quad is
        // folded with two shuffles (a fixed order), its first lane issues the RED.  The loop
        // bound is warp-uniform so that every lane reaches the shuffles.
        constexpr int kPairTasks = 2 * C::NP * kBins * 4;
        for (int t0 = (threadIdx.x & ~31); t0 < kPairTasks; t0 += blockDim.x) {
            const int t = t0 + lane;
            double part = 0.0;
            const int i = t >> 2, q = t & 3;
            if (t < kPairTasks) {
                const int j = i / kBins, b = i - j * kBins;
                const double* h = hist + (size_t)(j >> 1) * (kBins * kBins * C::JC);
                const int first = (j & 1) ? b : b * kBins, step = (j & 1) ? kBins : 1;
                for (int k = q; k < kBins; k += 4) {
                    int kk = k + b;  // rotate the start: neighbouring quads read different banks
                    if (kk >= kBins) kk -= kBins;
                    const double* cell = h + (first + kk * step) * C::JC;
#pragma unroll
                    for (int c = 0; c < C::JC; ++c) part += cell[c];
                }
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (q == 0 && t < kPairTasks) atomicAdd(acc + i, part);  // RED.E.ADD.F64
        }
        // single dimensions: reduce the HC copies
        const double* hs = hist + C::kPairEntries;
        for (int i = threadIdx.x; i < C::kSingles * kBins; i += blockDim.x) {
            double t = 0.0;
            // thread i starts at copy i: the lanes of a warp read different banks (a plain
            // c = 0.. walk has all 32 lanes on one bank: the rows are HC*8 = 256 B apart)
#pragma unroll
            for (int c = 0; c < C::HC; ++c) t += hs[i * C::HC + ((c + i) & (C::HC - 1))];
            atomicAdd(acc + 2 * C::NP * kBins + i, t);  // RED.E.ADD.F64
        }
    }
}

// Histogram update of one event: hist[j][bin_j][slot] += tmp2 for every dimension
// (vflow.py:370-387, utils.py:40-43).  sm_100 has no native shared-memory fp64 add: the
// reduction compiles to a load / DADD / ATOMS.CAST.SPIN loop per dimension.  That plain form
// measured 4 % faster over the whole kernel than issuing the eight loads, adds and ATOMS.CAS.64
// back to back with a common retry path (round 1), and faster than conflict-avoiding variants --
// half-warps on disjoint dimensions, one dimension at a time, 32 copies
// (profiles/r2_k1_variants.txt): ~5.7 CAS retries per warp-event happen either way and the cost
// is the instruction count.
// What the event loop keeps per dimension until the update (`keep`): the BIN for a paired
// dimension; for a single dimension the shared-window address of the event's TABLE row when the
// table and histogram row pitches agree and nothing is paired (then the histogram cell is row + a
// per-lane constant, one IADD3), else of its histogram row.  Dimension offsets fold into the
// address immediates.
template <class C>
struct HistAddr {
    static constexpr bool kSamePitch = (C::NP == 0) && (C::TC * 16 == C::HC * 8);
    uint32_t tbl_s, pair_s, hist_s, delta;
    __device__ __forceinline__ HistAddr(const void* tbl, const void* hist, int lane) {
        tbl_s = smem_u32(tbl) + (uint32_t)(lane % C::TC) * 16u;
        pair_s = smem_u32(hist) + (uint32_t)(lane % C::JC) * 8u;
        hist_s = smem_u32(hist) + (uint32_t)C::kPairEntries * 8u + (uint32_t)(lane % C::HC) * 8u;
        delta = hist_s - tbl_s;
    }
    __device__ __forceinline__ uint32_t keep(int j, int bin, uint32_t tbl_row) const {
        if (j < 2 * C::NP) return (uint32_t)bin;
        return kSamePitch ? tbl_row : row_addr<C::HC * 8>(bin, hist_s);
    }
};
template <class C, int NDIM>
__device__ __forceinline__ void hist_update(const HistAddr<C>& ha, const uint32_t (&row)[NDIM],
                                            double tmp2) {
#pragma unroll
    for (int p = 0; p < C::NP; ++p) {  // one update for dimensions 2p and 2p+1
        const uint32_t cell = row[2 * p] * (uint32_t)kBins + row[2 * p + 1];
        red_shared_f64(cell * (uint32_t)(C::JC * 8) + ha.pair_s,
                       (uint32_t)p * (kBins * kBins * C::JC * 8), tmp2);
    }
#pragma unroll
    for (int j = 2 * C::NP; j < NDIM; ++j)
        red_shared_f64(HistAddr<C>::kSamePitch ? row[j] + ha.delta : row[j],
                       (uint32_t)(j - 2 * C::NP) * (kBins * C::HC * 8), tmp2);
}

// ---------------------------------------------------------------------------
// K1: fused event kernel (VegasFlow._run_event, vflow.py:389-430; PlainFlow
// plain.py:18-35).  Thread t evaluates global events ev_begin + t, + stride...
// ---------------------------------------------------------------------------
// FAST = the training iteration on the unit hypercube (train != 0, no integration limits): the
// two launch-uniform branches of the generic kernel are compiled out (2 % of the kernel time --
// they fence the instruction scheduler; profiles/r2_k1_variants.txt).  The host picks it.
template <class I, int NDIM, int MODE, int RB, bool FAST = false>
__global__ void __launch_bounds__(Cfg<I, NDIM>::kThreads, 1)
event_kernel(const __grid_constant__ EventKernelArgs a) {
    using C = Cfg<I, NDIM>;
    static_assert(!FAST || MODE == VF_MODE_VEGAS, "the fast path is the VEGAS training iteration");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tbl = reinterpret_cast<double2*>(smem_raw);
    double* hist = reinterpret_cast<double*>(smem_raw + (size_t)C::kTblEntries * 16);
    const bool do_hist = FAST || ((MODE == VF_MODE_VEGAS) && a.train);
    pdl_launch_dependents();  // the reduce/refine kernel may get resident behind this grid
    if (MODE == VF_MODE_VEGAS) {
        stage_grid<C>(a.divisions, tbl, hist, do_hist);
        __syncthreads();
    } else {
        pdl_wait();  // the previous tail kernel still reads the per-block records
    }
    const int lane = threadIdx.x & 31;
    const HistAddr<C> ha(tbl, hist, lane);
    double sum = 0.0, sum2 = 0.0;
    // 32-bit trip count, 64-bit event index advanced by the launch-constant stride (a 64-bit
    // compare and a re-derived stride per event cost six integer instructions)
    const uint32_t stride = gridDim.x * C::kThreads;
    uint64_t n = a.ev_begin + (uint64_t)blockIdx.x * C::kThreads + threadIdx.x;
    const uint32_t trips = n < a.ev_end ? (uint32_t)((a.ev_end - n + stride - 1) / stride) : 0u;
    const uint32_t expo = a.pk.expo;
    auto one_event = [&](const uint64_t n) {
        double x[NDIM];
        uint32_t row[NDIM];
        double w = 1.0;
        constexpr int PC = Rng<RB>::kPerCall;  // uniforms per Philox block
#pragma unroll
        for (int p = 0; p < (NDIM + PC - 1) / PC; ++p) {
            const uint4 o = philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)p,
                                          a.iteration, a.pk);
#pragma unroll
            for (int h = 0; h < PC; ++h) {
                const int j = PC * p + h;
                if (j < NDIM) {
                    const double v = Rng<RB>::v(o, h, expo);  // the uniform is r = 2 - v, exactly
                    if (MODE == VF_MODE_VEGAS) {
                        // vflow.py:117: rn(50*(1-r)) with 1-r = v-1 exact == rn(50*v - 50)
                        const double xn = fma(v, kFBins, -kFBins);
                        double wfac;
                        int bin;
                        uint32_t trow;
                        vegas_map_dim_s<C::TC>(xn, ha.tbl_s,
                                               (uint32_t)j * (kBins * C::TC * 16),
                                               x[j], wfac, bin, trow);
                        row[j] = ha.keep(j, bin, trow);
                        w = (j == 0) ? wfac : __dmul_rn(w, wfac);  // reduce_prod, vflow.py:78
                    } else {
                        x[j] = __dsub_rn(2.0, v);  // r itself, monte_carlo.py:290-298
                    }
                }
            }
        }
        if (FAST) w = __dmul_rn(w, a.xjac);  // monte_carlo.py:270, no limits
        else w = apply_jacobians<NDIM>(w, x, a.xjac, a.lim);
        const double f = I::template eval<NDIM>(x, a.ic);  // vflow.py:412
        const double tmp = __dmul_rn(w, f);                // vflow.py:416
        const double tmp2 = __dmul_rn(tmp, tmp);           // vflow.py:417
        sum += tmp;                                        // vflow.py:420
        sum2 += tmp2;                                      // vflow.py:421
        if (do_hist) hist_update<C, NDIM>(ha, row, tmp2);
    };
    // Static striding.  The warps of a block leave this loop over the last 11 % of its duration
    // (the scheduler favours high warp ids), but handing the events out dynamically so that they
    // leave together is 1-2 % SLOWER: the kernel is bound by pipe throughput, not by the number
    // of resident warps (profiles/r2_k1_r3_tail_dynamic_variants.txt).
    for (uint32_t trip = 0; trip < trips; ++trip, n += stride) one_event(n);
    write_partials<C, NDIM>(sum, sum2, hist, do_hist, a.partials);
}

// ---------------------------------------------------------------------------
// K4: parity kernel -- same device code, external uniforms, per-event outputs.
// ---------------------------------------------------------------------------
struct DigestKernelArgs {
    const double* rnds;       // [n][NDIM]
    const double* divisions;
    double* x;
    double* w;
    int32_t* ind;
    double* wf;
    int64_t n;
    double xjac;
    Limits lim;
    IntegrandConsts ic;
};

template <class I, int NDIM, int MODE>
__global__ void __launch_bounds__(Cfg<I, NDIM>::kThreads, 1)
digest_kernel(const __grid_constant__ DigestKernelArgs a) {
    using C = Cfg<I, NDIM>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tbl = reinterpret_cast<double2*>(smem_raw);
    if (MODE == VF_MODE_VEGAS) {
        stage_grid<C>(a.divisions, tbl, nullptr, false);
        __syncthreads();
    }
    const char* tbl_lane = reinterpret_cast<const char*>(tbl) + ((threadIdx.x & 31) % C::TC) * 16;
    for (int64_t n = (int64_t)blockIdx.x * C::kThreads + threadIdx.x; n < a.n;
         n += (int64_t)gridDim.x * C::kThreads) {
        double x[NDIM];
        int bin[NDIM];
        double w = 1.0;
#pragma unroll
        for (int j = 0; j < NDIM; ++j) {
            const double r = a.rnds[n * NDIM + j];
            if (MODE == VF_MODE_VEGAS) {
                const double xn = __dmul_rn(kFBins, __dsub_rn(1.0, r));
                double wfac;
                vegas_map_dim<C::TC>(xn, tbl_lane + j * (kBins * C::TC * 16), x[j], wfac, bin[j]);
                w = (j == 0) ? wfac : __dmul_rn(w, wfac);
            } else {
                x[j] = r;
                bin[j] = 0;
            }
        }
        w = apply_jacobians<NDIM>(w, x, a.xjac, a.lim);
        const double f = I::template eval<NDIM>(x, a.ic);
        const double tmp = __dmul_rn(w, f);
#pragma unroll
        for (int j = 0; j < NDIM; ++j) {
            if (a.x) a.x[n * NDIM + j] = x[j];
            if (a.ind) a.ind[n * NDIM + j] = bin[j];
        }
        if (a.w) a.w[n] = w;
        if (a.wf) a.wf[n] = tmp;
    }
}

// ---------------------------------------------------------------------------
// VEGAS+ fused event kernel (generate_samples_in_hypercubes vflowplus.py:46-80,
// VegasFlowPlus._run_event vflowplus.py:187-220).  Events are ordered by cube;
// each block owns a contiguous slice of the event range, its threads stride
// through it, and a thread re-locates its cube only when it walks past the end
// of the current one.  Per-cube sums go to global memory with fp64 RED.
// ---------------------------------------------------------------------------
struct PlusKernelArgs {
    const double* divisions;
    double* partials;
    const int32_t* n_ev;       // [n_cubes]
    const int64_t* ev_offset;  // [n_cubes+1]
    double* ress;              // [n_cubes]
    double* ress2;             // [n_cubes]
    const double* rnds;        // external uniforms (EXT) or null
    double* x;
    double* w;
    int32_t* ind;
    double* wf;
    int64_t n_cubes;
    int64_t n_events;  // < 0: read ev_offset[n_cubes] on the device (no host copy of the count)
    int n_strat;
    int rank, world;   // multi-GPU: this rank evaluates the events of its cube range
    double xjac;
    uint32_t iteration;
    int train;
    PhiloxKeys pk;
    Limits lim;
    IntegrandConsts ic;
};

template <class I, int NDIM, bool EXT, int RB>
__global__ void __launch_bounds__(PlusCfg<I, NDIM>::kThreads, 1)
plus_event_kernel(const __grid_constant__ PlusKernelArgs a) {
    using C = PlusCfg<I, NDIM>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tbl = reinterpret_cast<double2*>(smem_raw);
    double* hist = reinterpret_cast<double*>(smem_raw + (size_t)C::kTblEntries * 16);
    const bool do_hist = a.train != 0;
    pdl_launch_dependents();
    stage_grid<C>(a.divisions, tbl, hist, do_hist);  // waits for the previous tail kernel
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const HistAddr<C> ha(tbl, hist, lane);
    const uint32_t expo = a.pk.expo;
    // the sample allocation lives on the device: after redistribute_samples the event count of
    // the next iteration is ev_offset[n_cubes] (vflowplus.py:163), never copied to the host
    const int64_t n_events = a.n_events >= 0 ? a.n_events : a.ev_offset[a.n_cubes];
    int64_t ev_lo = 0, ev_hi = n_events;
    if (a.world > 1) {  // contiguous cube range balanced on the event prefix sum
        ev_lo = a.ev_offset[first_cube_at_or_after(a.ev_offset, a.n_cubes,
                                                   n_events * a.rank / a.world)];
        ev_hi = a.ev_offset[first_cube_at_or_after(a.ev_offset, a.n_cubes,
                                                   n_events * (a.rank + 1) / a.world)];
    }
    // The block's slice is cut into one CONTIGUOUS range per warp, lanes striding by 32: the
    // lanes of a warp sit in the same cube almost always and a lane walks past the end of its
    // cube once per n_ev/32 events.  (Striding the whole block through the slice made every
    // thread re-locate its cube every n_ev/1024 ~ 8 events at 1e8 events / 6561 cubes: search,
    // coordinates and the two REDs were 21 % of the executed instructions,
    // profiles/r2_prof_c3_details.txt.)
    const int64_t per_block = (ev_hi - ev_lo + gridDim.x - 1) / gridDim.x;
    const int64_t bbegin = ev_lo + (int64_t)blockIdx.x * per_block;
    const int64_t bend = min(bbegin + per_block, ev_hi);
    constexpr int kWarps = C::kThreads / 32;
    const int64_t per_warp = bend > bbegin ? (bend - bbegin + kWarps - 1) / kWarps : 0;
    const int64_t begin = bbegin + (int64_t)(threadIdx.x >> 5) * per_warp;
    const int64_t end = min(begin + per_warp, bend);
    const double fstrat = (double)a.n_strat;
    const double rstrat = __ddiv_rn(1.0, fstrat);
    const uint32_t ustrat = (uint32_t)a.n_strat;

    int64_t cube = -1, hi = 0;
    double s1 = 0.0, s2 = 0.0, fn = 1.0, rfn = 1.0;
    double coords[NDIM];
    for (int64_t e = begin + lane; e < end; e += 32) {
        if (e >= hi) {
            if (cube >= 0) {
                atomicAdd(&a.ress[cube], s1);   // segment_sum, vflowplus.py:213
                atomicAdd(&a.ress2[cube], s2);  // vflowplus.py:214
                s1 = 0.0;
                s2 = 0.0;
            }
            // largest c with ev_offset[c] <= e  (tf.repeat(arange, n_ev), vflowplus.py:67):
            // gallop forward from the current cube (the next one, nearly always), then bisect
            int64_t lo_c = cube < 0 ? 0 : cube + 1, hi_c = a.n_cubes;  // answer in [lo_c, hi_c)
            if (cube >= 0) {
                int64_t step = 1;
                while (lo_c + step < hi_c && a.ev_offset[lo_c + step] <= e) {
                    lo_c += step;
                    step <<= 1;
                }
                hi_c = min(hi_c, lo_c + step);
            }
            while (lo_c < hi_c - 1) {
                const int64_t mid = (lo_c + hi_c) >> 1;
                if (a.ev_offset[mid] <= e) lo_c = mid; else hi_c = mid;
            }
            cube = lo_c;
            hi = a.ev_offset[cube + 1];
            fn = (double)a.n_ev[cube];  // vflowplus.py:69
            rfn = __ddiv_rn(1.0, fn);
            uint32_t rem = (uint32_t)cube;  // itertools.product order, vflowplus.py:126-128
#pragma unroll
            for (int j = NDIM - 1; j >= 0; --j) {
                const uint32_t q = rem / ustrat;
                coords[j] = (double)(rem - q * ustrat + (EXT ? 0u : 2u));
                rem = q;
            }
        }
        double x[NDIM];
        int bin[NDIM];
        uint32_t row[NDIM];
        double w = 1.0;
        constexpr int PC = Rng<RB>::kPerCall;
#pragma unroll
        for (int p = 0; p < (NDIM + PC - 1) / PC; ++p) {
            uint4 o;
            if (!EXT)
                o = philox4x32_10((uint32_t)e, (uint32_t)((uint64_t)e >> 32), (uint32_t)p,
                                  a.iteration, a.pk);
#pragma unroll
            for (int h = 0; h < PC; ++h) {
                const int j = PC * p + h;
                if (j < NDIM) {
                    // vflowplus.py:72: (points + rnds) * FBINS / n_strat.  On the engine's own
                    // stream r = 2 - v exactly, and coords[] holds points + 2 (an exact small
                    // integer): rn(points + r) == rn((points + 2) - v)
                    const double pr = EXT ? __dadd_rn(coords[j], a.rnds[e * NDIM + j])
                                          : __dsub_rn(coords[j], Rng<RB>::v(o, h, expo));
                    const double xn = div_rn_by(__dmul_rn(pr, kFBins), fstrat, rstrat);
                    double wfac;
                    uint32_t trow;
                    vegas_map_dim_s<C::TC>(xn, ha.tbl_s,
                                           (uint32_t)j * (kBins * C::TC * 16), x[j],
                                           wfac, bin[j], trow);
                    row[j] = ha.keep(j, bin[j], trow);
                    w = (j == 0) ? wfac : __dmul_rn(w, wfac);
                }
            }
        }
        w = div_rn_by(w, fn, rfn);  // vflowplus.py:77 (correctly rounded, see div_rn_by)
        w = apply_jacobians<NDIM>(w, x, a.xjac, a.lim);
        const double f = I::template eval<NDIM>(x, a.ic);
        const double tmp = __dmul_rn(w, f);       // vflowplus.py:209
        const double tmp2 = __dmul_rn(tmp, tmp);  // vflowplus.py:210
        s1 += tmp;
        s2 += tmp2;
        if (do_hist) hist_update<C, NDIM>(ha, row, tmp2);
        if (EXT) {
#pragma unroll
            for (int j = 0; j < NDIM; ++j) {
                if (a.x) a.x[e * NDIM + j] = x[j];
                if (a.ind) a.ind[e * NDIM + j] = bin[j];
            }
            if (a.w) a.w[e] = w;
            if (a.wf) a.wf[e] = tmp;
        }
    }
    if (cube >= 0) {
        atomicAdd(&a.ress[cube], s1);
        atomicAdd(&a.ress2[cube], s2);
    }
    // per-cube sums went to ress/ress2; the block record carries no scalars here
    write_partials<C, NDIM>(0.0, 0.0, hist, do_hist, a.partials);
}

// ---------------------------------------------------------------------------
// Host-side launchers (one explicit instantiation per integrand TU).
// ---------------------------------------------------------------------------
struct EventLaunch {
    int mode, n_dim;
    int rng_bits = 52;
    EventKernelArgs k;
    cudaStream_t stream;
    int* nblocks_out;
};
struct DigestLaunch {
    int mode, n_dim;
    DigestKernelArgs k;
    cudaStream_t stream;
};
struct PlusLaunch {
    int n_dim;
    int rng_bits = 52;
    PlusKernelArgs k;
    cudaStream_t stream;
    int* nblocks_out;
};

template <class I> int launch_event(const EventLaunch& L);
template <class I> int launch_digest(const DigestLaunch& L);
template <class I> int launch_plus(const PlusLaunch& L);
template <class I> int supported_dim(int n_dim);

// dims with a fused instantiation for the dimension-generic integrands
#define VF_FOREACH_DIM(X)                                                                  \
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) \
    X(17) X(18) X(19) X(20)

inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) dev = 0;
    return dev;
}

inline int grid_blocks_for(int64_t n_events, int threads, int min_blocks_per_sm) {
    const int64_t max_blocks = (int64_t)sm_count() * min_blocks_per_sm;
    // at least ~4 events per thread before adding blocks
    int64_t want = (n_events + (int64_t)threads * 4 - 1) / ((int64_t)threads * 4);
    if (want < 1) want = 1;
    if (want > max_blocks) want = max_blocks;
    if (want > kMaxBlocks) want = kMaxBlocks;
    return (int)want;
}

template <class I, int NDIM, int MODE, int RB, bool FAST = false>
int launch_event_dim(const EventLaunch& L) {
    using C = Cfg<I, NDIM>;
    auto kern = event_kernel<I, NDIM, MODE, RB, FAST>;
    const size_t smem = MODE == VF_MODE_VEGAS ? C::kSmemBytes : 0;
    // opt in to > 48 KB dynamic shared memory once per device (the attribute is per device)
    static std::atomic<uint64_t> configured{0};
    const int dev = current_device();
    if (!(configured.load(std::memory_order_acquire) >> dev & 1ull)) {
        VF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)C::kSmemBytes));
        configured.fetch_or(1ull << dev, std::memory_order_release);
    }
    const int64_t n = (int64_t)(L.k.ev_end - L.k.ev_begin);
    const int blocks = grid_blocks_for(n, C::kThreads, 1);
    timing_begin(L.stream);
    VF_CUDA_CHECK(launch_pdl(kern, blocks, C::kThreads, smem, L.stream, L.k));
    timing_end(L.stream);
    count_launch();
    *L.nblocks_out = blocks;
    return VF_OK;
}

template <class I, int NDIM, int MODE>
int launch_digest_dim(const DigestLaunch& L) {
    using C = Cfg<I, NDIM>;
    auto kern = digest_kernel<I, NDIM, MODE>;
    // opt in to > 48 KB dynamic shared memory once per device (the attribute is per device)
    static std::atomic<uint64_t> configured{0};
    const int dev = current_device();
    if (!(configured.load(std::memory_order_acquire) >> dev & 1ull)) {
        VF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)C::kSmemBytes));
        configured.fetch_or(1ull << dev, std::memory_order_release);
    }
    const int blocks = grid_blocks_for(L.k.n, C::kThreads, 1);
    kern<<<blocks, C::kThreads, C::kSmemBytes, L.stream>>>(L.k);
    count_launch();
    VF_CUDA_CHECK(cudaGetLastError());
    return VF_OK;
}

template <class I, int NDIM>
int launch_plus_dim(const PlusLaunch& L) {
    using C = PlusCfg<I, NDIM>;
    const bool ext = L.k.rnds != nullptr;
    auto kern = ext ? plus_event_kernel<I, NDIM, true, 52>
                    : (L.rng_bits == 32 ? plus_event_kernel<I, NDIM, false, 32>
                                        : plus_event_kernel<I, NDIM, false, 52>);
    // opt in to > 48 KB dynamic shared memory once per (variant, device)
    static std::atomic<uint64_t> configured[3];
    const int variant = ext ? 0 : (L.rng_bits == 32 ? 1 : 2), dev = current_device();
    if (!(configured[variant].load(std::memory_order_acquire) >> dev & 1ull)) {
        VF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)C::kSmemBytes));
        configured[variant].fetch_or(1ull << dev, std::memory_order_release);
    }
    // device-resident event count (n_events < 0): always the full grid, idle blocks exit at once
    const int blocks = L.k.n_events >= 0
                           ? grid_blocks_for(L.k.n_events, C::kThreads, 1)
                           : sm_count() * 1;
    timing_begin(L.stream);
    VF_CUDA_CHECK(launch_pdl(kern, blocks, C::kThreads, C::kSmemBytes, L.stream, L.k));
    timing_end(L.stream);
    count_launch();
    *L.nblocks_out = blocks;
    return VF_OK;
}

// Instantiates launch_event/launch_digest/launch_plus/supported_dim for integrand I.
#define VF_DIM_CASE_EVENT(D)                                                              \
    case D:                                                                               \
        if (L.mode == VF_MODE_VEGAS && L.k.train && !L.k.lim.has)                         \
            return L.rng_bits == 32 ? launch_event_dim<I, D, VF_MODE_VEGAS, 32, true>(L)  \
                                    : launch_event_dim<I, D, VF_MODE_VEGAS, 52, true>(L); \
        if (L.rng_bits == 32)                                                             \
            return L.mode == VF_MODE_VEGAS ? launch_event_dim<I, D, VF_MODE_VEGAS, 32>(L) \
                                           : launch_event_dim<I, D, VF_MODE_PLAIN, 32>(L); \
        return L.mode == VF_MODE_VEGAS ? launch_event_dim<I, D, VF_MODE_VEGAS, 52>(L)     \
                                       : launch_event_dim<I, D, VF_MODE_PLAIN, 52>(L);
#define VF_DIM_CASE_DIGEST(D)                                                             \
    case D:                                                                               \
        return L.mode == VF_MODE_VEGAS ? launch_digest_dim<I, D, VF_MODE_VEGAS>(L)        \
                                       : launch_digest_dim<I, D, VF_MODE_PLAIN>(L);
#define VF_DIM_CASE_PLUS(D) \
    case D:                 \
        return launch_plus_dim<I, D>(L);
#define VF_DIM_CASE_SUPPORTED(D) \
    case D:                      \
        return 1;

#define VF_INSTANTIATE_GENERIC_INTEGRAND(I_)                                        \
    template <> int launch_event<I_>(const EventLaunch& L) {                        \
        using I = I_;                                                               \
        switch (L.n_dim) { VF_FOREACH_DIM(VF_DIM_CASE_EVENT) default: break; }      \
        set_error("n_dim=%d has no fused instantiation", L.n_dim);                  \
        return VF_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    template <> int launch_digest<I_>(const DigestLaunch& L) {                      \
        using I = I_;                                                               \
        switch (L.n_dim) { VF_FOREACH_DIM(VF_DIM_CASE_DIGEST) default: break; }     \
        set_error("n_dim=%d has no fused instantiation", L.n_dim);                  \
        return VF_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    template <> int launch_plus<I_>(const PlusLaunch& L) {                          \
        using I = I_;                                                               \
        switch (L.n_dim) { VF_FOREACH_DIM(VF_DIM_CASE_PLUS) default: break; }       \
        set_error("n_dim=%d has no fused instantiation", L.n_dim);                  \
        return VF_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    template <> int supported_dim<I_>(int n_dim) {                                  \
        switch (n_dim) { VF_FOREACH_DIM(VF_DIM_CASE_SUPPORTED) default: break; }    \
        return 0;                                                                   \
    }

#define VF_INSTANTIATE_FIXED_INTEGRAND(I_)                                          \
    template <> int launch_event<I_>(const EventLaunch& L) {                        \
        using I = I_;                                                               \
        switch (L.n_dim) { VF_DIM_CASE_EVENT(I_::kFixedDim) default: break; }       \
        set_error("integrand needs n_dim=%d, got %d", I_::kFixedDim, L.n_dim);      \
        return VF_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    template <> int launch_digest<I_>(const DigestLaunch& L) {                      \
        using I = I_;                                                               \
        switch (L.n_dim) { VF_DIM_CASE_DIGEST(I_::kFixedDim) default: break; }      \
        set_error("integrand needs n_dim=%d, got %d", I_::kFixedDim, L.n_dim);      \
        return VF_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    template <> int launch_plus<I_>(const PlusLaunch& L) {                          \
        using I = I_;                                                               \
        switch (L.n_dim) { VF_DIM_CASE_PLUS(I_::kFixedDim) default: break; }        \
        set_error("integrand needs n_dim=%d, got %d", I_::kFixedDim, L.n_dim);      \
        return VF_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    template <> int supported_dim<I_>(int n_dim) { return n_dim == I_::kFixedDim; }

}  // namespace vf
