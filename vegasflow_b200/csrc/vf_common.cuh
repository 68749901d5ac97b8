// Shared device/host definitions for the vegasflow_b200 kernels (sm_100a).
// Reference citations are file:line relative to /root/reference.
#pragma once
#ifndef VF_HOST_SHIM  // tests/host_shim builds this header with g++ to test the device source
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <utility>

#include "../../include/vegasflow_b200.h"

namespace vf {

constexpr int kBins = VF_BINS_MAX;        // configflow.py:13
constexpr int kEdges = VF_BINS_MAX + 1;   // vflow.py:239
constexpr double kFBins = 50.0;           // vflow.py:30
constexpr double kTechCut = 1e-8;         // configflow.py:16
constexpr double kAlpha = 1.5;            // configflow.py:14
constexpr double kBeta = 0.75;            // configflow.py:15
constexpr int kMaxDim = 32;               // limits arrays passed by value
constexpr double kTwo52 = 4503599627370496.0;

// ---- host-side error plumbing (vf_abi.cu) ---------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
int sm_count();
// optional CUDA-event bracket around the event kernels (vf_kernel_timing)
void timing_begin(cudaStream_t stream, int category = 0);
void timing_end(cudaStream_t stream, int category = 0);

// ---- programmatic dependent launch (sm_90+) ------------------------------------------------
// Every kernel of the iteration chain (event kernel -> reduce/exchange/refine kernel -> next
// event kernel) is launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel
// releases its successor at its very start (the successor's blocks become resident as SM
// resources free up, hiding the launch latency) and waits for the COMPLETION of its predecessor
// (griddepcontrol.wait) before touching anything the predecessor writes.  Without the launch
// attribute both instructions are no-ops.
#ifndef VF_HOST_SHIM
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
// the same with thread-block clusters of `cluster` CTAs (grid must be a multiple of it)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), int grid, int block, int cluster,
                                      cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cluster;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#endif

#define VF_CUDA_CHECK(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return vf::cuda_fail(_e, #expr); \
    } while (0)

// Integration limits, passed by value into kernel parameter space
// (monte_carlo.py:159-175).  has == 0 -> unit hypercube.
struct Limits {
    int has;
    double xdeltajac;  // prod_j xdelta[j], left to right (monte_carlo.py:171)
    double xmin[kMaxDim];
    double xdelta[kMaxDim];
};

// Constants of the built-in integrands, filled on the host.
struct IntegrandConsts {
    double p[4];
};

// ---- Philox4x32-10 ---------------------------------------------------------
// The engine's own counter-based stream (the reference delegates to
// tf.random.uniform, monte_carlo.py:264-266).  Identical to
// oracle/vegas_oracle.c::vfo_philox4x32_10.
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// Round keys k_r = key + r*(W0, W1) depend only on the seed: they are expanded once on the
// host and travel in kernel-parameter (constant) space, so the per-event rounds cost exactly
// two IMAD.WIDE + two LOP3 each.
struct PhiloxKeys {
    uint32_t k0[10];
    uint32_t k1[10];
    // 0x3FF00000, the exponent word of [1,2).  It travels as a launch parameter so that ptxas
    // cannot see a constant: (word & 0xFFFFF) | expo then stays ONE LOP3 (immediate mask +
    // register) instead of an AND and an OR with two immediates -- LOP3 holds the integer dispatch
    // port two cycles on sm_100 (profiles/r2_pipes2.txt), and there is one per uniform.
    uint32_t expo;
    uint32_t pad_;
};
constexpr uint32_t kExpoOne = 0x3FF00000u;
inline PhiloxKeys make_philox_keys(uint64_t seed) {
    PhiloxKeys pk;
    pk.expo = kExpoOne;
    pk.pad_ = 0;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        pk.k0[r] = a;
        pk.k1[r] = b;
        a += kPhiloxW0;
        b += kPhiloxW1;
    }
    return pk;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKeys& pk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ pk.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ pk.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

// ---- Philox words -> uniforms ------------------------------------------------
// The stream is defined through v = fma(m, S, 3T) with m in [1,2) the mantissa fill,
// S = 1 - 2*TECH_CUT, T = TECH_CUT: v lies in [1+T, 2-T), i.e. on the 2^-52 grid of that
// binade, and the uniform handed to the reference's seam (monte_carlo.py:264-268) is
//     r = 2 - v     in (TECH_CUT, 1-TECH_CUT],   EXACT (Sterbenz).
// Because r is a multiple of 2^-52, everything the reference does first with r is exact too:
//   vflow.py:117      1 - r = v - 1 exactly, so  xn = rn(50*(1-r)) = rn(50*v - 50) = fma(v, 50, -50)
//   vflowplus.py:72   points + r = (points+2) - v in the reals, so rn(points+r) = rn((points+2) - v)
// -- the fused kernels consume v directly and never form r: one fp64 instruction per dimension
// less than r = fma(..), 1-r, 50*(..) (8 of the 152 fp64 instructions per event at d = 8), with
// bit-identical xn.  Identical to oracle/vegas_oracle.c::u52_to_uniform / u32_to_uniform.
// (word & 0xFFFFF) | expo in one LOP3 when expo is a register (LUT 0xEA = (a & b) | c)
__device__ __forceinline__ uint32_t mantissa_hi(uint32_t word, uint32_t expo) {
#ifndef VF_HOST_SHIM
    uint32_t h;
    asm("lop3.b32 %0, %1, 0xFFFFF, %2, 0xEA;" : "=r"(h) : "r"(word), "r"(expo));
    return h;
#else
    return (word & 0xFFFFFu) | expo;
#endif
}
__device__ __forceinline__ double u52_to_v(uint32_t hi, uint32_t lo, uint32_t expo = kExpoOne) {
    // two words, 52-bit mantissa fill (like tf.random.uniform for float64)
    const double m = __hiloint2double((int)mantissa_hi(hi, expo), (int)lo);
    constexpr double S = 1.0 - 2.0 * kTechCut;
    constexpr double C = 3.0 * kTechCut;  // (1 + T) - S
    return fma(m, S, C);
}
// One 32-bit word fills the TOP 32 mantissa bits, m = 1 + k*2^-32, and the offset puts v at the
// middle of its cell: v = 1 + T + (k + 1/2)*2^-32*S.  Optional stream (rng_bits = 32): four
// uniforms per Philox block instead of two, i.e. half the integer-multiply work; resolution
// 2.3e-10 (TECH_CUT is 1e-8).  Not the default.
__device__ __forceinline__ double u32_to_v(uint32_t k) {
    const double m = __hiloint2double((int)(0x3FF00000u | (k >> 12)), (int)(k << 20));
    constexpr double S = 1.0 - 2.0 * kTechCut;
    constexpr double C = 3.0 * kTechCut + S * 1.1641532182693481e-10;  // + S*2^-33
    return fma(m, S, C);
}
__device__ __forceinline__ double u52_to_uniform(uint32_t hi, uint32_t lo) {
    return __dsub_rn(2.0, u52_to_v(hi, lo));
}
__device__ __forceinline__ double u32_to_uniform(uint32_t k) {
    return __dsub_rn(2.0, u32_to_v(k));
}

// Uniforms per Philox block and word selection for the two stream definitions.
template <int RB>
struct Rng {
    static_assert(RB == 52 || RB == 32, "rng_bits is 52 or 32");
    static constexpr int kPerCall = RB == 32 ? 4 : 2;
    // v = 2 - r of uniform h of the block (see above)
    static __device__ __forceinline__ double v(const uint4& o, int h, uint32_t expo = kExpoOne) {
        if (RB == 32) return u32_to_v(h == 0 ? o.x : (h == 1 ? o.y : (h == 2 ? o.z : o.w)));
        return h == 0 ? u52_to_v(o.x, o.y, expo) : u52_to_v(o.z, o.w, expo);
    }
    static __device__ __forceinline__ double uniform(const uint4& o, int h) {
        return __dsub_rn(2.0, v(o, h));
    }
};
__device__ __forceinline__ double rng_uniform(const uint4& o, int h, int rng_bits) {
    return rng_bits == 32 ? Rng<32>::uniform(o, h) : Rng<52>::uniform(o, h);
}

// ---- VEGAS map, one dimension ----------------------------------------------
// vflow.py:117 (xn = 50*(1-r)), :67 (ind = trunc), :70-76 (x), :78 (Delta*50).
// `tbl` points at the (x_ini, Delta) pairs of this dimension in shared memory,
// replicated TC times: tbl[bin*TC + slot]; Delta = x_fin - x_ini (:73) is
// evaluated once per bin when the table is staged -- same IEEE operation on
// the same operands as the per-event form.
template <int TC>
__device__ __forceinline__ void vegas_map_dim(double xn, const char* __restrict__ tbl_lane,
                                              double& x, double& wfac, int& bin) {
    // floor via round-down add of 2^52: t = floor(xn) + 2^52 exactly for 0 <= xn < 2^31;
    // the integer sits in the low mantissa word (== C truncation, vflow.py:67).
    const double t = __dadd_rd(xn, kTwo52);
    bin = __double2loint(t);
    const double fl = __dsub_rn(t, kTwo52);   // tf.math.floor(xn), vflow.py:75
    const double aux = __dsub_rn(xn, fl);     // :75
    // tbl_lane already includes this lane's copy slot and the dimension offset
    const double2 e = *reinterpret_cast<const double2*>(tbl_lane + bin * (TC * 16));
    x = __dadd_rn(e.x, __dmul_rn(e.y, aux));  // :76, mul then add
    wfac = __dmul_rn(e.y, kFBins);            // :78
}

#ifndef VF_HOST_SHIM
// The fused kernels address shared memory with explicit 32-bit shared-window addresses: the row
// address  bin*(TC*16) + (table base + lane slot)  is ONE integer multiply-add, the dimension
// offset is the immediate of the LDS.128, and the histogram cell of the same bin is the row address
// plus a per-lane constant when the two row pitches agree (TC*16 == HC*8).  (The pointer form
// compiled to a shift, an OR and an add per dimension; the integer dispatch port is what the
// event kernel is short of, profiles/r2_pipes2.txt.)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
template <int PITCH>
__device__ __forceinline__ uint32_t row_addr(int bin, uint32_t base) {
    // a plain C expression: ptxas schedules it better than an inline mad.lo.u32 (-1.2 % at d = 8,
    // profiles/r2_k1_pairs.txt)
    return base + (uint32_t)bin * (uint32_t)PITCH;
}
// `off` is a compile-time constant after unrolling: ptxas folds it into the address immediate
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr, uint32_t off) {
    double2 e;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e.x), "=d"(e.y) : "r"(addr + off));
    return e;
}
__device__ __forceinline__ void red_shared_f64(uint32_t addr, uint32_t off, double v) {
    asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(addr + off), "d"(v) : "memory");
}
// vegas_map_dim on a shared-window address; `row` returns bin*(TC*16) + tbl_s for the histogram,
// `off` is the byte offset of the dimension's table
template <int TC>
__device__ __forceinline__ void vegas_map_dim_s(double xn, uint32_t tbl_s, uint32_t off, double& x,
                                                double& wfac, int& bin, uint32_t& row) {
    // floor via round-down add of 2^52 (see vegas_map_dim).  Taking the floor or the truncation
    // through I2F / F2I on the XU pipe instead saves fp64 instructions but not time
    // (profiles/r2_k1_r3_floor_exp_variants.txt).
    const double t = __dadd_rd(xn, kTwo52);
    bin = __double2loint(t);
    const double fl = __dsub_rn(t, kTwo52);   // tf.math.floor(xn), vflow.py:75
    const double aux = __dsub_rn(xn, fl);     // :75
    // (forming the row address on the fp64 pipe instead -- the low word of fl*PITCH + (2^52 + base),
    // one DFMA for the IMAD -- is 2 % slower: profiles/r2_k1_r3_floor_exp_variants.txt)
    row = row_addr<TC * 16>(bin, tbl_s);
    const double2 e = lds_f64x2(row, off);
    x = __dadd_rn(e.x, __dmul_rn(e.y, aux));  // :76, mul then add
    wfac = __dmul_rn(e.y, kFBins);            // :78
}
#endif

// y / b with IEEE round-to-nearest in three fp64 operations, given rb = rn(1/b):
// q0 = rn(y*rb) is within 1 ulp of y/b, r = y - b*q0 is exact in one FMA, and q0 + r*rb rounds
// to rn(y/b) (Markstein's theorem; b's significand must not be all ones -- true for the integer
// divisors used here).  0 mismatches vs `/` on 7.6e8 host samples over 20 022 integer divisors.
__device__ __forceinline__ double div_rn_by(double y, double b, double rb) {
    const double q0 = __dmul_rn(y, rb);
    const double r = __fma_rn(-b, q0, y);
    return __fma_rn(r, rb, q0);
}

// VEGAS+ multi-GPU partition: smallest cube c in [0, n_cubes] with ev_offset[c] >= target.
// Rank r of R owns cubes [f(n*r/R), f(n*(r+1)/R)), a pure function of the offsets.
__device__ __forceinline__ int64_t first_cube_at_or_after(const int64_t* __restrict__ ev_offset,
                                                          int64_t n_cubes, int64_t target) {
    int64_t lo = 0, hi = n_cubes;  // answer in [lo, hi]
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ev_offset[mid] >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Workspace layout (doubles), see vf_workspace_bytes:
//   scalars[kMaxBlocks][2]   per-block (sum wf, sum (wf)^2), reduced later in a fixed order
//   acc[n_dim*50]            histogram accumulator: every event-kernel block adds its 50*d bin
//                            sums with native fp64 RED; the reduce kernel reads AND ZEROES it,
//                            so it is all-zero between iterations (the caller zero-initialises
//                            the workspace once)
//   scratch[2]
constexpr int kMaxBlocks = 2048;  // upper bound on event-kernel grid size
__host__ __device__ inline size_t ws_acc_offset() { return (size_t)kMaxBlocks * 2; }
__host__ __device__ inline size_t ws_doubles(int n_dim) {
    return ws_acc_offset() + (size_t)n_dim * kBins + 2;
}

}  // namespace vf
