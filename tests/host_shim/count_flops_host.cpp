// Test infrastructure: fp64 operation count of the DEVICE integrand source.
// vf_integrands.cuh is compiled unchanged with `double` replaced by a counting scalar, as
// SURVEY.md 8(d) prescribes for the matrix elements ("count by instantiating the shared integrand
// header with an op-counting scalar type on the host"): add/sub/mul/div = 1, fused multiply-add
// = 2, sqrt and every transcendental = 1 (sincos = 2); comparisons, selects, negation, fabs and
// bit casts = 0.  The count follows the path each event takes, so it is an average over events.
#include <cfenv>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <utility>

#include "../../include/vegasflow_b200.h"

static long g_ops = 0;
struct cd {
    double v;
    constexpr cd(double x = 0.0) : v(x) {}
    constexpr cd(int x) : v(x) {}
    explicit operator int() const { return (int)v; }
};
// compile-time constant expressions (constexpr double k = 1.0 - 2.0 * T) are not operations
#define VF_COUNT(n) do { if (!__builtin_is_constant_evaluated()) g_ops += (n); } while (0)
constexpr cd operator+(cd a, cd b) { VF_COUNT(1); return a.v + b.v; }
constexpr cd operator-(cd a, cd b) { VF_COUNT(1); return a.v - b.v; }
constexpr cd operator*(cd a, cd b) { VF_COUNT(1); return a.v * b.v; }
constexpr cd operator/(cd a, cd b) { VF_COUNT(1); return a.v / b.v; }
constexpr cd operator-(cd a) { return -a.v; }
static inline cd& operator+=(cd& a, cd b) { a = a + b; return a; }
static inline bool operator<(cd a, cd b) { return a.v < b.v; }
static inline bool operator>(cd a, cd b) { return a.v > b.v; }
static inline bool operator<=(cd a, cd b) { return a.v <= b.v; }
static inline bool operator>=(cd a, cd b) { return a.v >= b.v; }
static inline bool operator==(cd a, cd b) { return a.v == b.v; }
static inline bool operator!=(cd a, cd b) { return a.v != b.v; }
static inline cd sqrt(cd a) { ++g_ops; return std::sqrt(a.v); }
static inline cd log(cd a) { ++g_ops; return std::log(a.v); }
static inline cd exp(cd a) { ++g_ops; return std::exp(a.v); }
static inline cd sin(cd a) { ++g_ops; return std::sin(a.v); }
static inline cd cos(cd a) { ++g_ops; return std::cos(a.v); }
static inline cd acos(cd a) { ++g_ops; return std::acos(a.v); }
static inline void sincos(cd a, cd* s, cd* c) { g_ops += 2; *s = std::sin(a.v); *c = std::cos(a.v); }
static inline cd fma(cd a, cd b, cd c) { g_ops += 2; return std::fma(a.v, b.v, c.v); }
static inline cd fabs(cd a) { return std::fabs(a.v); }

#define VF_HOST_SHIM 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __constant__
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
typedef int cudaError_t;
typedef void* cudaStream_t;
static const int cudaSuccess = 0;
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
struct double2 { cd x, y; };
static inline double2 make_double2(cd x, cd y) { return {x, y}; }
static inline cd __dadd_rn(cd a, cd b) { return a + b; }
static inline cd __dsub_rn(cd a, cd b) { return a - b; }
static inline cd __dmul_rn(cd a, cd b) { return a * b; }
static inline cd __ddiv_rn(cd a, cd b) { return a / b; }
static inline cd __fma_rn(cd a, cd b, cd c) { return fma(a, b, c); }
static inline cd __dadd_rd(cd a, cd b) { ++g_ops; return std::floor(a.v) + b.v; }  // only used as the floor add
static inline cd __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    std::memcpy(&d, &u, 8);
    return d;
}
static inline int __double2hiint(cd d) { uint64_t u; std::memcpy(&u, &d.v, 8); return (int)(u >> 32); }
static inline cd __shfl_xor_sync(unsigned, cd v, int) { return v; }  // unused on the host
static inline int __double2loint(cd d) { uint64_t u; std::memcpy(&u, &d.v, 8); return (int)(uint32_t)u; }

#define double cd
#include "vf_common.cuh"
#include "vf_integrands.cuh"
#undef double

namespace vf {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -3; }
void count_launch(int) {}
int sm_count() { return 1; }
void timing_begin(cudaStream_t, int) {}
void timing_end(cudaStream_t, int) {}
}  // namespace vf

template <class I, int NDIM>
static double count_avg(const double* x, long n, double pref, double c) {
    vf::IntegrandConsts ic{};
    ic.p[0] = pref;
    ic.p[1] = c;
    g_ops = 0;
    for (long i = 0; i < n; ++i) {
        cd xi[NDIM];
        for (int j = 0; j < NDIM; ++j) xi[j] = x[i * NDIM + j];
        volatile double sink = I::template eval<NDIM>(xi, ic).v;
        (void)sink;
    }
    return (double)g_ops / (double)n;
}

extern "C" double hs_count_flops(int integrand, int n_dim, long n, const double* x, double pref,
                                 double c) {
#define CASE(I, D) if (n_dim == D) return count_avg<vf::I, D>(x, n, pref, c);
    if (integrand == 0) { CASE(SymGauss, 4) CASE(SymGauss, 8) CASE(SymGauss, 20) }
    if (integrand == 1) { CASE(Product, 8) }
    if (integrand == 2) { CASE(DrellYanLO, 4) }
    if (integrand == 3) { CASE(SingleTopLO, 3) }
#undef CASE
    return -1.0;
}
