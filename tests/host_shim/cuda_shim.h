// Host shim (test infrastructure): lets g++ compile the DEVICE headers
// vegasflow_b200/csrc/vf_common.cuh and vf_integrands.cuh unchanged, so that the CPU test-suite
// exercises the very source the GPU runs (map, exact divisions, exp, spinor chains) against the
// oracle.  Every CUDA intrinsic used there is restated with its IEEE meaning; build with
// -ffp-contract=off so that plain * and + stay separately rounded like -fmad=false on nvcc.
#pragma once
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>

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
struct double2 { double x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dadd_rd(double a, double b) {  // round toward -infinity
    const int old = std::fegetround();
    std::fesetround(FE_DOWNWARD);
    volatile double va = a, vb = b;
    volatile double r = va + vb;
    std::fesetround(old);
    return r;
}
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    std::memcpy(&d, &u, 8);
    return d;
}
static inline int __double2hiint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline int __double2loint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; std::memcpy(&d, &u, 8); return d; }
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }  // unused on the host
using std::acos; using std::acosh; using std::cos; using std::cosh; using std::exp; using std::fabs;
using std::fma; using std::fmax; using std::hypot; using std::log; using std::sin; using std::sinh;
using std::sqrt;
