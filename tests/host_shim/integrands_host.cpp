// Host build of the device integrand / map source (test infrastructure, see cuda_shim.h).
#include "cuda_shim.h"

#include "vf_common.cuh"
#include "vf_integrands.cuh"

namespace vf {  // host stubs of the library plumbing declared in vf_common.cuh
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -3; }
void count_launch(int) {}
int sm_count() { return 1; }
void timing_begin(cudaStream_t, int) {}
void timing_end(cudaStream_t, int) {}
}  // namespace vf

using namespace vf;

template <class I, int NDIM>
static void eval_all(const double* x, long n, const IntegrandConsts& ic, double* out) {
    for (long i = 0; i < n; ++i) {
        double xi[NDIM];
        for (int j = 0; j < NDIM; ++j) xi[j] = x[i * NDIM + j];
        out[i] = I::template eval<NDIM>(xi, ic);
    }
}

extern "C" {
// integrand: 0 symgauss, 1 product, 2 drellyan_lo (d=4), 3 singletop_lo (d=3); pref/C for symgauss
int hs_integrand(int integrand, int n_dim, long n, const double* x, double pref, double c,
                 double* out) {
    IntegrandConsts ic{};
    ic.p[0] = pref;
    ic.p[1] = c;
#define CASE(I, D) if (n_dim == D) { eval_all<I, D>(x, n, ic, out); return 0; }
    if (integrand == 0) { CASE(SymGauss, 1) CASE(SymGauss, 2) CASE(SymGauss, 4) CASE(SymGauss, 8) CASE(SymGauss, 20) }
    if (integrand == 1) { CASE(Product, 1) CASE(Product, 3) CASE(Product, 8) }
    if (integrand == 2) { CASE(DrellYanLO, 4) }
    if (integrand == 3) { CASE(SingleTopLO, 3) }
#undef CASE
    return -1;
}
void hs_exp_nonpositive(long n, const double* x, double* out) {
    for (long i = 0; i < n; ++i) out[i] = exp_nonpositive(x[i]);
}
void hs_div_by_tenth(long n, const double* y, double* out) {
    for (long i = 0; i < n; ++i) out[i] = div_by_tenth(y[i]);
}
void hs_div_rn_by(long n, const double* y, double b, double* out) {
    const double rb = 1.0 / b;
    for (long i = 0; i < n; ++i) out[i] = div_rn_by(y[i], b, rb);
}
// the grid map of one dimension (vegas_map_dim with a single table copy)
void hs_vegas_map(long n, const double* xn, const double* divisions_row, double* x, double* wfac,
                  int* bin) {
    double2 tbl[kBins];
    for (int b = 0; b < kBins; ++b)
        tbl[b] = make_double2(divisions_row[b], __dsub_rn(divisions_row[b + 1], divisions_row[b]));
    for (long i = 0; i < n; ++i)
        vegas_map_dim<1>(xn[i], reinterpret_cast<const char*>(tbl), x[i], wfac[i], bin[i]);
}
void hs_uniforms(long n, const uint32_t* words, int rng_bits, double* out) {
    for (long i = 0; i < n; ++i)
        out[i] = rng_bits == 32 ? u32_to_uniform(words[2 * i]) : u52_to_uniform(words[2 * i], words[2 * i + 1]);
}
void hs_philox(const uint32_t* ctr, uint64_t seed, uint32_t* out) {
    const PhiloxKeys pk = make_philox_keys(seed);
    const uint4 o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], pk);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}
}
