// Experiment harness (not product code), round 2 session 3: times the PRODUCT VEGAS+ event kernel
// (vegasflow_b200/csrc/vf_event.cuh::plus_event_kernel) standalone on a uniform sample allocation
// (n_strat^d cubes, the same event count in every cube), like scripts/exp/k1_r3.cu does for the
// event kernel.  Build-time switches: -DEXP_DIM, -DEXP_STRAT; the product-header switches of the recorded
// variants (-DVF_PLUS_THREADS, -DVF_EXP_NOPAIRS: profiles/r2_plus_variants.txt) lived there up to commit 9142b36.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I include
//        -I vegasflow_b200/csrc [-D...] scripts/exp/plus_r3.cu -o scripts/exp/plus_r3_<variant>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "vf_event.cuh"

namespace vf {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t e, const char* w) { printf("CUDA fail %s at %s\n", cudaGetErrorString(e), w); exit(1); }
void count_launch(int) {}
int sm_count() { return 148; }
void timing_begin(cudaStream_t, int) {}
void timing_end(cudaStream_t, int) {}
}
using namespace vf;

#ifndef EXP_DIM
#define EXP_DIM 8
#endif
#ifndef EXP_NAME
#define EXP_NAME "product plus kernel"
#endif
#ifndef EXP_STRAT
#define EXP_STRAT 3
#endif

int main(int argc, char** argv) {
    constexpr int d = EXP_DIM;
    const int64_t per_cube = argc > 1 ? atoll(argv[1]) : 7620;  // c3: 6561 cubes x 7620 events
    using C = PlusCfg<SymGauss, d>;
    int64_t n_cubes = 1;
    for (int j = 0; j < d; ++j) n_cubes *= EXP_STRAT;
    const int64_t n = n_cubes * per_cube;
    std::vector<double> div(d * kEdges);
    for (int j = 0; j < d; ++j) for (int b = 0; b <= kBins; ++b) {
        const double u = (double)b / kBins; div[j * kEdges + b] = 0.5 + 0.5 * (2 * u - 1) * (0.2 + 0.8 * (2 * u - 1) * (2 * u - 1));
    }
    std::vector<int32_t> n_ev(n_cubes, (int32_t)per_cube);
    std::vector<int64_t> off(n_cubes + 1);
    for (int64_t c = 0; c <= n_cubes; ++c) off[c] = c * per_cube;
    double* ddiv; cudaMalloc(&ddiv, div.size() * 8); cudaMemcpy(ddiv, div.data(), div.size() * 8, cudaMemcpyHostToDevice);
    int32_t* dn; cudaMalloc(&dn, n_cubes * 4); cudaMemcpy(dn, n_ev.data(), n_cubes * 4, cudaMemcpyHostToDevice);
    int64_t* doff; cudaMalloc(&doff, (n_cubes + 1) * 8); cudaMemcpy(doff, off.data(), (n_cubes + 1) * 8, cudaMemcpyHostToDevice);
    double *ress, *ress2; cudaMalloc(&ress, n_cubes * 8); cudaMalloc(&ress2, n_cubes * 8);
    double* ws; const size_t wsn = ws_doubles(d); cudaMalloc(&ws, wsn * 8);
    PlusKernelArgs a{};
    a.divisions = ddiv; a.partials = ws; a.n_ev = dn; a.ev_offset = doff; a.ress = ress; a.ress2 = ress2;
    a.n_cubes = n_cubes; a.n_events = n; a.n_strat = EXP_STRAT; a.rank = 0; a.world = 1;
    a.xjac = 1.0 / n_cubes; a.iteration = 1; a.train = 1; a.pk = make_philox_keys(2024);
    a.ic.p[0] = pow(1.0 / 0.1 / sqrt(M_PI), (double)d); a.ic.p[1] = (100.0 * d + 1) * (100.0 * d) / 2.0;
    auto kern = plus_event_kernel<SymGauss, d, false, 52>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    const int blocks = 148;
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemsetAsync(ws, 0, wsn * 8);
        cudaMemsetAsync(ress, 0, n_cubes * 8);
        cudaMemsetAsync(ress2, 0, n_cubes * 8);
        cudaEventRecord(e0);
        kern<<<blocks, C::kThreads, C::kSmemBytes>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    std::vector<double> h(wsn), r1(n_cubes);
    cudaMemcpy(h.data(), ws, wsn * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r1.data(), ress, n_cubes * 8, cudaMemcpyDeviceToHost);
    double s1 = 0, sh = 0;
    for (int64_t c = 0; c < n_cubes; ++c) s1 += r1[c];
    for (int i = 0; i < d * kBins; ++i) sh += h[ws_acc_offset() + i];
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-44s d=%d threads=%d regs=%3d smem=%6zu  %8.4f ms  %.4e ev/s  sum=%.15g hist=%.15g %s\n", EXP_NAME, d,
           C::kThreads, fa.numRegs, (size_t)C::kSmemBytes, best, n / (best * 1e-3), s1, sh,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    return 0;
}
