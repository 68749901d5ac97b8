// Experiment harness (not product code), round 2: times the PRODUCT event kernel
// (vegasflow_b200/csrc/vf_event.cuh) standalone on 1e8 events.  The -DVF_EXP_* switches that
// selected the recorded variants (profiles/r2_k1_r3_*.txt) lived in the product headers up to commit
// "Experiments on the product event kernel: per-warp loop-exit clocks ..." and were removed after;
// the session-3 switches (VF_EXP_NP / _JC / _TC / _HC, VF_EXP_NOPAIRS, VF_EXP_EVENT_THREADS, VF_EXP_MAD_ADDR:
// profiles/r2_k1_pairs.txt, r2_k1_threads.txt) lived there up to commit 9142b36:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I include
//        -I vegasflow_b200/csrc [-DVF_EXP_...] scripts/exp/k1_r3.cu -o scripts/exp/k1_r3_<variant>
// Prints best-of-5 kernel time, events/s and checksums (sum wf, sum wf^2, sum of the histogram) so
// that variants can be compared for equality of results.
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
#define EXP_NAME "product kernel"
#endif
#ifndef EXP_INTEGRAND
#define EXP_INTEGRAND SymGauss
#endif

int main(int argc, char** argv) {
    constexpr int d = EXP_DIM;
    const int64_t n = argc > 1 ? atoll(argv[1]) : 100000000;
    using C = Cfg<EXP_INTEGRAND, d>;
    std::vector<double> div(d * kEdges);
    for (int j = 0; j < d; ++j) for (int b = 0; b <= kBins; ++b) {
        // a peaked grid like a trained symgauss one: bins concentrated around 1/2
        const double u = (double)b / kBins; div[j * kEdges + b] = 0.5 + 0.5 * (2 * u - 1) * (0.2 + 0.8 * (2 * u - 1) * (2 * u - 1));
    }
    double* ddiv; cudaMalloc(&ddiv, div.size() * 8); cudaMemcpy(ddiv, div.data(), div.size() * 8, cudaMemcpyHostToDevice);
    double* ws; const size_t wsn = ws_doubles(d); cudaMalloc(&ws, wsn * 8);
    EventKernelArgs a{};
    a.divisions = ddiv; a.partials = ws; a.ev_begin = 0; a.ev_end = (uint64_t)n; a.xjac = 1.0 / n; a.iteration = 1;
    a.train = 1; a.pk = make_philox_keys(2024);
    a.ic.p[0] = pow(1.0 / 0.1 / sqrt(M_PI), (double)d); a.ic.p[1] = (100.0 * d + 1) * (100.0 * d) / 2.0;
    auto kern = event_kernel<EXP_INTEGRAND, d, VF_MODE_VEGAS, 52, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    const int blocks = 148;
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemsetAsync(ws, 0, wsn * 8);
        cudaEventRecord(e0);
        kern<<<blocks, C::kThreads, C::kSmemBytes>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    std::vector<double> h(wsn);
    cudaMemcpy(h.data(), ws, wsn * 8, cudaMemcpyDeviceToHost);
    double s1 = 0, s2 = 0, sh = 0;
    for (int b = 0; b < blocks; ++b) { s1 += h[2 * b]; s2 += h[2 * b + 1]; }
    for (int i = 0; i < d * kBins; ++i) sh += h[ws_acc_offset() + i];
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-44s d=%d regs=%3d smem=%6zu  %8.4f ms  %.4e ev/s  sum=%.15g sum2=%.15g hist=%.15g %s\n", EXP_NAME, d,
           fa.numRegs, (size_t)C::kSmemBytes, best, n / (best * 1e-3), s1, s2, sh,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    return 0;
}
