// Experiment harness (not product code): phase timing of finalize_epilogue_kernel via clock64.
#define VF_PHASE_TIMING 1
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <vector>
#include <cmath>
#include "vf_aux.cu"
namespace vf {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t e, const char* w) { printf("CUDA fail %s at %s\n", cudaGetErrorString(e), w); exit(1); }
void count_launch(int) {}
int sm_count() { return 148; }
void timing_begin(cudaStream_t, int) {}
void timing_end(cudaStream_t, int) {}
}
using namespace vf;
int main() {
    const int d = 8, nblocks = 296;
    std::vector<double> part(ws_doubles(d));
    for (size_t i = 0; i < part.size(); ++i) part[i] = 1e-6 * (1 + (i * 2654435761u % 1000) / 1000.0) * (1 + 50.0 * exp(-0.02 * ((double)(i % 50) - 20) * ((double)(i % 50) - 20)));
    std::vector<double> div(d * 51);
    for (int j = 0; j < d; ++j) for (int b = 0; b < 51; ++b) div[j * 51 + b] = b / 50.0;
    double *dp, *dd, *dout, *dres;
    cudaMalloc(&dp, part.size() * 8); cudaMalloc(&dd, div.size() * 8); cudaMalloc(&dout, (d * 50 + 2) * 8); cudaMalloc(&dres, 16);
    cudaMemcpy(dp, part.data(), part.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemcpy(dd, div.data(), div.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dp, part.data(), part.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        launch_finalize_epilogue(dp, nblocks, d, true, 10000000, 1, dout + d * 50, dout, dd, dres, nullptr, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c[16]; cudaMemcpyFromSymbol(c, g_phase_clock, sizeof(c));
        int bad = 0; cudaMemcpyFromSymbol(&bad, g_phase_bad, sizeof(int));
        printf("rep %d: %.2f us | cycles: load+reduce %lld, refine: stage %lld, sum1 %lld, log/sqrt %lld, sum2+prefix %lld, scan %lld, interp %lld | total %lld | fallback %d\n",
               rep, ms * 1e3, c[1] - c[0], c[3] - c[2], c[4] - c[3], c[5] - c[4], c[6] - c[5], c[7] - c[6], c[8] - c[7], c[8] - c[0], bad);
    }
    return 0;
}
