// Experiment harness (not product code): variants of the fused event kernel for symgauss d=8,
// timed standalone.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
//   -std=c++17 -I include -I vegasflow_b200/csrc scripts/exp/k1_variants.cu -o /tmp/k1v
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "vf_common.cuh"
#include "vf_integrands.cuh"

namespace vf {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t e, const char* w) { printf("CUDA fail %s at %s\n", cudaGetErrorString(e), w); exit(1); }
void count_launch(int) {}
int sm_count() { return 148; }
}
using namespace vf;

struct Args {
    const double* divisions; double* out; uint64_t ev_begin, ev_end; double xjac; uint32_t iteration;
    PhiloxKeys pk; IntegrandConsts ic;
};

template <int NDIM, int TC, int HC, int THREADS, int MINB, bool PHILOX, bool HIST, bool MAP, bool INTEG, int EPT, bool D50>
__global__ void __launch_bounds__(THREADS, MINB) kvar(const __grid_constant__ Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int ENT = D50 ? 32 : 16;
    char* tblc = reinterpret_cast<char*>(smem_raw);
    double* hist = reinterpret_cast<double*>(smem_raw + (size_t)NDIM * kBins * TC * ENT);
    for (int i = threadIdx.x; i < NDIM * kBins * TC; i += THREADS) {
        const int jb = i / TC; const int j = jb / kBins, b = jb - j * kBins;
        const double x_ini = a.divisions[j * kEdges + b], x_fin = a.divisions[j * kEdges + b + 1];
        const double dl = __dsub_rn(x_fin, x_ini);
        double* e = reinterpret_cast<double*>(tblc + (size_t)i * ENT);
        e[0] = x_ini; e[1] = dl;
        if (D50) { e[2] = __dmul_rn(dl, 50.0); e[3] = 0.0; }
    }
    for (int i = threadIdx.x; i < NDIM * kBins * HC; i += THREADS) hist[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const char* tbl_lane = tblc + (lane % TC) * ENT;
    char* hist_lane = reinterpret_cast<char*>(hist) + (lane % HC) * 8;
    double sum = 0.0, sum2 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * THREADS * EPT;
    for (uint64_t n0 = a.ev_begin + ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * EPT; n0 < a.ev_end; n0 += stride) {
        double x[EPT][NDIM]; int bin[EPT][NDIM]; double w[EPT];
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const uint64_t n = n0 + q;
            w[q] = 1.0;
#pragma unroll
            for (int p = 0; p < (NDIM + 1) / 2; ++p) {
                uint4 o;
                if (PHILOX) o = philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)p, a.iteration, a.pk);
                else o = make_uint4((uint32_t)n * 2654435761u + p, (uint32_t)n ^ 0x9e3779b9u, (uint32_t)n * 40503u + 7 * p, (uint32_t)(n >> 3) + p);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 2 * p + h;
                    if (j < NDIM) {
                        const double r = h == 0 ? u52_to_uniform(o.x, o.y) : u52_to_uniform(o.z, o.w);
                        if (MAP) {
                            const double xn = __dmul_rn(kFBins, __dsub_rn(1.0, r));
                            const double t = __dadd_rd(xn, kTwo52);
                            const int b = __double2loint(t);
                            const double fl = __dsub_rn(t, kTwo52);
                            const double aux = __dsub_rn(xn, fl);
                            const char* ep = tbl_lane + j * (kBins * TC * ENT) + b * (TC * ENT);
                            const double2 e = *reinterpret_cast<const double2*>(ep);
                            x[q][j] = __dadd_rn(e.x, __dmul_rn(e.y, aux));
                            const double wfac = D50 ? *reinterpret_cast<const double*>(ep + 16) : __dmul_rn(e.y, kFBins);
                            w[q] = (j == 0) ? wfac : __dmul_rn(w[q], wfac);
                            bin[q][j] = b;
                        } else { x[q][j] = r; bin[q][j] = (o.x >> (h * 8)) % 50; }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            double wq = __dmul_rn(w[q], a.xjac);
            double f;
            if (INTEG) f = SymGauss::eval<NDIM>(x[q], a.ic);
            else { f = x[q][0]; for (int j = 1; j < NDIM; ++j) f += x[q][j]; }
            const double tmp = __dmul_rn(wq, f);
            const double tmp2 = __dmul_rn(tmp, tmp);
            sum += tmp; sum2 += tmp2;
            if (HIST) {
                unsigned long long* addr[NDIM]; unsigned long long old[NDIM], seen[NDIM]; unsigned long long lost = 0;
#pragma unroll
                for (int j = 0; j < NDIM; ++j) {
                    addr[j] = reinterpret_cast<unsigned long long*>(hist_lane + j * (kBins * HC * 8) + bin[q][j] * (HC * 8));
                    old[j] = *reinterpret_cast<volatile unsigned long long*>(addr[j]);
                }
#pragma unroll
                for (int j = 0; j < NDIM; ++j) {
                    const double upd = __longlong_as_double((long long)old[j]) + tmp2;
                    seen[j] = atomicCAS(addr[j], old[j], (unsigned long long)__double_as_longlong(upd));
                    lost |= seen[j] ^ old[j];
                }
                if (lost) {
#pragma unroll
                    for (int j = 0; j < NDIM; ++j) if (seen[j] != old[j]) atomicAdd(reinterpret_cast<double*>(addr[j]), tmp2);
                }
            }
        }
    }
    sum = warp_sum(sum); sum2 = warp_sum(sum2);
    __syncthreads();
    if (lane == 0) { atomicAdd(&a.out[0], sum); atomicAdd(&a.out[1], sum2); }
    for (int i = threadIdx.x; i < NDIM * kBins; i += THREADS) {
        double t = 0; for (int c = 0; c < HC; ++c) t += hist[i * HC + c];
        atomicAdd(&a.out[2 + i], t);
    }
}

template <int NDIM, int TC, int HC, int THREADS, int MINB, bool PHILOX, bool HIST, bool MAP, bool INTEG, int EPT, bool D50>
void run(const char* name, const Args& a0, int64_t n_events) {
    auto kern = kvar<NDIM, TC, HC, THREADS, MINB, PHILOX, HIST, MAP, INTEG, EPT, D50>;
    const size_t smem = (size_t)NDIM * kBins * TC * (D50 ? 32 : 16) + (size_t)NDIM * kBins * HC * 8;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
    if (occ < 1) { printf("%-44s does not fit (smem %zu)\n", name, smem); return; }
    Args a = a0; a.ev_end = a.ev_begin + n_events;
    const int blocks = 148 * occ;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemsetAsync(a.out, 0, (2 + NDIM * kBins) * 8);
        cudaEventRecord(e0);
        kern<<<blocks, THREADS, smem>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    std::vector<double> h(2 + NDIM * kBins);
    cudaMemcpy(h.data(), a.out, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-44s occ=%d regs=%3d smem=%6zu  %8.3f ms  %.3e ev/s  sum=%.6f %s\n", name, occ, fa.numRegs, smem, best,
           n_events / (best * 1e-3), h[0], err == cudaSuccess ? "" : cudaGetErrorString(err));
}

// Software-pipelined variant: the Philox blocks of the NEXT event are computed while the fp64
// work of the current event runs, so integer-multiply and fp64 pipes overlap inside one warp.
template <int NDIM, int TC, int HC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) kswp(const __grid_constant__ Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    char* tblc = reinterpret_cast<char*>(smem_raw);
    double* hist = reinterpret_cast<double*>(smem_raw + (size_t)NDIM * kBins * TC * 16);
    for (int i = threadIdx.x; i < NDIM * kBins * TC; i += THREADS) {
        const int jb = i / TC; const int j = jb / kBins, b = jb - j * kBins;
        const double x_ini = a.divisions[j * kEdges + b], x_fin = a.divisions[j * kEdges + b + 1];
        double* e = reinterpret_cast<double*>(tblc + (size_t)i * 16);
        e[0] = x_ini; e[1] = __dsub_rn(x_fin, x_ini);
    }
    for (int i = threadIdx.x; i < NDIM * kBins * HC; i += THREADS) hist[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const char* tbl_lane = tblc + (lane % TC) * 16;
    char* hist_lane = reinterpret_cast<char*>(hist) + (lane % HC) * 8;
    double sum = 0.0, sum2 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * THREADS;
    constexpr int NP = (NDIM + 1) / 2;
    uint64_t n = a.ev_begin + (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    uint4 o[NP];
    if (n < a.ev_end) {
#pragma unroll
        for (int p = 0; p < NP; ++p) o[p] = philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)p, a.iteration, a.pk);
    }
    for (; n < a.ev_end; n += stride) {
        uint4 on[NP];
        const uint64_t nn = n + stride;   // may run past the end: the extra block is simply unused
#pragma unroll
        for (int p = 0; p < NP; ++p) on[p] = philox4x32_10((uint32_t)nn, (uint32_t)(nn >> 32), (uint32_t)p, a.iteration, a.pk);
        double x[NDIM]; int bin[NDIM]; double w = 1.0;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * p + h;
                if (j < NDIM) {
                    const double r = h == 0 ? u52_to_uniform(o[p].x, o[p].y) : u52_to_uniform(o[p].z, o[p].w);
                    const double xn = __dmul_rn(kFBins, __dsub_rn(1.0, r));
                    double wfac;
                    vegas_map_dim<TC>(xn, tbl_lane + j * (kBins * TC * 16), x[j], wfac, bin[j]);
                    w = (j == 0) ? wfac : __dmul_rn(w, wfac);
                }
            }
        }
        w = __dmul_rn(w, a.xjac);
        const double f = SymGauss::eval<NDIM>(x, a.ic);
        const double tmp = __dmul_rn(w, f);
        const double tmp2 = __dmul_rn(tmp, tmp);
        sum += tmp; sum2 += tmp2;
        {
            unsigned long long* addr[NDIM]; unsigned long long old[NDIM], seen[NDIM]; unsigned long long lost = 0;
#pragma unroll
            for (int j = 0; j < NDIM; ++j) {
                addr[j] = reinterpret_cast<unsigned long long*>(hist_lane + j * (kBins * HC * 8) + bin[j] * (HC * 8));
                old[j] = *reinterpret_cast<volatile unsigned long long*>(addr[j]);
            }
#pragma unroll
            for (int j = 0; j < NDIM; ++j) {
                const double upd = __longlong_as_double((long long)old[j]) + tmp2;
                seen[j] = atomicCAS(addr[j], old[j], (unsigned long long)__double_as_longlong(upd));
                lost |= seen[j] ^ old[j];
            }
            if (lost) {
#pragma unroll
                for (int j = 0; j < NDIM; ++j) if (seen[j] != old[j]) atomicAdd(reinterpret_cast<double*>(addr[j]), tmp2);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) o[p] = on[p];
    }
    sum = warp_sum(sum); sum2 = warp_sum(sum2);
    __syncthreads();
    if (lane == 0) { atomicAdd(&a.out[0], sum); atomicAdd(&a.out[1], sum2); }
    for (int i = threadIdx.x; i < NDIM * kBins; i += THREADS) {
        double t = 0; for (int c = 0; c < HC; ++c) t += hist[i * HC + c];
        atomicAdd(&a.out[2 + i], t);
    }
}

template <int NDIM, int TC, int HC, int THREADS, int MINB>
void run_swp(const char* name, const Args& a0, int64_t n_events) {
    auto kern = kswp<NDIM, TC, HC, THREADS, MINB>;
    const size_t smem = (size_t)NDIM * kBins * TC * 16 + (size_t)NDIM * kBins * HC * 8;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
    if (occ < 1) { printf("%-44s does not fit\n", name); return; }
    Args a = a0; a.ev_end = a.ev_begin + n_events;
    const int blocks = 148 * occ;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemsetAsync(a.out, 0, (2 + NDIM * kBins) * 8);
        cudaEventRecord(e0);
        kern<<<blocks, THREADS, smem>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    std::vector<double> h(2 + NDIM * kBins);
    cudaMemcpy(h.data(), a.out, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-44s occ=%d regs=%3d spill=%zu smem=%6zu  %8.3f ms  %.3e ev/s  sum=%.6f %s\n", name, occ, fa.numRegs,
           (size_t)fa.localSizeBytes, smem, best, n_events / (best * 1e-3), h[0], err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    constexpr int D = 8;
    const int64_t N = 100000000;
    std::vector<double> div(D * kEdges);
    for (int j = 0; j < D; ++j) for (int b = 0; b < kEdges; ++b) div[j * kEdges + b] = (double)b / kBins;
    double *d_div, *d_out;
    cudaMalloc(&d_div, div.size() * 8); cudaMalloc(&d_out, (2 + D * kBins) * 8);
    cudaMemcpy(d_div, div.data(), div.size() * 8, cudaMemcpyHostToDevice);
    Args a{}; a.divisions = d_div; a.out = d_out; a.ev_begin = 0; a.xjac = 1.0 / N; a.iteration = 0;
    a.pk = make_philox_keys(2024);
    a.ic.p[0] = pow(1.0 / 0.1 / sqrt(M_PI), 8.0); a.ic.p[1] = 801.0 * 800.0 / 2.0;
    //            D  TC  HC  THR MINB PHILOX HIST  MAP   INTEG EPT D50
    run<D, 8, 16, 512, 2, true, true, true, true, 1, false>("base 512x2 TC8 HC16", a, N);
    run<D, 8, 16, 512, 2, true, false, true, true, 1, false>("no-hist", a, N);
    run<D, 8, 16, 512, 2, false, true, true, true, 1, false>("no-philox", a, N);
    run<D, 8, 16, 512, 2, false, false, true, true, 1, false>("no-philox no-hist", a, N);
    run<D, 8, 16, 512, 2, true, true, true, false, 1, false>("no-integrand", a, N);
    run<D, 8, 16, 512, 2, true, true, false, true, 1, false>("no-map", a, N);
    run<D, 8, 16, 512, 2, true, false, false, false, 1, false>("philox only", a, N);
    run<D, 8, 16, 512, 2, true, true, true, true, 1, true>("d50 table", a, N);
    run<D, 8, 32, 1024, 1, true, true, true, true, 1, false>("1024x1 TC8 HC32", a, N);
    run<D, 8, 16, 1024, 1, true, true, true, true, 1, false>("1024x1 TC8 HC16", a, N);
    run<D, 8, 16, 256, 4, true, true, true, true, 1, false>("256x4 TC8 HC16 (smem-limited)", a, N);
    run<D, 4, 8, 512, 3, true, true, true, true, 1, false>("512x3 TC4 HC8 (40 regs)", a, N);
    run<D, 4, 8, 256, 6, true, true, true, true, 1, false>("256x6 TC4 HC8", a, N);
    run<D, 8, 16, 512, 1, true, true, true, true, 2, false>("512x1 EPT2 (128 regs)", a, N);
    run<D, 8, 16, 256, 2, true, true, true, true, 2, false>("256x2 EPT2 (128 regs)", a, N);
    run<D, 8, 32, 512, 1, true, true, true, true, 2, false>("512x1 EPT2 HC32", a, N);
    run<D, 4, 4, 512, 4, true, true, true, true, 1, false>("512x4 TC4 HC4 (32 regs)", a, N);
    run_swp<D, 8, 16, 512, 2>("SWP 512x2 (64 regs)", a, N);
    run_swp<D, 8, 16, 512, 1>("SWP 512x1 (128 regs)", a, N);
    run_swp<D, 8, 16, 384, 2>("SWP 384x2 (80 regs)", a, N);
    run_swp<D, 8, 16, 768, 1>("SWP 768x1 (80 regs)", a, N);
    run_swp<D, 8, 16, 256, 3>("SWP 256x3 (80 regs)", a, N);
    run_swp<D, 8, 32, 640, 1>("SWP 640x1 HC32 (96 regs)", a, N);
    return 0;
}
