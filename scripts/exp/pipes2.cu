// Microbenchmark (not product code), round 2: which sm_100 pipes overlap?  Every instruction class
// is pinned with inline PTX (checked in SASS: cuobjdump -sass scripts/exp/pipes2), 8 independent
// dependency chains per thread, one 1024-thread block per SM (8 warps per scheduler, the event
// kernel's shape) or two (16 warps).  Reported: cycles per scheduler per "group" = one warp
// instruction of every enabled class.  If two classes overlap perfectly, the pair costs the max of
// the two; if they share an issue/dispatch resource, the sum.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/exp/pipes2.cu -o scripts/exp/pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum : unsigned {
    DFMA = 1, IMADW = 2, LOP = 4, IADD = 8, FFMA = 16, I2F = 32, LDS = 64, IMADLO = 128, DADD = 256,
    F2I = 512
};

template <unsigned M>
__global__ void __launch_bounds__(1024) k(int iters, uint32_t seed, double* sink) {
    __shared__ double sh[1024];
    sh[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a[8], e[8];
    uint64_t p[8];
    uint32_t l[8], s[8], q[8];
    float f[8];
    double g[8];
    int gi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = 1.0 + threadIdx.x + i; e[i] = 0.5 + i; p[i] = seed + threadIdx.x * 8 + i;
        l[i] = seed ^ (threadIdx.x + i); s[i] = seed + i + threadIdx.x * 77; q[i] = seed * (i + 3); f[i] = 1.0f + i;
        g[i] = i; gi[i] = threadIdx.x + i;
    }
    const double dm = 1.0000001, dc = 1e-9;
    const uint32_t M0 = 0xD2511F53u, K = 0x9E3779B9u;
    const float fm = 1.0000001f, fc = 1e-9f;
    const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(sh) + (threadIdx.x & 31) * 8;
    double ld_acc = 0.0;
#pragma unroll 16
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (M & DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(dm), "d"(dc));
            if (M & IMADW) {
                // Philox-shaped: both product words are consumed (ptxas narrows a mul.wide whose high
                // word is dead), so this class is IMAD.WIDE + one LOP3
                uint32_t lo, hi, x;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p[i]));
                asm volatile("xor.b32 %0, %1, %2;" : "=r"(x) : "r"(lo), "r"(hi));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p[i]) : "r"(x), "r"(M0));
            }
            if (M & LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(K), "r"(seed));
            if (M & IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(s[(i + 3) & 7]));
            if (M & FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fm), "f"(fc));
            if (M & I2F) {
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(g[i]) : "r"(gi[i]));
                uint32_t lo, hi;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(g[i]));
                gi[i] = (int)hi;
            }
            if (M & F2I) {
                asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(gi[i]) : "d"(g[i]));
                g[i] = __hiloint2double(0x40390000, gi[i]);
            }
            if (M & LDS) {
                double v;
                asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(sh_base + (uint32_t)i * 256));
                ld_acc += 0;  // keep the loads independent of the fp64 pipe
                e[i] = v;
            }
            if (M & IMADLO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(M0), "r"(K));
            if (M & DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(e[i]) : "d"(dc));
        }
    }
    double sd = ld_acc;
    uint32_t su = 0;
    float sf = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sd += a[i] + e[i] + g[i]; su ^= (uint32_t)p[i] ^ (uint32_t)(p[i] >> 32) ^ l[i] ^ s[i] ^ q[i] ^ (uint32_t)gi[i];
        sf += f[i];
    }
    if (sd == 123.456 || su == 0x12345678u || sf == 7.5f) sink[0] = sd + su + sf;
}

static float g_mhz = 1965.f;

template <unsigned M>
void run(const char* name, int n_classes, double* sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 3200;
    for (int blocks_per_sm = 1; blocks_per_sm <= 2; ++blocks_per_sm) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            k<M><<<148 * blocks_per_sm, 1024>>>(iters, 12345u, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        // per scheduler: 8*blocks_per_sm warps, each issuing iters*8 groups
        const double groups = 8.0 * blocks_per_sm * iters * 8.0;
        const double cycles = best * 1e-3 * g_mhz * 1e6;
        printf("%-34s %2d warps/sched  %8.3f ms  %6.2f cycles/group (%d instr)\n", name, 8 * blocks_per_sm, best,
               cycles / groups, n_classes);
    }
}

int main() {
    double* sink; cudaMalloc(&sink, 8);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    if (khz > 0) g_mhz = khz / 1000.f;
    printf("# clock used for the cycle conversion: %.0f MHz (max SM clock; check the clocks line of the run)\n", g_mhz);
    run<DFMA>("DFMA", 1, sink);
    run<DADD>("DADD", 1, sink);
    run<IMADW>("IMAD.WIDE+LOP3 (Philox-shaped)", 2, sink);
    run<IMADLO>("IMAD (lo)", 1, sink);
    run<LOP>("LOP3", 1, sink);
    run<IADD>("IADD", 1, sink);
    run<FFMA>("FFMA", 1, sink);
    run<I2F>("I2F.F64.S32", 1, sink);
    run<F2I>("F2I.S32.F64", 1, sink);
    run<LDS>("LDS.64", 1, sink);
    run<DFMA | DADD>("DFMA + DADD", 2, sink);
    run<DFMA | IMADW>("DFMA + IMAD.WIDE", 2, sink);
    run<DFMA | IMADLO>("DFMA + IMAD(lo)", 2, sink);
    run<DFMA | LOP>("DFMA + LOP3", 2, sink);
    run<DFMA | IADD>("DFMA + IADD", 2, sink);
    run<DFMA | FFMA>("DFMA + FFMA", 2, sink);
    run<DFMA | I2F>("DFMA + I2F", 2, sink);
    run<DFMA | LDS>("DFMA + LDS", 2, sink);
    run<IMADW | LOP>("IMAD.WIDE + LOP3", 2, sink);
    run<IMADW | FFMA>("IMAD.WIDE + FFMA", 2, sink);
    run<IMADW | LDS>("IMAD.WIDE + LDS", 2, sink);
    run<LOP | IADD>("LOP3 + IADD", 2, sink);
    run<LOP | FFMA>("LOP3 + FFMA", 2, sink);
    run<DFMA | LOP | LOP * 0 | IADD>("DFMA + LOP3 + IADD", 3, sink);
    run<DFMA | IMADW | LOP>("DFMA + IMAD.WIDE + LOP3", 3, sink);
    run<DFMA | IMADW | LOP | LDS>("DFMA + IMAD.WIDE + LOP3 + LDS", 4, sink);
    // the event kernel's mix per 16 groups would be ~ 152 fp64 : 68 IMAD.WIDE : 98 LOP3; closest
    // integer mix below: 2 fp64 (DFMA + DADD) : 1 IMAD.WIDE : 1 LOP3
    run<DFMA | DADD | IMADW | LOP>("DFMA + DADD + IMAD.WIDE + LOP3", 4, sink);
    return 0;
}
