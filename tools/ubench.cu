// FP32 pipe micro-benchmark for B200: how many FADD / FFMA / packed FFMA2 / FADD2 per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define NACC 8

__device__ __forceinline__ uint64_t pk(float a, float b) { return ((uint64_t)__float_as_uint(b) << 32) | __float_as_uint(a); }

template <int KIND>
__global__ void __launch_bounds__(512) k(float* out, float s) {
    float a[NACC], b[NACC];
    uint64_t pa[NACC], pb[NACC];
    for (int i = 0; i < NACC; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = s + i; pa[i] = pk(a[i], b[i]); pb[i] = pk(b[i], a[i]); }
    const uint64_t ps = pk(s, s * 0.5f);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (KIND == 0) a[i] = a[i] + b[i];                       // FADD
            else if (KIND == 1) a[i] = fmaf(a[i], b[i], s);          // FFMA, 3 distinct regs
            else if (KIND == 2) a[i] = fmaf(a[i], 1.0009765625f, b[i]);  // FFMA imm
            else if (KIND == 3) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pa[i]) : "l"(pa[i]), "l"(pb[i]), "l"(ps));
            else if (KIND == 4) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(pa[i]) : "l"(pa[i]), "l"(pb[i]));
            else if (KIND == 5) { a[i] = a[i] + b[i]; b[i] = fmaf(b[i], s, a[i]); }   // FADD + FFMA mix
            else if (KIND == 6) a[i] = a[i] * b[i];                  // FMUL
        }
    }
    float r = 0;
    for (int i = 0; i < NACC; ++i) r += a[i] + b[i] + __uint_as_float((uint32_t)pa[i]) + __uint_as_float((uint32_t)(pa[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int KIND>
void run(const char* name, int ops_per_inst, int inst_per_iter, float* d, int nsm, double clk_ghz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = nsm * 4;
    k<KIND><<<blocks, 512>>>(d, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<KIND><<<blocks, 512>>>(d, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * 512 * ITER * NACC * inst_per_iter;
    const double per_clk_sm = inst / (ms * 1e-3) / (clk_ghz * 1e9) / nsm;
    printf("%-22s %8.3f ms  %7.1f thread-inst/clk/SM (at %.2f GHz)  %7.2f Tflop-equiv/s\n", name, ms, per_clk_sm, clk_ghz,
           inst * ops_per_inst / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz / 1e6;
    printf("%s, %d SMs, max clock %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    float* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 4 * 512 * sizeof(float));
    run<0>("FADD", 1, 1, d, p.multiProcessorCount, ghz);
    run<6>("FMUL", 1, 1, d, p.multiProcessorCount, ghz);
    run<1>("FFMA 3-reg", 2, 1, d, p.multiProcessorCount, ghz);
    run<2>("FFMA imm", 2, 1, d, p.multiProcessorCount, ghz);
    run<3>("FFMA2 (f32x2)", 4, 1, d, p.multiProcessorCount, ghz);
    run<4>("FADD2 (f32x2)", 2, 1, d, p.multiProcessorCount, ghz);
    run<5>("FADD+FFMA mix", 1, 2, d, p.multiProcessorCount, ghz);
    return 0;
}
