// Check of csrc/sincos_cw.cuh against CUDA's sincos() on the GPU: maximum absolute difference and the number of
// arguments whose float32-rounded sine or cosine differs, per argument range.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sincos_check tools/sincos_check.cu && tools/sincos_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../iq_to_audio_b200/csrc/sincos_cw.cuh"

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_check(double lo, double hi, uint64_t seed, unsigned long long* flips, double* maxdiff) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double u = (double)(mix64(i ^ seed) >> 11) * (1.0 / 9007199254740992.0);
    const double x = lo + (hi - lo) * u;
    double s0, c0, s1, c1;
    sincos(x, &s0, &c0);
    iq2a::sincos_cw(x, &s1, &c1);
    const double d = fmax(fabs(s0 - s1), fabs(c0 - c1));
    unsigned f = ((float)s0 != (float)s1) + ((float)c0 != (float)c1);
    for (int off = 16; off > 0; off >>= 1) f += __shfl_xor_sync(0xffffffffu, f, off);
    double m = d;
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) {
        if (f) atomicAdd(flips, (unsigned long long)f);
        atomicMax(reinterpret_cast<unsigned long long*>(maxdiff), (unsigned long long)__double_as_longlong(m));
    }
}
int main() {
    unsigned long long* flips; double* maxdiff;
    cudaMalloc(&flips, 8); cudaMalloc(&maxdiff, 8);
    const double ranges[][2] = {{-1.0, 1.0}, {-100.0, 100.0}, {-1e5, 1e5}, {-3e7, 3e7}, {2.5e7, 2.7e7}, {-4.99e7, 4.99e7}, {0.78, 0.79}, {-1e-3, 1e-3}};
    for (auto& r : ranges) {
        cudaMemset(flips, 0, 8); cudaMemset(maxdiff, 0, 8);
        const int blocks = 1 << 20;                       // 2^28 arguments
        k_check<<<blocks, 256>>>(r[0], r[1], 0x1234567ull, flips, maxdiff);
        unsigned long long f; double m;
        cudaMemcpy(&f, flips, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&m, maxdiff, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("[%.3g, %.3g]: 2^28 arguments, float32 roundings that differ: %llu, max |diff| %.3e (%s)\n", r[0], r[1], f, m, cudaGetErrorString(e));
    }
    return 0;
}
