// Float64 shared-memory FFT shared by the spectrum previews (spectrum.cu) and the transform form of the bit-faithful
// channel filter (precise.cu).
#pragma once
#include <cuda_runtime.h>

namespace iq2a {

// tw[k] = exp(-2 pi i k / n)
static __global__ void k_fft64_twiddle(double2* __restrict__ tw, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s, c;
    sincospi(-2.0 * (double)k / (double)n, &s, &c);
    tw[k] = make_double2(c, s);
}

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 mul_neg_i(double2 a) { return make_double2(a.y, -a.x); }

// In-place forward DIF FFT of length n (power of two) on NC interleaved columns: element i of column c lives at
// s[i*NC + c].  X[k] ends up at index bitrev(k).  tw holds W_T^t for t < T, T a multiple of n.
template <int NC>
__device__ void fft_dif_shared(double2* s, int n, const double2* __restrict__ tw, int tw_n) {
    int len = n;
    for (; len >= 4; len >>= 2) {
        const int q = len >> 2;
        const int qshift = 31 - __clz(q);
        const int tstep = tw_n / len;
        for (int t = threadIdx.x; t < (n >> 2) * NC; t += blockDim.x) {
            const int col = t % NC, b = t / NC;
            const int blk = b >> qshift, j = b & (q - 1);
            double2* p = s + (size_t)(blk * len + j) * NC + col;
            const double2 a0 = p[0], a1 = p[(size_t)q * NC], a2 = p[(size_t)2 * q * NC], a3 = p[(size_t)3 * q * NC];
            const double2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = mul_neg_i(csub(a1, a3));
            // two radix-2 DIF stages at once: quarters hold k = 0, 2, 1, 3 (mod 4) so the final order is bit reversal
            p[0] = cadd(b0, b2);
            p[(size_t)q * NC] = cmul(csub(b0, b2), tw[2 * j * tstep]);
            p[(size_t)2 * q * NC] = cmul(cadd(b1, b3), tw[j * tstep]);
            p[(size_t)3 * q * NC] = cmul(csub(b1, b3), tw[3 * j * tstep]);
        }
        __syncthreads();
    }
    if (len == 2) {
        for (int t = threadIdx.x; t < (n >> 1) * NC; t += blockDim.x) {
            const int col = t % NC, b = t / NC;
            double2* p = s + (size_t)(2 * b) * NC + col;
            const double2 u = p[0], v = p[NC];
            p[0] = cadd(u, v);
            p[NC] = csub(u, v);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int bitrev_n(int k, int log2n) { return (int)(__brev((unsigned)k) >> (32 - log2n)); }

}  // namespace iq2a
