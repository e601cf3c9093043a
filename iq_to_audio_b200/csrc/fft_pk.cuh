// Packed float32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2) and an in-register DIF
// transform that runs TWO independent transforms per thread: every register pair holds the
// same element of transform A (low half) and transform B (high half), real and imaginary
// parts in separate pairs.  All butterflies, including the "free" multiplications by -j
// (operand swap + reversed subtraction), become packed instructions -- half the issue slots
// of the scalar version for the same flops (tools/ubench.cu: FFMA2 = 2 x FFMA per slot).
#pragma once
#include "fft_regs.cuh"

namespace iq2a {

typedef unsigned long long pk_t;

__device__ __forceinline__ pk_t pk_make(float lo, float hi) {
    pk_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void pk_split(pk_t a, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
__device__ __forceinline__ pk_t pk_add(pk_t a, pk_t b) {
    pk_t r;
    asm("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk_t pk_sub(pk_t a, pk_t b) {
    pk_t r;
    asm("sub.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk_t pk_mul(pk_t a, pk_t b) {
    pk_t r;
    asm("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk_t pk_fma(pk_t a, pk_t b, pk_t c) {   // a*b + c
    pk_t r;
    asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ pk_t pk_bc(float x) { return pk_make(x, x); }

// One radix-2 DIF butterfly on packed pairs: a' = a + b ; b' = (a - b) * W_N^T (forward).
template <int N, int T>
__device__ __forceinline__ void pk_bfly(pk_t& are, pk_t& aim, pk_t& bre, pk_t& bim) {
    static_assert(64 % N == 0 && T >= 0 && 2 * T < N, "twiddle range");
    constexpr float kS = 0.70710678118654752440f;
    const pk_t sre = pk_add(are, bre), sim = pk_add(aim, bim);
    if constexpr (T == 0) {
        const pk_t dre = pk_sub(are, bre), dim = pk_sub(aim, bim);
        bre = dre;
        bim = dim;
    } else if constexpr (4 * T == N) {            // * (-j): (x + jy) -> (y - jx)
        const pk_t nre = pk_sub(aim, bim);        //  re' =  (a-b).im
        const pk_t nim = pk_sub(bre, are);        //  im' = -(a-b).re
        bre = nre;
        bim = nim;
    } else if constexpr (8 * T == N) {            // * (1-j)/sqrt2
        const pk_t dre = pk_sub(are, bre), dim = pk_sub(aim, bim);
        bre = pk_mul(pk_add(dre, dim), pk_bc(kS));
        bim = pk_mul(pk_sub(dim, dre), pk_bc(kS));
    } else if constexpr (8 * T == 3 * N) {        // * (-1-j)/sqrt2
        const pk_t dre = pk_sub(are, bre), dim = pk_sub(aim, bim);
        bre = pk_mul(pk_sub(dim, dre), pk_bc(kS));
        bim = pk_mul(pk_add(dre, dim), pk_bc(-kS));
    } else {                                      // * (c - js)
        constexpr float c = kCos64[T * (64 / N)];
        constexpr float s = kSin64[T * (64 / N)];
        const pk_t dre = pk_sub(are, bre), dim = pk_sub(aim, bim);
        bre = pk_fma(dim, pk_bc(s), pk_mul(dre, pk_bc(c)));
        bim = pk_fma(dre, pk_bc(-s), pk_mul(dim, pk_bc(c)));
    }
    are = sre;
    aim = sim;
}

template <int N, int OFF, int RTOT>
__device__ __forceinline__ void pk_dif_rec(pk_t (&re)[RTOT], pk_t (&im)[RTOT]) {
    if constexpr (N >= 2) {
        constexpr int H = N / 2;
        static_for<H>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            pk_bfly<N, j>(re[OFF + j], im[OFF + j], re[OFF + j + H], im[OFF + j + H]);
        });
        pk_dif_rec<H, OFF, RTOT>(re, im);
        pk_dif_rec<H, OFF + H, RTOT>(re, im);
    }
}

// forward R-point DIF of two packed transforms; X[k] ends up in slot bitrev<R>(k)
template <int R>
__device__ __forceinline__ void pk_dif(pk_t (&re)[R], pk_t (&im)[R]) {
    pk_dif_rec<R, 0, R>(re, im);
}

}  // namespace iq2a
