// In-register radix-2 DIF transforms with compile-time twiddles.
//
// Every polyphase branch transform in the channel bank is an M-point FFT
// (M = 256..1024) split as R1 x R2 with R1, R2 in {16, 32}: each factor is done
// entirely in one thread's registers by the templates below, the two factors
// meet through shared memory (channelizer.cu).  Twiddles inside a factor are
// compile-time constants, so multiplications by 1, -j and (1-j)/sqrt2 cost no
// multiplies at all, and the rest become FMUL/FFMA with immediate operands.
//
// dif<R, DIR>(v): in-place, natural-order input, BIT-REVERSED output:
//   X[k] ends up in v[bitrev<R>(k)].   DIR=+1: e^{-2 pi j nk/R}, DIR=-1: e^{+...}.
#pragma once
#include <cuda_runtime.h>
#include <utility>
#include "tw_consts.cuh"

namespace iq2a {

template <typename F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
// static_for<N>(f): f(integral_constant<int,0>) ... f(integral_constant<int,N-1>)
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }
template <int R>
__host__ __device__ constexpr int bitrev(int k) {
    int r = 0;
    for (int b = 0; b < ilog2(R); ++b) r |= ((k >> b) & 1) << (ilog2(R) - 1 - b);
    return r;
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// v * W_R^T, W_R = exp(-DIR * 2 pi j / R), 0 <= T < R, R | 64.
template <int R, int T, int DIR>
__device__ __forceinline__ float2 mul_tw(float2 v) {
    static_assert(64 % R == 0, "twiddle table covers R | 64");
    constexpr int t = ((T % R) + R) % R;
    constexpr float kS = 0.70710678118654752440f;
    if constexpr (t == 0) {
        return v;
    } else if constexpr (2 * t == R) {
        return make_float2(-v.x, -v.y);
    } else if constexpr (4 * t == R) {          // -j (forward) / +j (inverse)
        return DIR > 0 ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
    } else if constexpr (4 * t == 3 * R) {
        return DIR > 0 ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
    } else if constexpr (8 * t == R) {          // (1 - j)/sqrt2 forward
        return DIR > 0 ? make_float2((v.x + v.y) * kS, (v.y - v.x) * kS)
                       : make_float2((v.x - v.y) * kS, (v.y + v.x) * kS);
    } else if constexpr (8 * t == 3 * R) {      // (-1 - j)/sqrt2 forward
        return DIR > 0 ? make_float2((v.y - v.x) * kS, -(v.x + v.y) * kS)
                       : make_float2(-(v.x + v.y) * kS, (v.x - v.y) * kS);
    } else {
        constexpr float c = kCos64[t * (64 / R)];
        constexpr float s = kSin64[t * (64 / R)];
        // forward: (c - j s) ; inverse: (c + j s)
        return DIR > 0 ? make_float2(v.x * c + v.y * s, v.y * c - v.x * s)
                       : make_float2(v.x * c - v.y * s, v.y * c + v.x * s);
    }
}

template <int N, int OFF, int DIR, int RTOT>
__device__ __forceinline__ void dif_rec(float2 (&v)[RTOT]) {
    if constexpr (N >= 2) {
        constexpr int H = N / 2;
        static_for<H>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const float2 a = v[OFF + j], b = v[OFF + j + H];
            v[OFF + j] = cadd(a, b);
            v[OFF + j + H] = mul_tw<N, j, DIR>(csub(a, b));
        });
        dif_rec<H, OFF, DIR, RTOT>(v);
        dif_rec<H, OFF + H, DIR, RTOT>(v);
    }
}

template <int R, int DIR>
__device__ __forceinline__ void dif(float2 (&v)[R]) {
    dif_rec<R, 0, DIR, R>(v);
}

}  // namespace iq2a
