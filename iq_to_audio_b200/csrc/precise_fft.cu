// Bit-faithful channel filter, fast form: float64 overlap-save with register-resident transform passes, plus a
// repair pass that keeps the result identical to the direct form.
//
// precise.cu's direct form (k_fir_decim_f64) spends 2 * ntaps DFMAs per output sample (65 k for the 32 769-tap SSB
// filter): 11 ms per 10 s of a 20 MS/s capture with two such channels.  The transform form needs ~20x fewer
// operations but (a) its shared-memory radix-4 passes made it memory-bound (five round trips of 16-byte elements per
// transform) and (b) about one sample in 10^4 lands on the neighbouring float32, because the polyphase sum cancels
// ~80 dB of out-of-band signal and the float64 rounding error of the transforms is no longer negligible against a
// float32 half-ulp.  This file fixes both:
//   k_fir_fft64r   M = 1024 = 32 x 32: each 1024-point transform is two passes of 32-point DIFs held in registers
//                  (two shared-memory round trips), 4 branch columns per tile and 128 threads so that two CTAs share
//                  an SM (one loads while the other transforms), branch spectra H in row order so the
//                  multiply-accumulate reads it coalesced.  Every output whose float64 value lies within `tol` of the
//                  midpoint between two float32 values is appended to a repair list.
//   k_fir_repair   one CTA per listed sample: the float64 direct form over all taps (the sum precise.cu computes),
//                  rounded once.  4 tol / ulp(|s|) of the samples are listed.
// tol(v) = tol_abs + tol_rel |v| = 1e-15 + 64 eps |v|.  Measured (tools/precise_tol_sweep.py, 20 s at int16 full scale,
// 3.8 M samples per channel): with NO repair the transform form adds ONE differing sample to the three that two
// float64 direct forms with different summation orders already disagree on (values within ~1e-17 of a rounding
// boundary, which no float64 evaluation -- the reference's included -- can decide), i.e. its error in a channel 60 dB
// below the wideband level is ~1e-17, 100x below tol_abs; the relative term covers strong channels, where the
// transform's error follows the signal (a few eps |v|).  Round 2 started with a flat 3e-14 (0.8 % of the rows of a
// quiet channel listed, 2.7 of 12 ms on cfg3's shape).
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "fft64.cuh"
#include "precise.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

namespace {

constexpr int kM = 1024;
constexpr int kCols = 4;                   // branch columns per tile
constexpr int kRS = kCols;                 // tile row stride (double2); bank conflicts are avoided by the swizzle of tix()
constexpr int kThreads = 128;              // kCols x 32 row groups; 248 registers each -> two CTAs per SM
constexpr int kBins = kM / kThreads;       // spectrum rows per thread in the multiply-accumulate
constexpr size_t kTileBytes = (size_t)kM * kRS * sizeof(double2);
constexpr size_t kSmem = kTileBytes + (size_t)kM * sizeof(double2);

template <typename F, int... I>
__device__ __forceinline__ void sfor_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void sfor(F&& f) {
    sfor_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}
__host__ __device__ constexpr int rev5(int k) {
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// Index of element (row, column) of the tile.  A row is four double2 (64 bytes), two rows share a 128-byte line of
// eight 16-byte units, and the unit inside the line is XOR-swizzled with ((row >> 1) & 3) | (((row >> 5) & 1) << 2),
// so that all three access patterns of a quarter-warp (8 threads x 16 bytes = one wavefront) touch eight different
// units: pass-1 stores (rows g, g+1 x 4 columns), pass-2 loads / stores (rows 32 g + i, 32 (g+1) + i x 4 columns) and
// the multiply-accumulate's reads (8 consecutive rows, one column).  The padded layout this replaces (row stride 5)
// served only the last one; the two passes took two wavefronts per quarter-warp (profiles: L1 data pipe 62 %).
__device__ __forceinline__ int tix(int row, int c) {
    return ((row << 2) | c) ^ (((row >> 1) & 3) | (((row >> 5) & 1) << 2));
}

// radix-2 DIF of N points held in v[OFF .. OFF+N), twiddles W_N^j = tw[j * (1024 / N)]; X[k] ends in slot rev(k)
template <int N, int OFF>
__device__ __forceinline__ void ddif(double2 (&v)[32], const double2* __restrict__ tw) {
    if constexpr (N >= 2) {
        constexpr int H = N / 2;
        sfor<H>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const double2 a = v[OFF + j], b = v[OFF + j + H];
            v[OFF + j] = make_double2(a.x + b.x, a.y + b.y);
            const double dx = a.x - b.x, dy = a.y - b.y;
            if constexpr (j == 0) {
                v[OFF + j + H] = make_double2(dx, dy);
            } else if constexpr (4 * j == N) {                    // * (-i)
                v[OFF + j + H] = make_double2(dy, -dx);
            } else {
                const double2 w = tw[j * (kM / N)];
                v[OFF + j + H] = make_double2(fma(dx, w.x, -dy * w.y), fma(dx, w.y, dy * w.x));
            }
        });
        ddif<H, OFF>(v, tw);
        ddif<H, OFF + H>(v, tw);
    }
}

// In-place forward 1024-point transforms of `ncol` columns of the tile (element n of column c at s[tix(n, c)]).
// X[k] ends up in row (k & 31) * 32 + (k >> 5).
// `load(i)` supplies element n = g + 32 i of this thread's column (c, g = thread & 3, thread / 4): from the tile itself
// (the inverse transform) or straight from global memory (the forward tiles: no staging pass through shared memory).
template <class Load>
__device__ __forceinline__ void fft1024_tile(double2* s, const double2* __restrict__ tw, int ncol, Load&& load) {
    const int c = threadIdx.x & (kCols - 1), g = threadIdx.x / kCols;        // g: 0..31
    double2 v[32];
    if (c < ncol) {
        // pass 1: n = m2 + 32 m1, 32-point DIF over m1 for m2 = g, then W_1024^{m2 k1}; result k1 -> row k1*32 + m2
        sfor<32>([&](auto ic) { constexpr int i = decltype(ic)::value; v[i] = load(i); });
        ddif<32, 0>(v, tw);
    }
    __syncthreads();                          // every thread is done with the tile's previous contents
    if (c < ncol) {
        sfor<32>([&](auto kc) {
            constexpr int k1 = decltype(kc)::value;
            double2 x = v[rev5(k1)];
            if constexpr (k1 != 0) {
                const double2 w = tw[(g * k1) & (kM - 1)];
                x = make_double2(fma(x.x, w.x, -x.y * w.y), fma(x.x, w.y, x.y * w.x));
            }
            s[tix(k1 * 32 + g, c)] = x;
        });
    }
    __syncthreads();
    if (c < ncol) {
        // pass 2: for k1 = g, 32-point DIF over m2; X[k1 + 32 k2] -> row k1*32 + k2
        sfor<32>([&](auto ic) { constexpr int i = decltype(ic)::value; v[i] = s[tix(g * 32 + i, c)]; });
        ddif<32, 0>(v, tw);
        sfor<32>([&](auto kc) { constexpr int k2 = decltype(kc)::value; s[tix(g * 32 + k2, c)] = v[rev5(k2)]; });
    }
    __syncthreads();
}
__device__ __forceinline__ void fft1024_tile(double2* s, const double2* __restrict__ tw, int ncol) {
    const int c = threadIdx.x & (kCols - 1), g = threadIdx.x / kCols;
    fft1024_tile(s, tw, ncol, [&](int i) { return s[tix(g + 32 * i, c)]; });
}

__global__ void k_fir_fftr_build(const double* __restrict__ taps, int ntaps, int D, int Q,
                                 const double2* __restrict__ tw, double2* __restrict__ H) {
    // H[p][rho]: spectrum / M of branch p at the bin that the tile keeps in row rho
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * kM) return;
    const int p = (int)(i / kM), rho = (int)(i - (int64_t)p * kM);
    const int k = (rho >> 5) + 32 * (rho & 31);
    double re = 0.0, im = 0.0;
    for (int q = 0; q <= Q; ++q) {
        const int64_t t = (int64_t)q * D - p;
        if (t < 0 || t >= ntaps) continue;
        const double2 w = tw[(int)(((int64_t)k * q) & (kM - 1))];
        re = fma(taps[t], w.x, re);
        im = fma(taps[t], w.y, im);
    }
    H[i] = make_double2(re / kM, im / kM);
}

// distance of v from the midpoint between the two float32 values that bracket it
__device__ __forceinline__ double boundary_distance(double v) {
    const float f = (float)v;
    const float g = nextafterf(f, v > (double)f ? INFINITY : -INFINITY);
    const double mid = 0.5 * ((double)f + (double)g);
    return fabs(v - mid);
}

__global__ void __launch_bounds__(kThreads, 2)
k_fir_fft64r(const float2* __restrict__ mixed, const double2* __restrict__ H, const double2* __restrict__ tw_g, int D, int Q,
             int64_t nrows, float2* __restrict__ out, double tol, double tol_rel, int* __restrict__ risky, int* __restrict__ n_risky,
             int risky_cap) {
    extern __shared__ __align__(16) unsigned char sm[];
    double2* s = reinterpret_cast<double2*>(sm);                          // [kM][kRS]
    double2* tw = reinterpret_cast<double2*>(sm + kTileBytes);            // [kM]
    for (int i = threadIdx.x; i < kM; i += kThreads) tw[i] = tw_g[i];
    const int ld = kM - Q;
    const int64_t row0 = (int64_t)blockIdx.x * ld;                         // first `mixed` row of this block
    const int64_t src_rows = nrows + Q;
    double2 acc[kBins];
#pragma unroll
    for (int i = 0; i < kBins; ++i) acc[i] = make_double2(0.0, 0.0);
    __syncthreads();
    const int tc = threadIdx.x & (kCols - 1), tg = threadIdx.x / kCols;
    for (int p0 = 0; p0 < D; p0 += kCols) {
        const int ncol = min(kCols, D - p0);
        // the 32 rows this thread transforms in pass 1 come straight from global memory (the four columns of a row
        // are one 32-byte sector); rows past the end of the signal are zero
        const float2* __restrict__ src = mixed + (row0 + tg) * (int64_t)D + p0 + tc;
        fft1024_tile(s, tw, ncol, [&](int i) {
            float2 f = make_float2(0.f, 0.f);
            if (row0 + tg + 32 * i < src_rows) f = __ldg(src + (int64_t)(32 * i) * D);
            return make_double2((double)f.x, (double)f.y);
        });
#pragma unroll
        for (int i = 0; i < kBins; ++i) {
            const int rho = threadIdx.x + i * kThreads;
            const double2* hs = H + (size_t)p0 * kM + rho;
            for (int c = 0; c < ncol; ++c) {
                const double2 h = hs[(size_t)c * kM], x = s[tix(rho, c)];
                acc[i].x = fma(h.x, x.x, fma(-h.y, x.y, acc[i].x));
                acc[i].y = fma(h.x, x.y, fma(h.y, x.x, acc[i].y));
            }
        }
        __syncthreads();
    }
    // inverse transform: IFFT(Y) = conj(FFT(conj(Y))) (1/M is in H); Y[k] goes to row k (natural order) of column 0
#pragma unroll
    for (int i = 0; i < kBins; ++i) {
        const int rho = threadIdx.x + i * kThreads;
        const int k = (rho >> 5) + 32 * (rho & 31);
        s[tix(k, 0)] = make_double2(acc[i].x, -acc[i].y);
    }
    __syncthreads();
    fft1024_tile(s, tw, 1);
    for (int j = Q + threadIdx.x; j < kM; j += kThreads) {
        const int64_t m = row0 + (j - Q);
        if (m < nrows) {
            const double2 v = s[tix((j & 31) * 32 + (j >> 5), 0)];
            const double re = v.x, im = -v.y;
            out[m] = make_float2((float)re, (float)im);
            if (boundary_distance(re) < fma(fabs(re), tol_rel, tol) || boundary_distance(im) < fma(fabs(im), tol_rel, tol)) {
                const int at = atomicAdd(n_risky, 1);
                if (at < risky_cap) risky[at] = (int)m;
            }
        }
    }
}

// float64 direct form for the listed rows: out[m] = c64( sum_k h[k] mixed[(m + Q) D - k] ), `mixed` row 0 = output row -Q.
// A CTA of 1024 threads takes kRepairBatch listed rows at a time (32 taps per thread and row): one tap load feeds
// that many independent multiply-add chains and that many sample loads are in flight per thread -- a single row per
// 256-thread CTA was bound by the latency of its own 128 dependent iterations.
constexpr int kRepairBatch = 4;
constexpr int kRepairThreads = 1024;

__global__ void __launch_bounds__(kRepairThreads) k_fir_repair(const float2* __restrict__ mixed, const double* __restrict__ taps, int ntaps,
                                                    int D, int Q, const int* __restrict__ risky, int* __restrict__ n_risky,
                                                    int risky_cap, float2* __restrict__ out) {
    __shared__ double2 part[kRepairBatch][kRepairThreads / 32];
    const int total = min(*n_risky, risky_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_risky + 1, total);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = blockIdx.x * kRepairBatch; it < total; it += gridDim.x * kRepairBatch) {
        const float2* top[kRepairBatch];                          // the sample that meets h[0]
        int64_t mrow[kRepairBatch];
#pragma unroll
        for (int j = 0; j < kRepairBatch; ++j) {
            mrow[j] = risky[min(it + j, total - 1)];
            top[j] = mixed + (mrow[j] + Q) * (int64_t)D;
        }
        double ar[kRepairBatch], ai[kRepairBatch];
#pragma unroll
        for (int j = 0; j < kRepairBatch; ++j) ar[j] = ai[j] = 0.0;
        for (int k = threadIdx.x; k < ntaps; k += kRepairThreads) {
            const double h = __ldg(taps + k);
#pragma unroll
            for (int j = 0; j < kRepairBatch; ++j) {
                const float2 x = __ldg(top[j] - k);
                ar[j] = fma(h, (double)x.x, ar[j]);
                ai[j] = fma(h, (double)x.y, ai[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < kRepairBatch; ++j) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ar[j] += __shfl_xor_sync(0xffffffffu, ar[j], off);
                ai[j] += __shfl_xor_sync(0xffffffffu, ai[j], off);
            }
            if (lane == 0) part[j][warp] = make_double2(ar[j], ai[j]);
        }
        __syncthreads();
        if (threadIdx.x < kRepairBatch && it + threadIdx.x < total) {
            double tr = 0.0, ti = 0.0;
            for (int w = 0; w < kRepairThreads / 32; ++w) { tr += part[threadIdx.x][w].x; ti += part[threadIdx.x][w].y; }
            out[mrow[threadIdx.x]] = make_float2((float)tr, (float)ti);
        }
        __syncthreads();
    }
}

}  // namespace

int fir_fftr_plan_create(FirFftPlan* pl, const double* d_taps, int ntaps, int D, int Q, cudaStream_t st) {
    *pl = FirFftPlan{};
    if (Q >= kM / 2) return IQ2A_OK;                                       // too much history: the direct form is used
    pl->M = kM;
    pl->Q = Q;
    pl->D = D;
    pl->ntaps = ntaps;
    pl->taps = d_taps;
    pl->reg = true;
    if (const char* env = std::getenv("IQ2A_PRECISE_TOL")) pl->tol = std::atof(env);
    if (const char* env = std::getenv("IQ2A_PRECISE_TOL_REL")) pl->tol_rel = std::atof(env);
    IQ2A_CUDA_TRY(cudaMalloc(&pl->tw, (size_t)kM * sizeof(double2)));
    IQ2A_CUDA_TRY(cudaMalloc(&pl->H, (size_t)D * kM * sizeof(double2)));
    IQ2A_CUDA_TRY(cudaMalloc(&pl->n_risky, 2 * sizeof(int)));
    IQ2A_CUDA_TRY(cudaMemsetAsync(pl->n_risky, 0, 2 * sizeof(int), st));
    k_fft64_twiddle<<<(kM + 255) / 256, 256, 0, st>>>(pl->tw, kM);
    const int64_t total = (int64_t)D * kM;
    k_fir_fftr_build<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_taps, ntaps, D, Q, pl->tw, pl->H);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// `d_mixed` row 0 is output row -Q.  Rows whose float64 value is within `tol` of a float32 rounding boundary are
// recomputed by the direct form; *repaired_total (device, optional) accumulates their number.
int launch_fir_fft64r(FirFftPlan& pl, const float2* d_mixed, int64_t nrows, float2* d_out, cudaStream_t st) {
    if (nrows <= 0) return IQ2A_OK;
    static bool cfg = false;
    if (!cfg) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_fir_fft64r, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        cfg = true;
    }
    if ((size_t)nrows > pl.risky_cap) {
        if (pl.risky) IQ2A_CUDA_TRY(cudaFree(pl.risky));
        pl.risky = nullptr;
        pl.risky_cap = 0;
        IQ2A_CUDA_TRY(cudaMalloc(&pl.risky, (size_t)nrows * sizeof(int)));
        pl.risky_cap = (size_t)nrows;
    }
    IQ2A_CUDA_TRY(cudaMemsetAsync(pl.n_risky, 0, sizeof(int), st));
    const int ld = kM - pl.Q;
    const unsigned grid = (unsigned)((nrows + ld - 1) / ld);
    k_fir_fft64r<<<grid, kThreads, kSmem, st>>>(d_mixed, pl.H, pl.tw, pl.D, pl.Q, nrows, d_out, pl.tol, pl.tol_rel, pl.risky, pl.n_risky,
                                                (int)pl.risky_cap);
    k_fir_repair<<<592, kRepairThreads, 0, st>>>(d_mixed, pl.taps, pl.ntaps, pl.D, pl.Q, pl.risky, pl.n_risky, (int)pl.risky_cap, d_out);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
