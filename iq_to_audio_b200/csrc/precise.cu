// Bit-faithful path for SSB channels with AGC on.
//
// Why it exists: the reference's AGC (src/iq_to_audio/decoders/ssb.py:74-79) moves its gain by
// 1e-3 * 0.251 / |x| per sample, so near an audio zero crossing (|x| ~ 1e-5) ONE sample changes
// the gain by ~25 and the following ~1000 output samples with it.  The output is therefore a
// discontinuous function of its input at the 1e-8 level, and the only way to stay within 1e-4 of
// the reference is to reproduce the reference's float32 values bit for bit up to that point:
//
//   k_mix_exact      mixed[n] = c64(x[n]) * c64(exp(j phi(n)))   -- the reference's complex64 mixer
//                    (processing.py:289-297), same float64 phase, numpy's FMA complex multiply
//   k_fir_decim_f64  s[m] = c64( sum_k h[k] mixed[mD-k] )        -- float64 polyphase direct form; the
//                    reference's complex128 FFT result rounds to the same complex64 (error 1e-16 vs
//                    half-ulp 3e-9) except with probability ~1e-7 per sample
//   k_seq_dc         DCBlocker.process (decoders/common.py:16-30) as the reference's float32 sequential
//                    recurrence, one thread per (channel, reference chunk), started 8192 rows early
//                    from zero state (0.995^8192 = 1e-18: the trajectories merge), VERIFIED against the
//                    previous chunk's true end state by k_seq_fix, which re-runs a chunk serially if the
//                    speculative start did not land on the same float32 -> bit-exact by induction
//   k_seq_agc        _apply_agc (ssb.py:67-80) as the float32 sequential loop, one thread per chunk
//                    (the gain restarts at 1.0 every chunk, so chunks are independent), then peak /
//                    clip / RMS exactly as the scan tail does.
//
// Cost: ~65 k DFMA per output sample (ntaps = 32 769); measured 18 ms per 10 s of a 20 MS/s capture with two such
// channels (tools/bench_cfg3.py).
// Only channels that need it take this path; everything else stays on the float32 transform path.
#include "common.cuh"
#include "fft64.cuh"
#include "precise.cuh"
#include "sincos_cw.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

__device__ __forceinline__ float2 cmul_np2(float2 a, float2 b) {   // numpy complex64 multiply (FMA form)
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}

// ---------------------------------------------------------------------------------------
constexpr int kMixPer = 4;             // samples per thread: the index arithmetic and parameter loads are paid once

template <int FMT>
__global__ void __launch_bounds__(256) k_mix_exact(const MixExactParams p) {
    using raw_t = typename RawT<FMT>::type;
    const raw_t* __restrict__ rp = reinterpret_cast<const raw_t*>(p.raw);
    const int64_t base = (int64_t)blockIdx.x * (256 * kMixPer);
    // nco_phase() finds the phase-table segment of a sample with a 64-bit division (~90 instructions; with the index
    // arithmetic of a one-sample thread the kernel ran ~200 instructions per sample and was issue-bound): the samples
    // of a CTA are consecutive, so the division is done once per CTA and a thread only steps over the (at most one)
    // segment boundary inside the CTA
    __shared__ int64_t s_k0;
    const int64_t seg_len = p.phase.seg_len, seg0 = p.phase.seg0_n;
    const int nseg = p.phase.nseg;
    const bool one_boundary = seg_len >= (int64_t)(256 * kMixPer);
    if (threadIdx.x == 0) {
        const int64_t rel0 = p.n0 + base - seg0;
        int64_t k = rel0 < 0 ? 0 : rel0 / seg_len;
        if (k >= nseg) k = nseg - 1;
        s_k0 = k;
    }
    __syncthreads();
    const int64_t k0 = s_k0;
    const double* __restrict__ tab = p.phase.tab + (int64_t)p.chan * nseg;
    const double t0 = tab[k0], t1 = tab[k0 + 1 < nseg ? k0 + 1 : k0];
    const int64_t rel_k0 = k0 * seg_len;
    {
        // interior CTAs (all samples present, one phase-table segment, no history rows): no per-sample tests, 32-bit
        // index arithmetic -- 64-bit compares, conversions and bounds tests were half of the general loop below
        const int64_t n_first = p.n0 + base, n_last = n_first + 256 * kMixPer - 1;
        const int64_t l_first = n_first - seg0 - rel_k0;
        const bool fast = one_boundary && base + 256 * kMixPer <= p.count && n_first >= seg0 && n_first >= 0 &&
                          n_first >= p.raw_n0 && n_last - p.raw_n0 < p.raw_len && l_first >= 0 &&
                          l_first + 256 * kMixPer < (int64_t)0x7fffffff && (l_first + 256 * kMixPer <= seg_len || k0 == nseg - 1);
        if (fast) {
            const raw_t* __restrict__ src = rp + (n_first - p.raw_n0) + threadIdx.x;
            float2* __restrict__ dst = p.mixed + base + threadIdx.x;
            const int l0 = (int)l_first + (int)threadIdx.x;
            const double w = p.w;
#pragma unroll
            for (int j = 0; j < kMixPer; ++j) {
                const float2 x = raw_to_c64<FMT>(src[j * 256], p.iq_swap, p.q_neg);
                const double ph = __dadd_rn(t0, __dmul_rn(w, (double)(l0 + j * 256)));
                double sn, cs;
                sincos_cw(ph, &sn, &cs);
                dst[j * 256] = cmul_np2(x, make_float2((float)cs, (float)sn));
            }
            return;
        }
    }
#pragma unroll
    for (int j = 0; j < kMixPer; ++j) {
        const int64_t i = base + threadIdx.x + j * 256;           // coalesced per j
        if (i >= p.count) break;
        const int64_t n = p.n0 + i;
        float2 out = make_float2(0.f, 0.f);
        const int64_t f = n - p.raw_n0;
        if (n >= 0 && f >= 0 && f < p.raw_len) {
            const float2 x = raw_to_c64<FMT>(rp[f], p.iq_swap, p.q_neg);
            double ph;
            if ((n >= seg0 || p.nhist == 0) && one_boundary) {
                const int64_t rel = n - seg0;
                int64_t local = rel - rel_k0;
                double tb = t0;
                if (rel >= 0 && local >= seg_len && k0 < nseg - 1) {
                    local -= seg_len;
                    tb = t1;
                }
                ph = __dadd_rn(tb, __dmul_rn(p.w, (double)local));
            } else if (n >= seg0 || p.nhist == 0) {
                ph = nco_phase(p.phase, p.chan, p.w, n);
            } else {
                int h = p.nhist - 1;
                while (h > 0 && n < p.hist_start[h]) --h;
                ph = __dadd_rn(p.hist_phase[h], __dmul_rn(p.w, (double)(n - p.hist_start[h])));
            }
            double sn, cs;
            sincos_cw(ph, &sn, &cs);       // within 1 ulp of sincos(), a third of its instructions at these arguments
            out = cmul_np2(x, make_float2((float)cs, (float)sn));
        }
        p.mixed[i] = out;
    }
}

int launch_mix_exact(const MixExactParams& p, int codec, cudaStream_t st) {
    if (p.count <= 0) return IQ2A_OK;
    const unsigned grid = (unsigned)((p.count + 256 * kMixPer - 1) / (256 * kMixPer));
    switch (codec) {
        case CODEC_S16: k_mix_exact<CODEC_S16><<<grid, 256, 0, st>>>(p); break;
        case CODEC_U8: k_mix_exact<CODEC_U8><<<grid, 256, 0, st>>>(p); break;
        case CODEC_F32: k_mix_exact<CODEC_F32><<<grid, 256, 0, st>>>(p); break;
        default: set_error("unknown codec %d", codec); return IQ2A_ERR_INVALID;
    }
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// ---------------------------------------------------------------------------------------
// Polyphase direct form in float64, register-tiled.
//   s[m] = sum_{p<D} sum_{q<=Q} h[qD - p] * mixed[(m - q) D + p]
// `mixed` starts at row (row0 - Q): element i <-> n = (row0 - Q) D + i.
// CTA = kFirGroups row groups x kFirLanes branch lanes; a thread owns ONE branch lane and kFirTile consecutive
// output rows, so a tap h[q] and a sliding window of kFirTile samples stay in registers: per tap one 16-byte and
// one 8-byte shared load feed 2*kFirTile DFMAs (the untiled form, one row per thread, was bound by its float ->
// double conversions and shared loads at ~12 % of the FP64 rate).  Samples are widened to double once, when the
// tile is staged.  Rows are summed over q in ascending order per lane, lanes combined by shuffles at the end.
// ---------------------------------------------------------------------------------------
constexpr int kFirLanes = 16;
constexpr int kFirTile = 8;
constexpr int kFirGroups = 16;
constexpr int kFirRows = kFirTile * kFirGroups;       // output rows per CTA

__global__ void __launch_bounds__(kFirGroups * kFirLanes)
k_fir_decim_f64(const float2* __restrict__ mixed, const double* __restrict__ taps, int ntaps, int D, int Q,
                int64_t nrows, float2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int Qp = (Q + kFirTile) & ~(kFirTile - 1);      // taps padded to a multiple of the tile (zeros): Qp >= Q + 1
    double2* sx = reinterpret_cast<double2*>(sm);                                   // [kFirRows + Qp][kFirLanes]
    double* st = reinterpret_cast<double*>(sx + (size_t)(kFirRows + Qp) * kFirLanes);   // [Qp][kFirLanes]
    const int lane = threadIdx.x % kFirLanes;
    const int g = threadIdx.x / kFirLanes;
    const int64_t m0 = (int64_t)blockIdx.x * kFirRows;       // first output row of this CTA (relative)
    double ar[kFirTile], ai[kFirTile];
#pragma unroll
    for (int i = 0; i < kFirTile; ++i) ar[i] = ai[i] = 0.0;
    // buffer row j <-> output row m0 - Qp + j; `mixed` row 0 is output row -Q, so source row = m0 + j - (Qp - Q)
    const int64_t src0 = m0 - (Qp - Q);
    const int64_t src_rows = nrows + Q;                      // rows available in `mixed`
    for (int p0 = 0; p0 < D; p0 += kFirLanes) {
        const int p = p0 + lane;
        for (int j = g; j < kFirRows + Qp; j += kFirGroups) {
            const int64_t sr = src0 + j;
            double2 v = make_double2(0.0, 0.0);
            if (p < D && sr >= 0 && sr < src_rows) {
                const float2 f = mixed[sr * (int64_t)D + p];
                v = make_double2((double)f.x, (double)f.y);
            }
            sx[j * kFirLanes + lane] = v;
        }
        for (int q = g; q < Qp; q += kFirGroups) {
            const int64_t k = (int64_t)q * D - p;
            st[q * kFirLanes + lane] = (p < D && q <= Q && k >= 0 && k < ntaps) ? taps[k] : 0.0;
        }
        __syncthreads();
        // row m0 + T g + i at tap q reads buffer row (T g + i + Qp - q), T = kFirTile; window slot of buffer row j is j mod T
        const double2* xw = sx + (size_t)(kFirTile * g + Qp) * kFirLanes + lane;     // buffer row of (i = 0, q = 0)
        double2 c[kFirTile];
#pragma unroll
        for (int i = 0; i < kFirTile; ++i) c[i] = xw[i * kFirLanes];                  // rows base .. base+T-1, base % T == 0
        const double* hp = st + lane;
        for (int q4 = 0; q4 < Qp; q4 += kFirTile) {
#pragma unroll
            for (int u = 0; u < kFirTile; ++u) {
                const double h = hp[(q4 + u) * kFirLanes];
#pragma unroll
                for (int i = 0; i < kFirTile; ++i) {
                    const double2 v = c[(i - u) & (kFirTile - 1)];
                    ar[i] = fma(h, v.x, ar[i]);
                    ai[i] = fma(h, v.y, ai[i]);
                }
                // the sample one row further back replaces the one that leaves the window
                c[(kFirTile - 1 - u) & (kFirTile - 1)] = xw[-(q4 + u + 1) * kFirLanes];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < kFirTile; ++i) {
#pragma unroll
        for (int off = kFirLanes / 2; off > 0; off >>= 1) {
            ar[i] += __shfl_xor_sync(0xffffffffu, ar[i], off);
            ai[i] += __shfl_xor_sync(0xffffffffu, ai[i], off);
        }
        const int64_t m = m0 + kFirTile * g + i;
        if (lane == 0 && m < nrows) out[m] = make_float2((float)ar[i], (float)ai[i]);
    }
}

// The same sum with one thread per output row (taps in ascending order): for the low-rate captures whose history does
// not fit the tiled kernel -- a 144 kS/s capture is 144 k rows x 1025 taps per second of signal.
__global__ void __launch_bounds__(128) k_fir_rows_f64(const float2* __restrict__ mixed, const double* __restrict__ taps,
                                                      int ntaps, int D, int Q, int64_t nrows, float2* __restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nrows) return;
    const float2* x = mixed + (m + Q) * (int64_t)D;          // the sample that meets h[0]; `mixed` row 0 is output row -Q
    double ar = 0.0, ai = 0.0;
    for (int k = 0; k < ntaps; ++k) {
        const float2 v = x[-k];
        const double h = __ldg(taps + k);
        ar = fma(h, (double)v.x, ar);
        ai = fma(h, (double)v.y, ai);
    }
    out[m] = make_float2((float)ar, (float)ai);
}

int launch_fir_decim_f64(const float2* d_mixed, const double* d_taps, int ntaps, int D, int Q, int64_t nrows,
                         float2* d_out, cudaStream_t st) {
    if (nrows <= 0) return IQ2A_OK;
    const int Qp = (Q + kFirTile) & ~(kFirTile - 1);
    const size_t smem = (size_t)(kFirRows + Qp) * kFirLanes * sizeof(double2) + (size_t)Qp * kFirLanes * sizeof(double);
    if (smem > 220 * 1024) {
        // more history rows than the tiled kernel's shared memory holds (D = 1..2): one thread per output row
        k_fir_rows_f64<<<(unsigned)((nrows + 127) / 128), 128, 0, st>>>(d_mixed, d_taps, ntaps, D, Q, nrows, d_out);
        IQ2A_CUDA_TRY(cudaGetLastError());
        return IQ2A_OK;
    }
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_fir_decim_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const unsigned grid = (unsigned)((nrows + kFirRows - 1) / kFirRows);
    k_fir_decim_f64<<<grid, kFirGroups * kFirLanes, smem, st>>>(d_mixed, d_taps, ntaps, D, Q, nrows, d_out);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// ---------------------------------------------------------------------------------------
// Transform form: s = sum_p h_p * x_p per overlap-save block of M rows, x_p[m] = mixed[mD + p], h_p[q] = h[qD - p].
//   Y[k] = sum_p H_p[k] X_p[k],  H_p = FFT_M(h_p) / M,  rows >= Q of IFFT_M(Y) are exact convolution outputs.
// One CTA per block: branch columns are transformed eight at a time in shared memory (fft64.cuh), every thread
// accumulates its bins of Y in registers, then one inverse transform (forward transform of the conjugate).
// Float64 throughout, so the result rounds to the same complex64 as the direct form and as the reference's own
// complex128 transform, up to the same ~1e-8 chance per sample of landing on the other side of a rounding boundary.
// ---------------------------------------------------------------------------------------
constexpr int kFftCols = 8;
constexpr int kFftThreads = 512;

__global__ void k_fir_fft_build(const double* __restrict__ taps, int ntaps, int D, int Q, int M,
                                const double2* __restrict__ tw, double2* __restrict__ H) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * M) return;
    const int p = (int)(i / M), k = (int)(i - (int64_t)p * M);
    double re = 0.0, im = 0.0;
    for (int q = 0; q <= Q; ++q) {
        const int64_t t = (int64_t)q * D - p;
        if (t < 0 || t >= ntaps) continue;
        const double2 w = tw[(int)(((int64_t)k * q) & (M - 1))];
        re = fma(taps[t], w.x, re);
        im = fma(taps[t], w.y, im);
    }
    H[i] = make_double2(re / M, im / M);
}

template <int M>
__global__ void __launch_bounds__(kFftThreads) k_fir_fft64(const float2* __restrict__ mixed, const double2* __restrict__ H,
                                                           const double2* __restrict__ tw, int D, int Q, int64_t nrows,
                                                           float2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm[];
    double2* s = reinterpret_cast<double2*>(sm);                          // [M][kFftCols]
    constexpr int BPT = M / kFftThreads;                                  // bins per thread
    constexpr int LOG2M = M == 1024 ? 10 : 9;
    const int ld = M - Q;
    const int64_t row0 = (int64_t)blockIdx.x * ld;                         // first `mixed` row of this block
    const int64_t src_rows = nrows + Q;
    double2 acc[BPT];
    int kr[BPT];
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        acc[i] = make_double2(0.0, 0.0);
        kr[i] = bitrev_n(threadIdx.x + i * kFftThreads, LOG2M);
    }
    for (int p0 = 0; p0 < D; p0 += kFftCols) {
        for (int idx = threadIdx.x; idx < M * kFftCols; idx += kFftThreads) {
            const int j = idx / kFftCols, c = idx % kFftCols;
            const int p = p0 + c;
            const int64_t sr = row0 + j;
            double2 v = make_double2(0.0, 0.0);
            if (p < D && sr < src_rows) {
                const float2 f = mixed[sr * (int64_t)D + p];
                v = make_double2((double)f.x, (double)f.y);
            }
            s[idx] = v;
        }
        __syncthreads();
        fft_dif_shared<kFftCols>(s, M, tw, M);                             // X_p[k] at row bitrev(k)
        const int ncol = min(kFftCols, D - p0);
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int k = threadIdx.x + i * kFftThreads;
            const double2* xs = s + (size_t)kr[i] * kFftCols;
            const double2* hs = H + (size_t)p0 * M + k;
            for (int c = 0; c < ncol; ++c) {
                const double2 h = hs[(size_t)c * M], x = xs[c];
                acc[i].x = fma(h.x, x.x, fma(-h.y, x.y, acc[i].x));
                acc[i].y = fma(h.x, x.y, fma(h.y, x.x, acc[i].y));
            }
        }
        __syncthreads();
    }
    // inverse transform: IFFT(Y) = conj(FFT(conj(Y))) (1/M is in H)
#pragma unroll
    for (int i = 0; i < BPT; ++i) s[threadIdx.x + i * kFftThreads] = make_double2(acc[i].x, -acc[i].y);
    __syncthreads();
    fft_dif_shared<1>(s, M, tw, M);
    for (int j = Q + threadIdx.x; j < M; j += kFftThreads) {
        const int64_t m = row0 + (j - Q);
        if (m < nrows) {
            const double2 v = s[bitrev_n(j, LOG2M)];
            out[m] = make_float2((float)v.x, (float)(-v.y));
        }
    }
}

int fir_fft_plan_create(FirFftPlan* pl, const double* d_taps, int ntaps, int D, int Q, cudaStream_t st) {
    *pl = FirFftPlan{};
    const int M = Q < 96 ? 512 : 1024;                                    // keep at least ~80 % of a block useful
    if (Q >= M / 2) return IQ2A_OK;                                       // too much history: the direct form is used
    pl->M = M;
    pl->Q = Q;
    pl->D = D;
    IQ2A_CUDA_TRY(cudaMalloc(&pl->tw, (size_t)M * sizeof(double2)));
    IQ2A_CUDA_TRY(cudaMalloc(&pl->H, (size_t)D * M * sizeof(double2)));
    k_fft64_twiddle<<<(M + 255) / 256, 256, 0, st>>>(pl->tw, M);
    const int64_t total = (int64_t)D * M;
    k_fir_fft_build<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_taps, ntaps, D, Q, M, pl->tw, pl->H);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

void fir_fft_plan_destroy(FirFftPlan* pl) {
    if (pl->H) cudaFree(pl->H);
    if (pl->tw) cudaFree(pl->tw);
    if (pl->risky) cudaFree(pl->risky);
    if (pl->n_risky) cudaFree(pl->n_risky);
    *pl = FirFftPlan{};
}

int launch_fir_fft64(const FirFftPlan& pl, const float2* d_mixed, int64_t nrows, float2* d_out, cudaStream_t st) {
    if (nrows <= 0) return IQ2A_OK;
    const int ld = pl.M - pl.Q;
    const unsigned grid = (unsigned)((nrows + ld - 1) / ld);
    const size_t smem = (size_t)pl.M * kFftCols * sizeof(double2);
    if (pl.M == 1024) {
        static bool cfg = false;
        if (!cfg) { IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_fir_fft64<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
        k_fir_fft64<1024><<<grid, kFftThreads, smem, st>>>(d_mixed, pl.H, pl.tw, pl.D, pl.Q, nrows, d_out);
    } else {
        static bool cfg = false;
        if (!cfg) { IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_fir_fft64<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
        k_fir_fft64<512><<<grid, kFftThreads, smem, st>>>(d_mixed, pl.H, pl.tw, pl.D, pl.Q, nrows, d_out);
    }
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// ---------------------------------------------------------------------------------------
// sequential float32 tail
// ---------------------------------------------------------------------------------------
// rows of reference chunk k (global chunk index) intersected with [mg0, mg0 + n)
__device__ __forceinline__ void chunk_rows(const SeqParams& p, int64_t kchunk, int64_t* lo, int64_t* hi) {
    const int64_t d = p.decim;
    int64_t a = (p.seg_origin + kchunk * p.seg_len + d - 1) / d;
    int64_t b = (p.seg_origin + (kchunk + 1) * p.seg_len + d - 1) / d;
    a = max(a, p.mg0);
    b = min(b, p.mg0 + p.n);
    *lo = a - p.mg0;
    *hi = max(b - p.mg0, a - p.mg0);
}

// One step of the reference recurrence (common.py:24): y = (x - x1) + r*y1 in float32; on the first
// sample of a call r*y1 is a float64 product rounded once to float32 (common.py:20-21, Python floats).
__device__ __forceinline__ float dc_step(float x, float x1, float y1, bool first_of_call) {
    const float a = __fsub_rn(x, x1);
    const float b = first_of_call ? (float)(0.995 * (double)y1) : __fmul_rn(0.995f, y1);
    return __fadd_rn(a, b);
}

constexpr int kSeqWarm = 8192;

// The float32 recurrences are run by one WARP per (chunk, channel): the lanes load 32 consecutive rows at once
// (coalesced, and whatever does not depend on the carried value -- input differences, the AGC's target/|x| division --
// is computed lane-parallel), then all lanes step through the 32 dependent updates together, fetching row j's term by
// shuffle; lane j keeps the j-th result for a coalesced store.  One thread walking the rows alone spent ~350 cycles
// per row waiting for its own loads (8.9 ms per 4 s of capture; 0.3 ms this way).  The arithmetic per row is the
// same sequence of float32 operations as before.

// rows [a, b) of the DC blocker from state (x1, y1); stores when y_out != nullptr.  All lanes return the same state.
__device__ __forceinline__ void dc_rows_warp(const float* __restrict__ x, float* __restrict__ y_out, int64_t a, int64_t b,
                                             float& x1, float& y1) {
    const int lane = threadIdx.x & 31;
    // the rows of the next kAhead groups are requested before the dependent chain of the current group runs: the chain
    // (32 x (FMUL, FADD), ~260 cycles) is shorter than a load from HBM, so one group of look-ahead is not enough
    constexpr int kAhead = 4;
    float nxt[kAhead];
#pragma unroll
    for (int i = 0; i < kAhead; ++i) {
        const int64_t r = a + 32 * i + lane;
        nxt[i] = r < b ? x[r] : 0.f;
    }
    for (int64_t base = a; base < b; base += 32 * kAhead) {
#pragma unroll
        for (int i = 0; i < kAhead; ++i) {
            const int64_t gbase = base + 32 * i;
            if (gbase >= b) break;
            const int64_t r = gbase + lane;
            const int cnt = (int)min((int64_t)32, b - gbase);
            const float xv = nxt[i];
            {
                const int64_t rn = gbase + 32 * kAhead + lane;
                nxt[i] = rn < b ? x[rn] : 0.f;
            }
            float xprev = __shfl_up_sync(0xffffffffu, xv, 1);
            if (lane == 0) xprev = x1;
            const float diff = __fsub_rn(xv, xprev);
            float mine = 0.f;
            if (cnt == 32) {
                // full group: the 32 shuffles do not depend on the carried value, so with the loop unrolled they are
                // all in flight before the chain of 32 x (FMUL, FADD) starts -- the chain is what remains
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float dj = __shfl_sync(0xffffffffu, diff, j);
                    y1 = __fadd_rn(dj, __fmul_rn(0.995f, y1));
                    if (lane == j) mine = y1;
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    const float dj = __shfl_sync(0xffffffffu, diff, j);
                    y1 = __fadd_rn(dj, __fmul_rn(0.995f, y1));
                    if (lane == j) mine = y1;
                }
            }
            if (y_out && r < b) y_out[r] = mine;
            x1 = __shfl_sync(0xffffffffu, xv, cnt - 1);
        }
    }
}

// pass A: one warp per (chunk, channel): DC blocker over the chunk's rows from a speculative start
__global__ void __launch_bounds__(32) k_seq_dc(const SeqParams p) {
    const int kc = blockIdx.x;                                // local chunk index
    const int ci = blockIdx.y;
    const int lane = threadIdx.x;
    const int c = p.chan_idx[ci];
    int64_t lo, hi;
    chunk_rows(p, p.chunk0 + kc, &lo, &hi);
    SeqChunk& rec = p.rec[(size_t)ci * p.nchunks + kc];
    const float* x = p.pre + (size_t)c * p.work_stride;
    float* y = p.tmp + (size_t)c * p.work_stride;
    float x1, y1;
    const bool exact_start = (kc == 0 && !p.fresh);
    if (exact_start) {
        x1 = p.state[c].dc_x;
        y1 = p.state[c].dc_y;
    } else {
        // speculative: run the recurrence over up to kSeqWarm rows before the chunk from zero state
        const int64_t w0 = max((int64_t)0, lo - kSeqWarm);
        x1 = w0 > 0 ? x[w0 - 1] : (p.fresh ? 0.f : p.state[c].dc_x);
        y1 = (w0 == 0 && !p.fresh) ? p.state[c].dc_y : 0.f;
        dc_rows_warp(x, nullptr, w0, lo, x1, y1);
    }
    const float y_start = y1;
    if (hi > lo) {
        // first row of the call: r * y1 is a float64 product there (dc_step)
        const float xv = x[lo];
        y1 = dc_step(xv, x1, y1, true);
        if (lane == 0) y[lo] = y1;
        x1 = xv;
        dc_rows_warp(x, y, lo + 1, hi, x1, y1);
    }
    if (lane == 0) {
        rec.lo = lo;
        rec.hi = hi;
        rec.y_start = y_start;
        rec.y_end = y1;
    }
}

// pass B: one thread per channel walks the chunks; a chunk whose speculative start state differs from
// the true end state of its predecessor is recomputed serially (rare: the recurrence is a contraction).
__global__ void k_seq_fix(const SeqParams p) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= p.nprecise) return;
    const int c = p.chan_idx[ci];
    const float* x = p.pre + (size_t)c * p.work_stride;
    float* y = p.tmp + (size_t)c * p.work_stride;
    SeqChunk* rec = p.rec + (size_t)ci * p.nchunks;
    int repaired = 0;
    for (int kc = 1; kc < p.nchunks; ++kc) {
        if (rec[kc].hi <= rec[kc].lo) { rec[kc].y_end = rec[kc - 1].y_end; continue; }
        const float y_true = rec[kc - 1].y_end;
        if (__float_as_uint(rec[kc].y_start) == __float_as_uint(y_true)) continue;
        const int64_t lo = rec[kc].lo, hi = rec[kc].hi;
        float x1 = lo > 0 ? x[lo - 1] : 0.f, y1 = y_true;
        for (int64_t r = lo; r < hi; ++r) {
            const float xv = x[r];
            y1 = dc_step(xv, x1, y1, r == lo);
            y[r] = y1;
            x1 = xv;
        }
        rec[kc].y_end = y1;
        ++repaired;
    }
    if (p.n > 0) {
        p.state[c].dc_x = x[p.n - 1];
        p.state[c].dc_y = rec[p.nchunks - 1].y_end;
    }
    if (p.repaired) atomicAdd(p.repaired, repaired);
}

// pass C: AGC (ssb.py:67-80) per chunk, then the writer-side peak / clip and the chunk statistic; one warp per
// (chunk, channel), see above
__global__ void __launch_bounds__(32) k_seq_agc(const SeqParams p) {
    const int kc = blockIdx.x;
    const int ci = blockIdx.y;
    const int lane = threadIdx.x;
    const int c = p.chan_idx[ci];
    const SeqChunk rec = p.rec[(size_t)ci * p.nchunks + kc];
    const float* __restrict__ x = p.tmp + (size_t)c * p.work_stride;
    const float target = (float)p.agc_target, decay = (float)p.agc_decay;
    float gain = 1.0f;                                                   // restarts every call (ssb.py:72)
    float peak = 0.f;
    double ss = 0.0;
    constexpr int kAhead = 4;                                            // groups of 32 rows requested ahead (see dc_rows_warp)
    float nxt[kAhead];
#pragma unroll
    for (int i = 0; i < kAhead; ++i) {
        const int64_t r = rec.lo + 32 * i + lane;
        nxt[i] = r < rec.hi ? x[r] : 0.f;
    }
    for (int64_t base0 = rec.lo; base0 < rec.hi; base0 += 32 * kAhead) {
#pragma unroll
        for (int i = 0; i < kAhead; ++i) {
            const int64_t base = base0 + 32 * i;
            if (base >= rec.hi) break;
            const int64_t r = base + lane;
            const int cnt = (int)min((int64_t)32, rec.hi - base);
            const float sv = nxt[i];
            {
                const int64_t rn = base + 32 * kAhead + lane;
                nxt[i] = rn < rec.hi ? x[rn] : 0.f;
            }
            const float mag = fabsf(sv);
            const bool live = r < rec.hi && mag > 1e-6f;
            const float desired = live ? __fdiv_rn(target, mag) : 0.f;
            // a row below the floor leaves the gain alone (ssb.py:75-76): with a zero step factor the update
            // gain + 0 * (desired - gain) is exactly that, and the factor does not wait for the carried gain
            const float step = live ? decay : 0.f;
            const unsigned mask = __ballot_sync(0xffffffffu, live);
            float mine = 1.0f;
            if (cnt == 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float dj = __shfl_sync(0xffffffffu, desired, j);
                    const float sj = __shfl_sync(0xffffffffu, step, j);
                    gain = __fadd_rn(gain, __fmul_rn(sj, __fsub_rn(dj, gain)));
                    if (lane == j) mine = gain;
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    const float dj = __shfl_sync(0xffffffffu, desired, j);
                    if (mask & (1u << j)) gain = __fadd_rn(gain, __fmul_rn(decay, __fsub_rn(dj, gain)));
                    if (lane == j) mine = gain;
                }
            }
            if (r >= rec.hi || r < p.n_skip) continue;
            const float o = __fmul_rn(sv, mine);
            const int64_t oi = r - p.n_skip;
            if (p.audio) p.audio[(size_t)c * p.out_stride + oi] = o;
            if (p.clipped) p.clipped[(size_t)c * p.out_stride + oi] = fminf(fmaxf(o, -0.99f), 0.99f);
            peak = fmaxf(peak, fabsf(o));
            ss = fma((double)o, (double)o, ss);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, off));
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if (lane != 0) return;
    if (peak > 0.f) atomicMax(reinterpret_cast<unsigned int*>(&p.state[c].peak), __float_as_uint(peak));
    if (p.sumsq) {
        int64_t w = p.chunk0 + kc - p.win_chunk0;
        if (w < 0) w = 0;
        if (w >= p.nwin) w = p.nwin - 1;
        atomicAdd(p.sumsq + (size_t)c * p.nwin + w, ss);
    }
}

int launch_seq_tail(const SeqParams& p, cudaStream_t st, int64_t* launches) {
    if (p.nprecise <= 0 || p.nchunks <= 0 || p.n <= 0) return IQ2A_OK;
    const dim3 grid(p.nchunks, p.nprecise);                   // one warp per (chunk, channel)
    k_seq_dc<<<grid, 32, 0, st>>>(p);
    k_seq_fix<<<(p.nprecise + 31) / 32, 32, 0, st>>>(p);
    k_seq_agc<<<grid, 32, 0, st>>>(p);
    if (launches) *launches += 3;
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
