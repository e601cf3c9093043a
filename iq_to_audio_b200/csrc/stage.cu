// Stage-level kernels: the stand-alone drop-ins for the individual reference classes,
// the float64 "head" fix-up for the start of a stream, and the G-table builder.
#include "common.cuh"
#include "stage.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

// ---------------------------------------------------------------------------------------
// K1+K2+K3 fused: PCM frames -> IQ order/polarity -> complex NCO multiply -> complex64.
// ref: processing.py:261-279 (_extract_iq) + :289-297 (ComplexOscillator.mix).
// HBM-bound: reads 2/4/8 B and writes 8 B per sample; each thread handles 4 consecutive
// frames (one 128-bit load for s16, two 128-bit stores).
// LO = complex64(exp(1j * (phase + w*n))) with the float64 ramp of the reference; the
// product is the numpy complex64 multiply (cmul_np below).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float2 lo_f64(double phase, double w, int64_t n) {
    const double ph = __dadd_rn(phase, __dmul_rn(w, (double)n));
    double s, c;
    sincos(ph, &s, &c);
    return make_float2((float)c, (float)s);
}
// numpy's complex64 multiply as its SIMD loops evaluate it on FMA-capable x86 (AVX2/AVX-512):
// re = fma(ar, br, -(ai*bi)), im = fma(ar, bi, ai*br) -- pinned by tests/golden/stage_vectors.npz.
__device__ __forceinline__ float2 cmul_np(float2 a, float2 b) {
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)),
                       __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}

template <int FMT>
__global__ void __launch_bounds__(256) k_unpack_mix(const void* __restrict__ raw, int64_t n, int iq_swap,
                                                     int q_neg, double phase, double w,
                                                     float2* __restrict__ out) {
    using raw_t = typename RawT<FMT>::type;
    const raw_t* rp = reinterpret_cast<const raw_t*>(raw);
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    raw_t r[4];
    if (i0 + 4 <= n) {
        if constexpr (FMT == CODEC_S16) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp + i0));   // 128-bit: 4 frames
            r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
        } else if constexpr (FMT == CODEC_U8) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(rp + i0));
            r[0] = (uint16_t)(q.x & 0xffffu); r[1] = (uint16_t)(q.x >> 16);
            r[2] = (uint16_t)(q.y & 0xffffu); r[3] = (uint16_t)(q.y >> 16);
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(rp + i0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(rp + i0 + 2));
            r[0] = make_float2(a.x, a.y); r[1] = make_float2(a.z, a.w);
            r[2] = make_float2(b.x, b.y); r[3] = make_float2(b.z, b.w);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = (i0 + k < n) ? rp[i0 + k] : raw_zero<FMT>();
    }
    float2 y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = cmul_np(raw_to_c64<FMT>(r[k], iq_swap, q_neg), lo_f64(phase, w, i0 + k));
    if (i0 + 4 <= n) {
        float4* o = reinterpret_cast<float4*>(out + i0);
        o[0] = make_float4(y[0].x, y[0].y, y[1].x, y[1].y);
        o[1] = make_float4(y[2].x, y[2].y, y[3].x, y[3].y);
    } else {
        for (int k = 0; k < 4 && i0 + k < n; ++k) out[i0 + k] = y[k];
    }
}

int launch_unpack_mix(const void* d_raw, int64_t n, int codec, int iq_order, double phase, double w,
                      float2* d_out, cudaStream_t st) {
    if (n <= 0) return IQ2A_OK;
    const int swap = (iq_order == IQ2A_ORDER_QI || iq_order == IQ2A_ORDER_QI_INV);
    const int neg = (iq_order == IQ2A_ORDER_IQ_INV || iq_order == IQ2A_ORDER_QI_INV);
    const unsigned grid = (unsigned)((n + 1023) / 1024);
    switch (codec) {
        case CODEC_S16: k_unpack_mix<CODEC_S16><<<grid, 256, 0, st>>>(d_raw, n, swap, neg, phase, w, d_out); break;
        case CODEC_U8: k_unpack_mix<CODEC_U8><<<grid, 256, 0, st>>>(d_raw, n, swap, neg, phase, w, d_out); break;
        case CODEC_F32: k_unpack_mix<CODEC_F32><<<grid, 256, 0, st>>>(d_raw, n, swap, neg, phase, w, d_out); break;
        default: set_error("unknown codec %d", codec); return IQ2A_ERR_INVALID;
    }
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// ---------------------------------------------------------------------------------------
// OverlapSaveFIR.process drop-in (processing.py:325-346): exact causal FIR, full rate,
// complex64 in/out, float64 accumulation.  `x` is [history(ntaps-1) | new samples].
// Direct form: the stage-level API is for parity tests and the sign probe, not the hot
// path (the hot path never materialises the full-rate filter output).
// Tile: 256 outputs per CTA, taps streamed through shared memory in chunks of 1024.
// ---------------------------------------------------------------------------------------
constexpr int kFirOut = 256;
constexpr int kFirTapChunk = 1024;

__global__ void __launch_bounds__(kFirOut) k_fir_direct(const float2* __restrict__ x, int64_t n_out,
                                                         const double* __restrict__ taps, int ntaps,
                                                         float2* __restrict__ y) {
    __shared__ double s_taps[kFirTapChunk];
    __shared__ float2 s_x[kFirOut + kFirTapChunk];
    const int64_t o0 = (int64_t)blockIdx.x * kFirOut;
    const int64_t o = o0 + threadIdx.x;
    double ar = 0.0, ai = 0.0;
    // y[o] = sum_k taps[k] * x[(ntaps-1) + o - k]
    for (int k0 = 0; k0 < ntaps; k0 += kFirTapChunk) {
        const int kc = min(kFirTapChunk, ntaps - k0);
        for (int i = threadIdx.x; i < kc; i += kFirOut) s_taps[i] = taps[k0 + i];
        // samples needed: indices (ntaps-1) + o0 + [ -(k0+kc-1) , kFirOut-1 - k0 ]
        const int64_t lo = (int64_t)(ntaps - 1) + o0 - (k0 + kc - 1);
        const int span = kFirOut + kc - 1;
        const int64_t total = n_out + ntaps - 1;
        for (int i = threadIdx.x; i < span; i += kFirOut) {
            const int64_t idx = lo + i;
            s_x[i] = (idx >= 0 && idx < total) ? x[idx] : make_float2(0.f, 0.f);
        }
        __syncthreads();
        if (o < n_out) {
            // x index for tap k0+j: lo + (kc-1-j) + threadIdx.x  -> s_x[kc-1-j+threadIdx.x]
            for (int j = 0; j < kc; ++j) {
                const float2 v = s_x[kc - 1 - j + threadIdx.x];
                const double h = s_taps[j];
                ar = fma(h, (double)v.x, ar);
                ai = fma(h, (double)v.y, ai);
            }
        }
        __syncthreads();
    }
    if (o < n_out) y[o] = make_float2((float)ar, (float)ai);
}

int launch_fir_direct(const float2* d_x, int64_t n_out, const double* d_taps, int ntaps, float2* d_y,
                      cudaStream_t st) {
    if (n_out <= 0) return IQ2A_OK;
    k_fir_direct<<<(unsigned)((n_out + kFirOut - 1) / kFirOut), kFirOut, 0, st>>>(d_x, n_out, d_taps, ntaps, d_y);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// Decimator.process (processing.py:354-360): y[i] = x[start + i*factor]
__global__ void k_decimate(const float2* __restrict__ x, int64_t start, int factor, int64_t n_out,
                           float2* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) y[i] = x[start + i * factor];
}
int launch_decimate(const float2* d_x, int64_t start, int factor, int64_t n_out, float2* d_y, cudaStream_t st) {
    if (n_out <= 0) return IQ2A_OK;
    k_decimate<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(d_x, start, factor, n_out, d_y);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// ---------------------------------------------------------------------------------------
// Head fix-up.  The first ~ntaps/D channel samples of a stream are sums of a few tiny
// taps (|s| ~ 1e-7 .. 1e-3): the float32 transform path has ~3e-8 absolute error, which
// the NFM discriminator (angle of s) would amplify to > 1e-4.  Those rows are recomputed
// the way the reference defines them -- mixed[n] = c64(x[n]) * c64(exp(j phi(n))) in
// complex64, y = sum_k h[k] mixed[n-k] in float64 -- one CTA per (row, channel).
// ---------------------------------------------------------------------------------------
// mixed[c][idx] for idx = 0 .. n_in-1 (global sample index), exactly as k_head_direct forms it
template <int FMT>
__global__ void __launch_bounds__(256) k_head_mix(const HeadParams p, int64_t n_in) {
    using raw_t = typename RawT<FMT>::type;
    const raw_t* rp = reinterpret_cast<const raw_t*>(p.raw);
    const int c = blockIdx.y;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_in) return;
    const int64_t f = idx - p.raw_n0;
    float2 xv = make_float2(0.f, 0.f);
    if (f >= 0 && f < p.raw_len) xv = raw_to_c64<FMT>(rp[f], p.iq_swap, p.q_neg);
    const double ph = nco_phase(p.phase, c, p.w[c], idx);
    double s, co;
    sincos(ph, &s, &co);
    p.mixed[(size_t)c * p.mixed_stride + idx] = cmul_np(xv, make_float2((float)co, (float)s));
}

template <int FMT>
__global__ void __launch_bounds__(256) k_head_direct(const HeadParams p) {
    using raw_t = typename RawT<FMT>::type;
    const raw_t* rp = reinterpret_cast<const raw_t*>(p.raw);
    const int c = blockIdx.y;
    const int64_t mg = p.mg_begin + blockIdx.x;
    const int64_t n = mg * (int64_t)p.decim;
    const double* taps = p.taps + p.tap_offset[c];
    const int ntaps = p.ntaps[c];
    const int64_t kmax = min((int64_t)ntaps - 1, n);    // x[n-k] with n-k >= 0
    double ar = 0.0, ai = 0.0;
    if (p.mixed) {
        const float2* __restrict__ mx = p.mixed + (size_t)c * p.mixed_stride;
        for (int64_t k = threadIdx.x; k <= kmax; k += blockDim.x) {
            const float2 mixed = mx[n - k];
            const double h = taps[k];
            ar = fma(h, (double)mixed.x, ar);
            ai = fma(h, (double)mixed.y, ai);
        }
    } else
    for (int64_t k = threadIdx.x; k <= kmax; k += blockDim.x) {
        const int64_t idx = n - k;
        const int64_t f = idx - p.raw_n0;
        float2 xv = make_float2(0.f, 0.f);
        if (f >= 0 && f < p.raw_len) xv = raw_to_c64<FMT>(rp[f], p.iq_swap, p.q_neg);
        const double ph = nco_phase(p.phase, c, p.w[c], idx);
        double s, co;
        sincos(ph, &s, &co);
        const float2 mixed = cmul_np(xv, make_float2((float)co, (float)s));
        const double h = taps[k];
        ar = fma(h, (double)mixed.x, ar);
        ai = fma(h, (double)mixed.y, ai);
    }
    __shared__ double sr[8], si[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, off);
        ai += __shfl_xor_sync(0xffffffffu, ai, off);
    }
    if ((threadIdx.x & 31) == 0) { sr[threadIdx.x >> 5] = ar; si[threadIdx.x >> 5] = ai; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, ti = 0.0;
        for (int w = 0; w < 8; ++w) { tr += sr[w]; ti += si[w]; }
        p.out[(size_t)c * p.out_stride + (mg - p.out_mg0)] = make_float2((float)tr, (float)ti);
    }
}

int launch_head_direct(const HeadParams& p, int codec, int nrows, int nchan, cudaStream_t st) {
    if (nrows <= 0) return IQ2A_OK;
    const dim3 grid(nrows, nchan);
    if (p.mixed) {
        const int64_t n_in = (p.mg_begin + nrows - 1) * (int64_t)p.decim + 1;
        const dim3 gm((unsigned)((n_in + 255) / 256), nchan);
        switch (codec) {
            case CODEC_S16: k_head_mix<CODEC_S16><<<gm, 256, 0, st>>>(p, n_in); break;
            case CODEC_U8: k_head_mix<CODEC_U8><<<gm, 256, 0, st>>>(p, n_in); break;
            case CODEC_F32: k_head_mix<CODEC_F32><<<gm, 256, 0, st>>>(p, n_in); break;
            default: set_error("unknown codec %d", codec); return IQ2A_ERR_INVALID;
        }
    }
    switch (codec) {
        case CODEC_S16: k_head_direct<CODEC_S16><<<grid, 256, 0, st>>>(p); break;
        case CODEC_U8: k_head_direct<CODEC_U8><<<grid, 256, 0, st>>>(p); break;
        case CODEC_F32: k_head_direct<CODEC_F32><<<grid, 256, 0, st>>>(p); break;
        default: set_error("unknown codec %d", codec); return IQ2A_ERR_INVALID;
    }
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// ---------------------------------------------------------------------------------------
// G-table builder (setup, once per bank).  One CTA per (branch p, channel c):
//   g[q] = h[qD - p] * exp(-j w (qD - p)),   G[j'] = (1/M) sum_q g[q] W_M^{k(j') q}
// float64 throughout, twiddles from an exact float64 table, result rounded to float32.
// k(j') = k1 + R1*k2 for slot j' = k1*R2 + k2 (the channelizer's spectrum order).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_g(const double* __restrict__ taps, int ntaps, double w, int D,
                                                  int M, int R1, int qn, const double2* __restrict__ wtab,
                                                  float2* __restrict__ gout, int cg, int c_in_group, int layout,
                                                  double scale) {
    extern __shared__ double2 s_g[];    // [qn]
    const int p = blockIdx.x;
    for (int q = threadIdx.x; q < qn; q += blockDim.x) {
        const int64_t k = (int64_t)q * D - p;
        double2 g = make_double2(0.0, 0.0);
        if (k >= 0 && k < ntaps) {
            double s, c;
            sincos(-w * (double)k, &s, &c);
            g = make_double2(taps[k] * c, taps[k] * s);
        }
        s_g[q] = g;
    }
    __syncthreads();
    const int R2 = M / R1;
    const double inv_m = scale / (double)M;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        int bin;
        if (layout == 0) {
            const int k1 = j / R2, k2 = j % R2;
            bin = k1 + R1 * k2;
        } else {      // layouts 1 and 2 share the slot -> bin map of the packed-transform kernels
            const int r = j & 255;
            bin = (r >> 4) + 16 * (r & 15) + 256 * (j >> 8);
        }
        double ar = 0.0, ai = 0.0;
        for (int q = 0; q < qn; ++q) {
            const double2 t = wtab[(int)(((int64_t)bin * q) % M)];   // exp(-2 pi j bin q / M)
            const double2 g = s_g[q];
            ar += g.x * t.x - g.y * t.y;
            ai += g.x * t.y + g.y * t.x;
        }
        if (layout == 2) {
            // two-bins-per-thread kernel: float4 (g_k'.x, g_k'+256.x, g_k'.y, g_k'+256.y) per (p, c, r)
            float* gf = reinterpret_cast<float*>(gout) + (((size_t)p * cg + c_in_group) * 256 + (j & 255)) * 4;
            gf[j >> 8] = (float)(ar * inv_m);
            gf[2 + (j >> 8)] = (float)(ai * inv_m);
        } else {
            gout[((size_t)p * cg + c_in_group) * M + j] = make_float2((float)(ar * inv_m), (float)(ai * inv_m));
        }
    }
}

// Table of the mirror-pair kernel (channelizer5.cuh): one block per table entry e.
//   e = 4 t + q, q < 3: pair (column 4u + q + 1, column 4w + 3 - q) of tile t = (forward group u, mirror group w);
//                       weight 1/2 when the column is its own mirror, 0 when the pair was already taken (u == w)
//   e = 4 t + 3:        column 4u with the spectrum carried from the previous tile (or nothing: first tile of a class)
//   e = 4 ntiles + cls: the column left at the end of class cls (its own mirror, staged rotated by Q = A + cls):
//                       plain complex entry times W^{-Qk}; zero when the class has an odd number of groups
// The column p of the entry carries g[q] = h[qD - p] e^{-j w (qD - p - (L-1)/2)}  (= kappa^{-1/2} times the
// unpaired entry); its spectrum a + j b is stored as float4 (a_k', a_k'+256, b_k', b_k'+256) per (e, c, r),
// r <-> bin k' = (r>>4) + 16 (r&15).
__global__ void __launch_bounds__(256) k_build_gpair(const double* __restrict__ taps, int ntaps, double w, int D,
                                                     int qn, const double2* __restrict__ wtab,
                                                     float4* __restrict__ gout, int cg, int c_in_group, double scale,
                                                     PairGeo geo) {
    extern __shared__ double2 s_g[];    // [qn]
    constexpr int M = 512;
    const int e = blockIdx.x;
    int p = 0, rot = 0;
    double weight = 1.0;
    if (e < 4 * geo.ntiles) {
        const int t = e >> 2, q = e & 3;
        const int cls = t >= geo.tiles1, j = cls ? t - geo.tiles1 : t;
        const int gl = cls ? geo.g1 : 0, g = cls ? geo.g2 : geo.g1;
        const int u = gl + j, wg = gl + g - 1 - j;
        if (q < 3) {
            const int f = 4 * u + q + 1, m = 4 * wg + 3 - q;
            p = f;
            weight = f < m ? 1.0 : (f == m ? 0.5 : 0.0);
        } else {
            p = 4 * u;
        }
    } else {
        const int cls = e - 4 * geo.ntiles;
        const int gl = cls ? geo.g1 : 0, g = cls ? geo.g2 : geo.g1;
        if (g <= 0 || (g & 1)) weight = 0.0;
        else {
            p = 4 * (gl + g / 2);
            rot = geo.a + cls;
        }
    }
    const double centre = 0.5 * (double)(ntaps - 1);
    for (int q = threadIdx.x; q < qn; q += blockDim.x) {
        const int64_t k = (int64_t)q * D - p;
        double2 g = make_double2(0.0, 0.0);
        if (weight != 0.0 && k >= 0 && k < ntaps) {
            double sn, cs;
            sincos(-w * ((double)k - centre), &sn, &cs);
            g = make_double2(taps[k] * cs * weight, taps[k] * sn * weight);
        }
        s_g[q] = g;
    }
    __syncthreads();
    const double inv_m = scale / (double)M;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const int r = j & 255;
        const int bin = (r >> 4) + 16 * (r & 15) + 256 * (j >> 8);
        double ar = 0.0, ai = 0.0;
        for (int q = 0; q < qn; ++q) {
            const double2 tw = wtab[(int)(((int64_t)bin * q) % M)];   // exp(-2 pi j bin q / M)
            const double2 g = s_g[q];
            ar += g.x * tw.x - g.y * tw.y;
            ai += g.x * tw.y + g.y * tw.x;
        }
        if (rot) {                                                    // times W^{-rot bin} = conj(wtab[rot bin])
            const double2 tw = wtab[(int)(((int64_t)bin * rot) % M)];
            const double br = ar * tw.x + ai * tw.y, bi = ai * tw.x - ar * tw.y;
            ar = br;
            ai = bi;
        }
        float* gf = reinterpret_cast<float*>(gout + ((size_t)e * cg + c_in_group) * 256 + r);
        gf[j >> 8] = (float)(ar * inv_m);
        gf[2 + (j >> 8)] = (float)(ai * inv_m);
    }
}

int launch_build_gpair(const double* d_taps, int ntaps, double w, int D, int qn, const double2* d_wtab,
                       float4* d_gout, int cg, int c_in_group, double scale, const PairGeo& geo, cudaStream_t st) {
    k_build_gpair<<<pair_table_entries(geo), 256, (size_t)qn * sizeof(double2), st>>>(d_taps, ntaps, w, D, qn, d_wtab,
                                                                                      d_gout, cg, c_in_group, scale, geo);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// in-block NCO rotation table Q[c][r] = exp(j w_c D r), r < ld
__global__ void k_build_rot(const double* __restrict__ w, int D, int ld, float2* __restrict__ rot) {
    const int c = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ld) return;
    double s, co;
    sincos(w[c] * (double)D * (double)r, &s, &co);
    rot[(size_t)c * ld + r] = make_float2((float)co, (float)s);
}
int launch_build_rot(const double* d_w, int nchan, int D, int ld, float2* d_rot, cudaStream_t st) {
    const dim3 grid((ld + 127) / 128, nchan);
    k_build_rot<<<grid, 128, 0, st>>>(d_w, D, ld, d_rot);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

int launch_build_g(const double* d_taps, int ntaps, double w, int D, int M, int R1, int qn,
                   const double2* d_wtab, float2* d_gout, int cg, int c_in_group, int layout, double scale,
                   cudaStream_t st) {
    k_build_g<<<D, 256, (size_t)qn * sizeof(double2), st>>>(d_taps, ntaps, w, D, M, R1, qn, d_wtab, d_gout, cg,
                                                            c_in_group, layout, scale);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
