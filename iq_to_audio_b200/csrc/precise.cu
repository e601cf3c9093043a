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
// Cost: ~65 k DFMA per output sample (ntaps = 32 769) -> ~25 ms per channel for 60 s at 20 MS/s.
// Only channels that need it take this path; everything else stays on the float32 transform path.
#include "common.cuh"
#include "precise.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

__device__ __forceinline__ float2 cmul_np2(float2 a, float2 b) {   // numpy complex64 multiply (FMA form)
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}

// ---------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) k_mix_exact(const MixExactParams p) {
    using raw_t = typename RawT<FMT>::type;
    const raw_t* rp = reinterpret_cast<const raw_t*>(p.raw);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.count) return;
    const int64_t n = p.n0 + i;
    float2 out = make_float2(0.f, 0.f);
    const int64_t f = n - p.raw_n0;
    if (n >= 0 && f >= 0 && f < p.raw_len) {
        const float2 x = raw_to_c64<FMT>(rp[f], p.iq_swap, p.q_neg);
        const double ph = nco_phase(p.phase, p.chan, p.w, n);
        double s, c;
        sincos(ph, &s, &c);
        out = cmul_np2(x, make_float2((float)c, (float)s));
    }
    p.mixed[i] = out;
}

int launch_mix_exact(const MixExactParams& p, int codec, cudaStream_t st) {
    if (p.count <= 0) return IQ2A_OK;
    const unsigned grid = (unsigned)((p.count + 255) / 256);
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
// Polyphase direct form in float64.  CTA = kFirRows output rows x kFirLanes branch lanes.
//   s[m] = sum_{p<D} sum_{q<=Q} h[qD - p] * mixed[(m - q) D + p]
// `mixed` starts at row (row0 - Q): element i <-> n = (row0 - Q) D + i.
// ---------------------------------------------------------------------------------------
constexpr int kFirRows = 32;
constexpr int kFirLanes = 16;

__global__ void __launch_bounds__(kFirRows * kFirLanes)
k_fir_decim_f64(const float2* __restrict__ mixed, const double* __restrict__ taps, int ntaps, int D, int Q,
                int64_t nrows, float2* __restrict__ out) {
    extern __shared__ unsigned char sm[];
    float2* sx = reinterpret_cast<float2*>(sm);                                   // [(kFirRows + Q)][kFirLanes]
    double* st = reinterpret_cast<double*>(sx + (size_t)(kFirRows + Q) * kFirLanes);   // [Q + 1][kFirLanes]
    const int lane = threadIdx.x % kFirLanes;
    const int r = threadIdx.x / kFirLanes;
    const int64_t m0 = (int64_t)blockIdx.x * kFirRows;       // first output row of this CTA (relative)
    double ar = 0.0, ai = 0.0;
    for (int p0 = 0; p0 < D; p0 += kFirLanes) {
        const int p = p0 + lane;
        // rows (m0 - Q .. m0 + kFirRows - 1) of branch columns p0..p0+15; buffer row j <-> output row m0 - Q + j
        for (int j = r; j < kFirRows + Q; j += kFirRows) {
            float2 v = make_float2(0.f, 0.f);
            if (p < D) v = mixed[(m0 + j) * (int64_t)D + p];
            sx[j * kFirLanes + lane] = v;
        }
        for (int q = r; q <= Q; q += kFirRows) {
            const int64_t k = (int64_t)q * D - p;
            st[q * kFirLanes + lane] = (p < D && k >= 0 && k < ntaps) ? taps[k] : 0.0;
        }
        __syncthreads();
        // output row m0 + r uses buffer rows (r + Q - q)
        const float2* xr = sx + (size_t)(r + Q) * kFirLanes + lane;
        for (int q = 0; q <= Q; ++q) {
            const float2 v = xr[-(q * kFirLanes)];
            const double h = st[q * kFirLanes + lane];
            ar = fma(h, (double)v.x, ar);
            ai = fma(h, (double)v.y, ai);
        }
        __syncthreads();
    }
#pragma unroll
    for (int off = kFirLanes / 2; off > 0; off >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, off);
        ai += __shfl_xor_sync(0xffffffffu, ai, off);
    }
    if (lane == 0 && m0 + r < nrows) out[m0 + r] = make_float2((float)ar, (float)ai);
}

int launch_fir_decim_f64(const float2* d_mixed, const double* d_taps, int ntaps, int D, int Q, int64_t nrows,
                         float2* d_out, cudaStream_t st) {
    if (nrows <= 0) return IQ2A_OK;
    const size_t smem = (size_t)(kFirRows + Q) * kFirLanes * sizeof(float2) + (size_t)(Q + 1) * kFirLanes * sizeof(double);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_fir_decim_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const unsigned grid = (unsigned)((nrows + kFirRows - 1) / kFirRows);
    k_fir_decim_f64<<<grid, kFirRows * kFirLanes, smem, st>>>(d_mixed, d_taps, ntaps, D, Q, nrows, d_out);
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

// pass A: one thread per (chunk, channel): DC blocker over the chunk's rows from a speculative start
__global__ void k_seq_dc(const SeqParams p) {
    const int kc = blockIdx.x * blockDim.x + threadIdx.x;     // local chunk index
    const int ci = blockIdx.y;
    if (kc >= p.nchunks) return;
    const int c = p.chan_idx[ci];
    int64_t lo, hi;
    chunk_rows(p, p.chunk0 + kc, &lo, &hi);
    SeqChunk& rec = p.rec[(size_t)ci * p.nchunks + kc];
    rec.lo = lo;
    rec.hi = hi;
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
        for (int64_t r = w0; r < lo; ++r) {
            const float xv = x[r];
            y1 = dc_step(xv, x1, y1, false);
            x1 = xv;
        }
    }
    rec.y_start = y1;
    for (int64_t r = lo; r < hi; ++r) {
        const float xv = x[r];
        y1 = dc_step(xv, x1, y1, r == lo);
        y[r] = y1;
        x1 = xv;
    }
    rec.y_end = y1;
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

// pass C: AGC (ssb.py:67-80) per chunk, then the writer-side peak / clip and the chunk statistic
__global__ void k_seq_agc(const SeqParams p) {
    const int kc = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci = blockIdx.y;
    if (kc >= p.nchunks) return;
    const int c = p.chan_idx[ci];
    const SeqChunk rec = p.rec[(size_t)ci * p.nchunks + kc];
    const float* x = p.tmp + (size_t)c * p.work_stride;
    const float target = (float)p.agc_target, decay = (float)p.agc_decay;
    float gain = 1.0f;                                                   // restarts every call (ssb.py:72)
    float peak = 0.f;
    double ss = 0.0;
    for (int64_t r = rec.lo; r < rec.hi; ++r) {
        const float s = x[r];
        const float mag = fabsf(s);
        if (mag > 1e-6f) {
            const float desired = __fdiv_rn(target, mag);
            gain = __fadd_rn(gain, __fmul_rn(decay, __fsub_rn(desired, gain)));
        }
        const float o = __fmul_rn(s, gain);
        if (r < p.n_skip) continue;
        const int64_t oi = r - p.n_skip;
        if (p.audio) p.audio[(size_t)c * p.out_stride + oi] = o;
        if (p.clipped) p.clipped[(size_t)c * p.out_stride + oi] = fminf(fmaxf(o, -0.99f), 0.99f);
        peak = fmaxf(peak, fabsf(o));
        ss = fma((double)o, (double)o, ss);
    }
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
    const dim3 grid((p.nchunks + 63) / 64, p.nprecise);
    k_seq_dc<<<grid, 64, 0, st>>>(p);
    k_seq_fix<<<(p.nprecise + 31) / 32, 32, 0, st>>>(p);
    k_seq_agc<<<grid, 64, 0, st>>>(p);
    if (launches) *launches += 3;
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
