// Channel-rate tail: discriminator / envelope / product detector, then the decoder
// recurrences as parallel affine scans in float64, then the writer-side peak / clip and
// the per-chunk statistics.
//
// Reference arithmetic replaced (src/iq_to_audio/decoders/):
//   nfm.py:17-24   QuadratureDemod.process   angle(s[n] conj(s[n-1]))            -> k_pre
//   am.py:28       abs(s)                                                        -> k_pre
//   ssb.py:42-43   real(s) (USB) / real(conj s) (LSB)                            -> k_pre
//   nfm.py:49-62   DeemphasisFilter: y = beta x + z ; z = alpha y (lfilter DF2T) -> scan DEEMPH
//   common.py:16-30 DCBlocker: y = x - x1 + r y1                                 -> scan DC
//   ssb.py:67-80   AGC: g += decay (target/|x| - g), g restarts at 1 per chunk   -> scan AGC
//   processing.py:449-453 AudioWriter.write: running pre-clip peak, clip +-0.99  -> emit
//   nfm.py:87-89   per-chunk RMS dBFS (sum of squares accumulated here)          -> emit
//
// Every recurrence is first-order and affine in its carried variable,
//   v[m] = A_m v[m-1] + B_m,
// so it composes associatively: (A2,B2) o (A1,B1) = (A2 A1, A2 B1 + B2).  Three kernels:
// per-tile aggregate, a serial walk over tile aggregates (a few thousand steps per
// channel), per-tile apply.  All in float64: the reference runs the de-emphasis in
// float64 (lfilter) and the DC blocker / AGC as float32 sequential loops; float64 scans
// differ from those by < 1e-5 pre-clip, well inside the 1e-4 tolerance.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>

#include "common.cuh"
#include "tail.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

constexpr int kScanThreads = 256;
constexpr int kScanPer = 4;
constexpr int kScanTile = kScanThreads * kScanPer;

struct Aff {
    double a, b;
};
__device__ __forceinline__ Aff compose(Aff first, Aff second) {   // second o first
    return Aff{second.a * first.a, fma(second.a, first.b, second.b)};
}

__device__ __forceinline__ bool is_chunk_start(const TailParams& p, int64_t r) {
    const int64_t n = (p.mg0 + r) * (int64_t)p.decim - p.seg_origin;
    if (n < 0) return false;
    return (n % p.seg_len) < p.decim;
}

// ---------------------------------------------------------------------------------------
// detector front: complex64 channel samples -> float32 "pre-audio"
// ---------------------------------------------------------------------------------------
__global__ void k_pre(const TailParams p) {
    const int c = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n) return;
    const TailChan ch = p.chan[c];
    const float2* bb = p.bb + (size_t)c * p.bb_stride;
    const float2 s = bb[r];
    float out;
    if (ch.mode == MODE_NFM) {
        float2 prev;
        if (r > 0) prev = bb[r - 1];
        else if (p.fresh) prev = make_float2(1.f, 0.f);          // nfm.py:15
        else prev = make_float2(p.state[c].prev_re, p.state[c].prev_im);
        // numpy complex64 product s * conj(prev) (FMA form of its SIMD loop), then arctan2 (nfm.py:21-22)
        const float re = __fmaf_rn(s.x, prev.x, __fmul_rn(s.y, prev.y));
        const float im = __fmaf_rn(s.x, -prev.y, __fmul_rn(s.y, prev.x));
        out = atan2f(im, re);
    } else if (ch.mode == MODE_AM) {
        out = hypotf(s.x, s.y);                                  // am.py:28
    } else {
        out = s.x;                                               // ssb.py:42-43
    }
    p.pre[(size_t)c * p.work_stride + r] = out;
}

// ---------------------------------------------------------------------------------------
// scan element definitions
// ---------------------------------------------------------------------------------------
enum ScanKind : int { SCAN_NONE = -1, SCAN_DEEMPH = 0, SCAN_DC = 1, SCAN_AGC = 2 };

__device__ __forceinline__ int scan_kind(const TailChan& ch, int pass) {
    if (ch.precise) return SCAN_NONE;
    if (pass == 0) {
        if (ch.mode == MODE_NFM || ch.mode == MODE_RAW_DEEMPH) return SCAN_DEEMPH;
        if (ch.mode == MODE_AM || ch.mode == MODE_USB || ch.mode == MODE_LSB || ch.mode == MODE_RAW_DC) return SCAN_DC;
        if (ch.mode == MODE_RAW_AGC) return SCAN_AGC;
        return SCAN_NONE;
    }
    if ((ch.mode == MODE_USB || ch.mode == MODE_LSB) && ch.agc) return SCAN_AGC;
    return SCAN_NONE;
}
__device__ __forceinline__ bool scan_is_final(const TailChan& ch, int pass, int kind) {
    if (pass == 1) return true;
    return !(kind == SCAN_DC && (ch.mode == MODE_USB || ch.mode == MODE_LSB) && ch.agc);
}

struct ScanCtx {
    int kind;
    const float* x;     // input row for this channel
    double alpha, beta; // DEEMPH
    double r;           // DC radius
    float x_before;     // DC: x[-1]
};

__device__ __forceinline__ Aff scan_elem(const TailParams& p, const ScanCtx& k, int64_t r) {
    if (k.kind == SCAN_DEEMPH) {
        // z[m] = alpha*y[m] = alpha*z[m-1] + alpha*(beta*x[m])
        return Aff{k.alpha, k.alpha * (k.beta * (double)k.x[r])};
    }
    if (k.kind == SCAN_DC) {
        const float xm1 = r > 0 ? k.x[r - 1] : k.x_before;
        return Aff{k.r, (double)__fsub_rn(k.x[r], xm1)};        // common.py:24 float32 difference
    }
    // AGC (ssb.py:74-78)
    const float mag = fabsf(k.x[r]);
    double a = 1.0, b = 0.0;
    if (mag > 1e-6f) {
        a = 1.0 - p.agc_decay;
        b = p.agc_decay * (p.agc_target / (double)mag);
    }
    if (is_chunk_start(p, r)) return Aff{0.0, a + b};            // gain restarts at 1.0 (ssb.py:72)
    return Aff{a, b};
}

__device__ __forceinline__ float scan_emit(const ScanCtx& k, int64_t r, double v_prev, double v_cur) {
    if (k.kind == SCAN_DEEMPH) return (float)(k.beta * (double)k.x[r] + v_prev);   // y = b0 x + z
    if (k.kind == SCAN_DC) return (float)v_cur;
    return __fmul_rn(k.x[r], (float)v_cur);                                        // ssb.py:79
}

__device__ __forceinline__ ScanCtx make_ctx(const TailParams& p, int c, int pass, const TailChan& ch) {
    ScanCtx k;
    k.kind = scan_kind(ch, pass);
    k.alpha = ch.alpha;
    k.beta = ch.beta;
    k.r = p.dc_radius;
    k.x = (pass == 0 ? p.pre : p.tmp) + (size_t)c * p.work_stride;
    k.x_before = p.fresh ? 0.f : p.state[c].dc_x;
    return k;
}

// block-wide inclusive scan of affine maps (thread order), returns this thread's inclusive
// prefix; *total gets the block aggregate.
__device__ __forceinline__ Aff block_scan(Aff mine, Aff* total) {
    __shared__ Aff warp_tot[kScanThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Aff inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        Aff o;
        o.a = __shfl_up_sync(0xffffffffu, inc.a, off);
        o.b = __shfl_up_sync(0xffffffffu, inc.b, off);
        if (lane >= off) inc = compose(o, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    Aff pre{1.0, 0.0};
    for (int w = 0; w < wid; ++w) pre = compose(pre, warp_tot[w]);
    Aff all = pre;
    for (int w = wid; w < kScanThreads / 32; ++w) all = compose(all, warp_tot[w]);
    *total = all;
    __syncthreads();
    return compose(pre, inc);
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const TailParams p, int pass) {
    const int c = blockIdx.y;
    const TailChan ch = p.chan[c];
    const ScanCtx k = make_ctx(p, c, pass, ch);
    if (k.kind == SCAN_NONE) return;
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPer;
    Aff mine{1.0, 0.0};
#pragma unroll
    for (int i = 0; i < kScanPer; ++i)
        if (base + i < p.n) mine = compose(mine, scan_elem(p, k, base + i));
    Aff tot;
    block_scan(mine, &tot);
    if (threadIdx.x == 0) p.agg[(size_t)c * p.ntiles + blockIdx.x] = make_double2(tot.a, tot.b);
}

// one CTA per channel: block-level scan over the tile aggregates.  Each thread composes a
// contiguous run of tiles, the runs are scanned across the block, then every thread
// rewrites its run with the carry-in of each tile.
__global__ void __launch_bounds__(kScanThreads) k_scan_carry(const TailParams p, int pass) {
    const int c = blockIdx.x;
    const TailChan ch = p.chan[c];
    const int kind = scan_kind(ch, pass);
    if (kind == SCAN_NONE) return;
    double v0 = 0.0;
    if (!p.fresh) {
        if (kind == SCAN_DEEMPH) v0 = p.state[c].deemph_z;
        else if (kind == SCAN_DC) v0 = (double)p.state[c].dc_y;
    }
    if (kind == SCAN_AGC) v0 = 1.0;
    double2* agg = p.agg + (size_t)c * p.ntiles;
    const int64_t per = (p.ntiles + kScanThreads - 1) / kScanThreads;
    const int64_t t0 = (int64_t)threadIdx.x * per;
    const int64_t t1 = min(p.ntiles, t0 + per);
    Aff mine{1.0, 0.0};
    for (int64_t t = t0; t < t1; ++t) {
        const double2 ab = agg[t];
        mine = compose(mine, Aff{ab.x, ab.y});
    }
    Aff tot;
    const Aff inc = block_scan(mine, &tot);
    Aff exc;
    exc.a = __shfl_up_sync(0xffffffffu, inc.a, 1);
    exc.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
    __shared__ Aff last_of_warp[kScanThreads / 32];
    if ((threadIdx.x & 31) == 31) last_of_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) exc = threadIdx.x == 0 ? Aff{1.0, 0.0} : last_of_warp[(threadIdx.x >> 5) - 1];
    double v = fma(exc.a, v0, exc.b);
    for (int64_t t = t0; t < t1; ++t) {
        const double2 ab = agg[t];
        agg[t].x = v;                    // becomes the carry-in of tile t
        v = fma(ab.x, v, ab.y);
    }
    if (threadIdx.x == 0) {
        const double vend = fma(tot.a, v0, tot.b);
        if (kind == SCAN_DEEMPH) p.state[c].deemph_z = vend;
        else if (kind == SCAN_DC) p.state[c].dc_y = (float)vend;
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const TailParams p, int pass) {
    const int c = blockIdx.y;
    const TailChan ch = p.chan[c];
    const ScanCtx k = make_ctx(p, c, pass, ch);
    if (k.kind == SCAN_NONE) return;
    const bool fin = scan_is_final(ch, pass, k.kind);
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPer;
    Aff el[kScanPer];
    Aff mine{1.0, 0.0};
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {
        el[i] = Aff{1.0, 0.0};
        if (base + i < p.n) el[i] = scan_elem(p, k, base + i);
        mine = compose(mine, el[i]);
    }
    Aff tot;
    const Aff inc = block_scan(mine, &tot);
    // exclusive prefix of this thread = inclusive prefix of the previous thread
    Aff exc;
    exc.a = __shfl_up_sync(0xffffffffu, inc.a, 1);
    exc.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
    __shared__ Aff last_of_warp[kScanThreads / 32];
    if ((threadIdx.x & 31) == 31) last_of_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) exc = threadIdx.x == 0 ? Aff{1.0, 0.0} : last_of_warp[(threadIdx.x >> 5) - 1];
    const double carry = p.agg[(size_t)c * p.ntiles + blockIdx.x].x;
    double v = fma(exc.a, carry, exc.b);

    float peak = 0.f;
    // statistics: a thread's kScanPer consecutive rows touch at most two windows
    int64_t w_cur = -1;
    double ss_cur = 0.0;
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {
        const int64_t r = base + i;
        if (r >= p.n) break;
        const double v_new = fma(el[i].a, v, el[i].b);
        const float y = scan_emit(k, r, v, v_new);
        v = v_new;
        if (!fin) {
            p.tmp[(size_t)c * p.work_stride + r] = y;
            continue;
        }
        if (r < p.n_skip) continue;                               // warm-up rows are not emitted
        const int64_t o = r - p.n_skip;
        if (p.audio) p.audio[(size_t)c * p.out_stride + o] = y;
        if (p.clipped) p.clipped[(size_t)c * p.out_stride + o] = fminf(fmaxf(y, -0.99f), 0.99f);   // processing.py:452
        peak = fmaxf(peak, fabsf(y));
        if (p.sumsq) {
            int64_t w = ((p.mg0 + r) * (int64_t)p.decim - p.seg_origin) / p.seg_len - p.win0;
            if (w < 0) w = 0;
            if (w >= p.nwin) w = p.nwin - 1;
            if (w != w_cur) {
                if (w_cur >= 0) atomicAdd(p.sumsq + (size_t)c * p.nwin + w_cur, ss_cur);
                w_cur = w;
                ss_cur = 0.0;
            }
            ss_cur = fma((double)y, (double)y, ss_cur);
        }
    }
    if (fin && p.sumsq) {
        // warp-aggregate when the whole warp sits in one window (the common case)
        const int64_t w0 = __shfl_sync(0xffffffffu, w_cur, 0);
        const bool uniform = __all_sync(0xffffffffu, w_cur == w0);
        if (uniform) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss_cur += __shfl_xor_sync(0xffffffffu, ss_cur, off);
            if ((threadIdx.x & 31) == 0 && w0 >= 0) atomicAdd(p.sumsq + (size_t)c * p.nwin + w0, ss_cur);
        } else if (w_cur >= 0) {
            atomicAdd(p.sumsq + (size_t)c * p.nwin + w_cur, ss_cur);
        }
    }
    if (fin) {
        // running pre-clip peak (processing.py:449-451); non-negative floats order like their bit patterns
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, off));
        if ((threadIdx.x & 31) == 0 && peak > 0.f)
            atomicMax(reinterpret_cast<unsigned int*>(&p.state[c].peak), __float_as_uint(peak));
    }
}

// ---------------------------------------------------------------------------------------
// single-pass tail for constant-pole recurrences (NFM de-emphasis, AM / SSB-without-AGC DC blocker)
// ---------------------------------------------------------------------------------------
// v[m] = a v[m-1] + b[m] with |a| < 1 forgets: a^W < 1e-18 means rows more than W back cannot change v[m] in
// float64.  So a CTA that owns rows [t0, t1) starts W rows early from zero state (from the carried state when
// that reaches row 0), walks forward in 1024-row steps and is exact to rounding without any communication between
// CTAs: no aggregate pass, no carry pass, and the detector output never goes to memory.  Per step: detector
// (atan2 / hypot / real part) on 4 consecutive rows per thread, thread-local recurrence, warp + block prefix with
// the constant multipliers a^4, a^8, ..., a^128, emit.  Reads the channel samples once (plus W/len of them twice).
constexpr int kFusedThreads = 256;
constexpr int kFusedStep = kFusedThreads * 4;

__device__ __forceinline__ float detect(int mode, float2 s, float2 prev) {
    if (mode == MODE_NFM) {
        const float re = __fmaf_rn(s.x, prev.x, __fmul_rn(s.y, prev.y));       // same form as k_pre
        const float im = __fmaf_rn(s.x, -prev.y, __fmul_rn(s.y, prev.x));
        return atan2f(im, re);
    }
    if (mode == MODE_AM) return hypotf(s.x, s.y);
    return s.x;
}

__global__ void __launch_bounds__(kFusedThreads) k_tail_fused(const TailParams p) {
    __shared__ double warp_tot[kFusedThreads / 32];
    const int c = blockIdx.y;
    const TailChan ch = p.chan[c];
    const int kind = scan_kind(ch, 0);
    const bool deemph = kind == SCAN_DEEMPH;
    const double a = deemph ? ch.alpha : p.dc_radius;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // constant multipliers
    const double a2 = a * a, a4 = a2 * a2;
    double step_pw[5];                       // a^(4 * 2^k)
    step_pw[0] = a4;
#pragma unroll
    for (int k = 1; k < 5; ++k) step_pw[k] = step_pw[k - 1] * step_pw[k - 1];
    const double a128 = step_pw[4] * step_pw[4];
    double a_lane = 1.0, a_warp = 1.0;       // a^(4 lane), a^(128 wid)
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (lane & (1 << k)) a_lane *= step_pw[k];
    {
        double q = a128;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (wid & (1 << k)) a_warp *= q;
            q *= q;
        }
    }
    double a1024 = a128;
#pragma unroll
    for (int k = 0; k < 3; ++k) a1024 *= a1024;

    const int64_t t0 = (int64_t)blockIdx.x * p.fused_len;
    const int64_t t1 = min(p.n, t0 + p.fused_len);
    const int64_t w0 = max((int64_t)0, t0 - p.fused_w);
    const float2* bb = p.bb + (size_t)c * p.bb_stride;
    const iq2a_channel_state st0 = p.state[c];

    double v_in = 0.0;                       // state before row w0
    if (w0 == 0 && !p.fresh) v_in = deemph ? st0.deemph_z : (double)st0.dc_y;

    float peak = 0.f;
    int64_t win_a = 0, split = INT64_MIN;    // statistics window of the current step / first row of the next one (not located yet)
    for (int64_t sub = w0; sub < t1; sub += kFusedStep) {
        const int64_t r0 = sub + 4 * tid;
        // ---- detector on rows r0 .. r0+3 (and the row before, for the differences) ----
        float2 s[5];
        if (r0 + 3 < p.n) {
            const float4 u0 = *reinterpret_cast<const float4*>(bb + r0);
            const float4 u1 = *reinterpret_cast<const float4*>(bb + r0 + 2);
            s[1] = make_float2(u0.x, u0.y);
            s[2] = make_float2(u0.z, u0.w);
            s[3] = make_float2(u1.x, u1.y);
            s[4] = make_float2(u1.z, u1.w);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) s[1 + i] = r0 + i < p.n ? bb[r0 + i] : make_float2(0.f, 0.f);
        }
        if (r0 > 0) s[0] = r0 - 1 < p.n ? bb[r0 - 1] : make_float2(0.f, 0.f);
        else s[0] = p.fresh ? make_float2(1.f, 0.f) : make_float2(st0.prev_re, st0.prev_im);    // nfm.py:15
        float x[5];
#pragma unroll
        for (int i = 1; i < 5; ++i) x[i] = detect(ch.mode, s[i], s[i - 1]);
        double b[4];
        if (deemph) {
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = a * (ch.beta * (double)x[1 + i]);          // z = alpha z + alpha beta x
        } else {
            // DC blocker input difference in float32 (common.py:24); the row before row 0 is the carried x
            if (r0 > 0) x[0] = ch.mode == MODE_AM ? hypotf(s[0].x, s[0].y) : s[0].x;
            else x[0] = p.fresh ? 0.f : st0.dc_x;
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = (double)__fsub_rn(x[1 + i], x[i]);
        }
        // ---- prefix: state before this thread's first row ----
        double agg = b[0];
#pragma unroll
        for (int i = 1; i < 4; ++i) agg = fma(a, agg, b[i]);
        double inc = agg;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double o = __shfl_up_sync(0xffffffffu, inc, 1 << k);
            if (lane >= (1 << k)) inc = fma(step_pw[k], o, inc);
        }
        double exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 0.0;
        __syncthreads();                     // previous step's readers are done with warp_tot
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        double before_warp = 0.0, all = 0.0;
#pragma unroll
        for (int w = 0; w < kFusedThreads / 32; ++w) {
            if (w == wid) before_warp = all;
            all = fma(a128, all, warp_tot[w]);
        }
        double v = fma(a_lane, fma(a_warp, v_in, before_warp), exc);
        v_in = fma(a1024, v_in, all);        // state before the next step
        if (sub + kFusedStep <= t0) continue;                      // pure warm-up step: nothing to emit
        // ---- emit ----
        // statistics windows are reference chunks (tens of thousands of rows): a 1024-row step touches at most two.
        // Window of the step's first row and the first row of the next window: located with two 64-bit divisions
        // (~180 instructions, a quarter of this kernel when done every step) only when a step starts past the last
        // located boundary -- the steps of a CTA walk forward, so that is once per window.
        if (p.sumsq && (split == INT64_MIN || sub >= split)) {
            // window(r) = clamp(((mg0 + r) decim - origin) / seg_len - win0, 0, nwin - 1), as in k_scan_apply
            const int64_t na = (p.mg0 + sub) * (int64_t)p.decim - p.seg_origin;
            const int64_t wa = max(max(na / p.seg_len, p.win0), (int64_t)0);
            win_a = min(wa - p.win0, p.nwin - 1);
            split = INT64_MAX;
            if (win_a < p.nwin - 1)      // first row whose sample index reaches chunk wa + 1
                split = (p.seg_origin + (wa + 1) * p.seg_len + p.decim - 1) / p.decim - p.mg0;
        }
        const int64_t emit_lo = max(t0, p.n_skip);
        double ss_a = 0.0, ss_b = 0.0;
        if (r0 >= emit_lo && r0 + 3 < t1 && r0 + 3 < p.n - 1 && (r0 + 3 < split || r0 >= split)) {
            // interior rows (all four emitted, one statistics window, not the capture's last row): no per-row tests
            float y4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double v_new = fma(a, v, b[i]);
                y4[i] = deemph ? (float)(ch.beta * (double)x[1 + i] + v) : (float)v_new;
                v = v_new;
            }
            const int64_t o = r0 - p.n_skip;
            if (p.audio) {
                float* q = p.audio + (size_t)c * p.out_stride + o;
#pragma unroll
                for (int i = 0; i < 4; ++i) q[i] = y4[i];
            }
            if (p.clipped) {
                float* q = p.clipped + (size_t)c * p.out_stride + o;
#pragma unroll
                for (int i = 0; i < 4; ++i) q[i] = fminf(fmaxf(y4[i], -0.99f), 0.99f);               // processing.py:452
            }
            double ss = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                peak = fmaxf(peak, fabsf(y4[i]));
                ss += (double)y4[i] * (double)y4[i];          // same order of additions as the general loop below
            }
            if (r0 < split) ss_a = ss;
            else ss_b = ss;
        } else
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t r = r0 + i;
            const double v_new = fma(a, v, b[i]);
            const float y = deemph ? (float)(ch.beta * (double)x[1 + i] + v) : (float)v_new;
            v = v_new;
            if (r == p.n - 1) {
                // carried state after the last row; committed by k_state_tail once every CTA has read the old one
                p.agg[(size_t)c * p.ntiles] = make_double2(v_new, (double)x[1 + i]);
            }
            if (r < emit_lo || r >= t1) continue;
            const int64_t o = r - p.n_skip;
            if (p.audio) p.audio[(size_t)c * p.out_stride + o] = y;
            if (p.clipped) p.clipped[(size_t)c * p.out_stride + o] = fminf(fmaxf(y, -0.99f), 0.99f);   // processing.py:452
            peak = fmaxf(peak, fabsf(y));
            const double yy = (double)y * (double)y;
            if (r < split) ss_a += yy;
            else ss_b += yy;
        }
        if (p.sumsq) {
            const bool straddle = sub + kFusedStep > split;                            // uniform
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss_a += __shfl_xor_sync(0xffffffffu, ss_a, off);
            if (lane == 0 && ss_a != 0.0) atomicAdd(p.sumsq + (size_t)c * p.nwin + win_a, ss_a);
            if (straddle) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) ss_b += __shfl_xor_sync(0xffffffffu, ss_b, off);
                if (lane == 0 && ss_b != 0.0) atomicAdd(p.sumsq + (size_t)c * p.nwin + min(win_a + 1, p.nwin - 1), ss_b);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, off));
    if (lane == 0 && peak > 0.f) atomicMax(reinterpret_cast<unsigned int*>(&p.state[c].peak), __float_as_uint(peak));
}

// carried quantities that are plain copies of the last row
__global__ void k_state_tail(const TailParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.nchan || p.n <= 0) return;
    const int mode = p.chan[c].mode;
    if (!p.skip_pre) {
        const float2 last = p.bb[(size_t)c * p.bb_stride + p.n - 1];
        p.state[c].prev_re = last.x;
        p.state[c].prev_im = last.y;
    }
    const bool dc_mode = mode == MODE_AM || mode == MODE_USB || mode == MODE_LSB || mode == MODE_RAW_DC;
    if (p.fused_w > 0) {
        // k_tail_fused left (state after the last row, detector output of the last row) in agg[c][0]
        const double2 e = p.agg[(size_t)c * p.ntiles];
        if (dc_mode) {
            p.state[c].dc_y = (float)e.x;
            p.state[c].dc_x = (float)e.y;
        } else {
            p.state[c].deemph_z = e.x;
        }
        return;
    }
    // (channels on the sequential path update their DC-blocker state themselves, after reading it)
    if (!p.chan[c].precise && dc_mode) p.state[c].dc_x = p.pre[(size_t)c * p.work_stride + p.n - 1];
}

int64_t tail_memory_rows(double pole) {
    if (!(pole > 0.0) || !(pole < 1.0)) return pole == 0.0 ? 1024 : 0;
    const double rows = std::ceil(std::log(1e-18) / std::log(pole));
    if (rows > 65536.0) return 0;                                  // too slow a pole: use the three-kernel scan
    return ((int64_t)rows + kFusedStep - 1) / kFusedStep * kFusedStep;
}

int launch_tail(const TailParams& p_in, bool any_agc, cudaStream_t st, int64_t* launches) {
    if (p_in.n <= 0) return IQ2A_OK;
    TailParams p = p_in;
    if (p.skip_pre || any_agc) p.fused_w = 0;
    // the single-pass kernel lets a 1024-row step straddle at most one statistics-window boundary
    if (p.sumsq && p.seg_len < (int64_t)p.decim * (kFusedStep + 1)) p.fused_w = 0;
    if (p.fused_w > 0) {
        // rows per CTA: enough CTAs to fill the machine on short calls, at most 1/8 of redundant history on long ones
        int64_t len = (p.n * p.nchan / 296 + kFusedStep - 1) / kFusedStep * kFusedStep;
        len = std::max(len, p.fused_w);
        len = std::min(len, 8 * p.fused_w);
        p.fused_len = len;
        const dim3 grid((unsigned)((p.n + len - 1) / len), p.nchan);
        k_tail_fused<<<grid, kFusedThreads, 0, st>>>(p);
        k_state_tail<<<(p.nchan + 63) / 64, 64, 0, st>>>(p);
        if (launches) *launches += 2;
        IQ2A_CUDA_TRY(cudaGetLastError());
        return IQ2A_OK;
    }
    const dim3 gpre((unsigned)((p.n + 255) / 256), p.nchan);
    if (!p.skip_pre) k_pre<<<gpre, 256, 0, st>>>(p);
    const dim3 gscan((unsigned)p.ntiles, p.nchan);
    const int npass = any_agc ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
        k_scan_reduce<<<gscan, kScanThreads, 0, st>>>(p, pass);
        k_scan_carry<<<p.nchan, kScanThreads, 0, st>>>(p, pass);
        k_scan_apply<<<gscan, kScanThreads, 0, st>>>(p, pass);
    }
    k_state_tail<<<(p.nchan + 63) / 64, 64, 0, st>>>(p);
    if (launches) *launches += 2 + 3 * npass;
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

int64_t tail_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }

}  // namespace iq2a
