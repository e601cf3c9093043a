// Shared declarations for the channel-bank kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace iq2a {

constexpr int kMaxGroup = 8;      // channels handled by one channelizer launch
constexpr int kBlocksPerSet = 4;  // overlap-save blocks a CTA transforms together (amortises the G-table reads)

enum Codec : int { CODEC_S16 = 0, CODEC_U8 = 1, CODEC_F32 = 2 };
enum Mode : int { MODE_NFM = 0, MODE_AM = 1, MODE_USB = 2, MODE_LSB = 3, MODE_IQ = 4,
                  // stand-alone recurrences on real input (iq2a_scan): DeemphasisFilter, DCBlocker, AGC
                  MODE_RAW_DEEMPH = 5, MODE_RAW_DC = 6, MODE_RAW_AGC = 7 };

// Piecewise-linear NCO phase, exactly the reference's bookkeeping
// (src/iq_to_audio/processing.py:292-295): inside reference chunk k the phase is
// tab[k] + w * (n - n_chunk_start), tab[k+1] = (tab[k] + w*chunk) mod 2 pi.
struct PhaseModel {
    const double* tab;   // [n_channels_in_launch][nseg]
    int64_t seg_len;     // reference chunk length in input samples (>= 1)
    int64_t seg0_n;      // global sample index where tab[.][0] starts
    int nseg;
};

struct ChannelizeParams {
    const void* raw;     // device pointer, interleaved PCM frames
    int64_t raw_n0;      // global sample index of frame 0 of `raw`
    int64_t raw_len;     // frames available
    int iq_swap;         // 1: first value of a frame is Q
    int q_neg;           // 1: negate Q
    int decim;           // D
    int vd, ld;          // overlap rows / new rows per block
    int64_t mg_begin;    // first decimated (channel-rate) index to produce
    int64_t mg_end;      // one past the last
    int nblocks;         // ceil((mg_end - mg_begin) / ld)
    int nchan;           // channels in this launch (<= CG)
    const float2* gtab;  // [D][CG][M], slot order (plan.py: spectrum_slot_to_bin)
    const float2* twid;  // [M] forward twiddles W_M^t
    const float2* rot;   // [CG][ld] e^{j w_c D r}: in-block part of the NCO rotation
    float2* out;         // [CG][out_stride] complex64 channel samples, index mg - mg_begin
    int64_t out_stride;
    PhaseModel phase;
    double w[kMaxGroup]; // signed NCO increment per channel (rad/sample)
    int* set_counter;    // generation 5: device counter, zero at launch (dynamic block-set scheduling), or null
    double phase_bias[kMaxGroup];   // constant added to the NCO phase of the output rotation (generation 5: -w (L-1)/2)
};

// Mirror-pair geometry of the generation-5 channel bank (channelizer5.cuh), shared by host and device.
struct PairGeo {
    int a;                  // A = ceil((ntaps-1)/D): window rotation of class 1 (class 2: A + 1)
    int g1, g2;             // aligned 4-column groups of class 1 (columns 0..r-1, r = 4 g1) and class 2 (r..D-1)
    int tiles1;             // tiles of class 1 = ceil(g1 / 2)
    int ntiles;             // + ceil(g2 / 2)
    int dm[2], hw[2], hl[2];   // per class: mirror row offset, rows of the wrap box, rows of a linear box
};

// Many-channel (split) form of the mirror-pair kernel (channelizer5s.cuh)
struct SplitGroup {
    int first, count;         // channels [first, first + count) of the bank
    size_t g5_off;            // float4 offset of the group's table in the bank's pair table
};

struct SplitParams {
    float4* scratch;          // [sets_in_wave][ntiles][16][256]
    int set0;                 // first block set of the wave
    int nsets;                // block sets in the wave
    int tiles_per_cta;        // k_forward5: consecutive tiles one CTA transforms
    const SplitGroup* groups; // device
    const float4* gtab5;      // the bank's pair table
    const double* w;          // [C] signed NCO increments (device)
    const double* phase_bias; // [C] (device)
    const float2* rot;        // [C][ld]
    const double* phase_tab;  // [C][nseg]
    float2* out;              // [C][out_stride]
};

// table entries of the mirror-pair kernel per channel: 4 per tile + the column left at the end of each class
inline int pair_table_entries(const PairGeo& g) { return 4 * g.ntiles + 2; }

#define IQ2A_CUDA_TRY(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::iq2a::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                              __FILE__, __LINE__);                                       \
            return IQ2A_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

void set_error(const char* fmt, ...);

// Reference NCO phase at global sample n for channel row `c` of the phase table.
// The product and the sum are rounded separately, as numpy does (no FMA).
__device__ __forceinline__ double nco_phase(const PhaseModel& pm, int c, double w, int64_t n) {
    int64_t rel = n - pm.seg0_n;
    int64_t k = rel / pm.seg_len;
    if (rel < 0) k = 0;
    if (k >= pm.nseg) k = pm.nseg - 1;
    const int64_t local = rel - k * pm.seg_len;
    return __dadd_rn(pm.tab[(int64_t)c * pm.nseg + k], __dmul_rn(w, (double)local));
}

// cos/sin of a float64 phase, returned as float (the reference rounds its
// float64 LO to complex64, processing.py:294).  Cody-Waite reduction in float64,
// then the float32 polynomial on |r| <= pi.
__device__ __forceinline__ float2 phasor_f32(double ph) {
    const double inv_2pi = 0.15915494309189533577;
    const double two_pi_hi = 6.283185307179586232;       // double(2*pi)
    const double two_pi_lo = 2.4492935982947064e-16;     // 2*pi - two_pi_hi
    const double k = rint(ph * inv_2pi);
    double r = fma(-k, two_pi_hi, ph);
    r = fma(-k, two_pi_lo, r);
    float s, c;
    sincosf((float)r, &s, &c);
    return make_float2(c, s);
}

// ---- PCM frame decoding shared by the channel bank and the stage-level kernels ----
template <int FMT>
struct RawT;
template <>
struct RawT<CODEC_S16> { using type = uint32_t; };
template <>
struct RawT<CODEC_U8> { using type = uint16_t; };
template <>
struct RawT<CODEC_F32> { using type = float2; };

template <int FMT>
__device__ __forceinline__ typename RawT<FMT>::type raw_zero() {
    if constexpr (FMT == CODEC_S16) return 0u;
    else if constexpr (FMT == CODEC_U8) return (uint16_t)0x8080u;   // (128,128) -> 0.0
    else return make_float2(0.f, 0.f);
}

// ffmpeg's sample-format rule: s16 -> x/32768, u8 -> (x-128)/128, f32 as is.
template <int FMT>
__device__ __forceinline__ float2 raw_to_c64(typename RawT<FMT>::type r, int iq_swap, int q_neg) {
    float a, b;
    if constexpr (FMT == CODEC_S16) {
        a = (float)(int16_t)(r & 0xffffu) * (1.0f / 32768.0f);
        b = (float)(int16_t)(r >> 16) * (1.0f / 32768.0f);
    } else if constexpr (FMT == CODEC_U8) {
        a = ((float)(r & 0xffu) - 128.0f) * (1.0f / 128.0f);
        b = ((float)(r >> 8) - 128.0f) * (1.0f / 128.0f);
    } else {
        a = r.x;
        b = r.y;
    }
    float i = iq_swap ? b : a;
    float q = iq_swap ? a : b;
    if (q_neg) q = -q;
    return make_float2(i, q);
}


}  // namespace iq2a
