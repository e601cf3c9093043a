// Parameters / launchers of the bit-faithful SSB+AGC path (precise.cu).
#pragma once
#include "common.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

constexpr int kMixHist = 16;

struct MixExactParams {
    const void* raw;
    int64_t raw_n0, raw_len;
    int iq_swap, q_neg;
    int64_t n0;            // global index of mixed[0]
    int64_t count;
    PhaseModel phase;      // tab rows indexed by global channel
    int chan;
    double w;
    float2* mixed;
    // streaming: reference chunks of EARLIER calls whose samples are mixed again as filter history.  The reference
    // mixed them with their own chunk's start phase (processing.py:292-295), and tab[0] + w * (negative offset)
    // differs from that by the rounding of a ~1e5 rad product -- enough to move ~1e-3 of the complex64 LO values by
    // one ulp.  Newest last; samples older than hist_start[0] extrapolate from it.
    int nhist;
    int64_t hist_start[kMixHist];
    double hist_phase[kMixHist];
};

struct SeqChunk {
    int64_t lo, hi;        // rows of the chunk, relative to row 0 of the call
    float y_start, y_end;  // DC-blocker state the chunk started from / ended with
};

struct SeqParams {
    const float* pre;      // [C][work_stride] real(s)
    float* tmp;            // [C][work_stride] DC-blocked audio
    int64_t work_stride;
    float* audio;          // [C][out_stride] or null
    float* clipped;
    int64_t out_stride;
    double* sumsq;         // [C][nwin] or null
    int64_t nwin, win_chunk0;
    iq2a_channel_state* state;
    const int* chan_idx;   // [nprecise] channel numbers on this path (device)
    int nprecise;
    SeqChunk* rec;         // [nprecise][nchunks] (device)
    int nchunks;
    int64_t chunk0;
    int64_t seg_origin, seg_len;
    int64_t mg0, n, n_skip;
    int decim;
    int fresh;
    double agc_target, agc_decay;
    int* repaired;         // device counter of serially repaired chunks, or null
};

int launch_mix_exact(const MixExactParams& p, int codec, cudaStream_t st);
int launch_fir_decim_f64(const float2* d_mixed, const double* d_taps, int ntaps, int D, int Q, int64_t nrows,
                         float2* d_out, cudaStream_t st);
int launch_seq_tail(const SeqParams& p, cudaStream_t st, int64_t* launches);

// Transform form of the same float64 channel filter (overlap-save over the D polyphase branches of the mixed
// signal, M-point float64 transforms): ~20x fewer operations than the direct form, same rounding to complex64.
struct FirFftPlan {
    int M = 0, Q = 0, D = 0;
    double2* H = nullptr;      // [D][M] branch filter spectra / M (device); natural bin order, or row order (reg = true)
    double2* tw = nullptr;     // [M] exp(-2 pi i k / M) (device)
    // register-pass form with repair (precise_fft.cu)
    bool reg = false;
    int ntaps = 0;
    const double* taps = nullptr;   // device, owned by the bank
    int* risky = nullptr;           // rows to recompute by the direct form
    size_t risky_cap = 0;
    int* n_risky = nullptr;         // [0]: entries of `risky` in the current launch, [1]: total repaired so far
    // distance to a float32 rounding boundary below which a sample is recomputed: tol + tol_rel |v| (precise_fft.cu
    // header; IQ2A_PRECISE_TOL / IQ2A_PRECISE_TOL_REL override)
    double tol = 1e-15;
    double tol_rel = 64.0 * 1.1102230246251565e-16;
};
int fir_fft_plan_create(FirFftPlan* plan, const double* d_taps, int ntaps, int D, int Q, cudaStream_t st);
void fir_fft_plan_destroy(FirFftPlan* plan);
int launch_fir_fft64(const FirFftPlan& plan, const float2* d_mixed, int64_t nrows, float2* d_out, cudaStream_t st);
// register-pass transform form + repair of every sample near a float32 rounding boundary (precise_fft.cu): the
// direct form's result at ~1/15 of its cost
int fir_fftr_plan_create(FirFftPlan* plan, const double* d_taps, int ntaps, int D, int Q, cudaStream_t st);
int launch_fir_fft64r(FirFftPlan& plan, const float2* d_mixed, int64_t nrows, float2* d_out, cudaStream_t st);

}  // namespace iq2a
