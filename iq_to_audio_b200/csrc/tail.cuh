// Parameters of the channel-rate tail kernels (tail.cu).
#pragma once
#include "common.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

struct TailChan {
    int mode;        // Mode
    int agc;         // SSB only
    double alpha;    // de-emphasis pole (decoders/nfm.py:42)
    double beta;     // 1 - alpha
    int precise;     // 1: the bit-faithful sequential tail (precise.cu) owns this channel
    int pad;
};

struct TailParams {
    const float2* bb;          // [C][bb_stride] channel samples; row 0 <-> decimated index mg0
    int64_t bb_stride;
    float* pre;                // [C][work_stride] detector output
    float* tmp;                // [C][work_stride] DC-blocked audio ahead of the AGC
    int64_t work_stride;
    float* audio;              // [C][out_stride] or null
    float* clipped;            // [C][out_stride] or null
    int64_t out_stride;
    double2* agg;              // [C][ntiles] tile aggregates, then carry-ins
    int64_t ntiles;
    double* sumsq;             // [C][nwin] sum of squares per statistics window, or null
    int64_t nwin, win0;
    iq2a_channel_state* state; // [C] device
    const TailChan* chan;      // [C] device
    int nchan;
    int64_t n;                 // rows to process (warm-up rows included)
    int64_t n_skip;            // leading warm-up rows that are not emitted
    int64_t mg0;               // decimated index of row 0
    int decim;
    int64_t seg_origin;        // input-sample index of a reference chunk start
    int64_t seg_len;           // reference chunk length (AGC restarts, statistics windows)
    int fresh;                 // 1: start the recurrences from zero state
    int skip_pre;              // 1: `pre` already holds real float32 input (stand-alone recurrences)
    double dc_radius, agc_target, agc_decay;
    int64_t fused_w;           // > 0: every channel is a constant-pole recurrence; rows of history that make its
                               // state exact to float64 (pole^fused_w < 1e-18), multiple of 1024 -> k_tail_fused
    int64_t fused_len;         // rows per CTA on that path (set by launch_tail)
};

// rows after which a first-order recurrence with this pole has forgotten its state to below float64 resolution
int64_t tail_memory_rows(double pole);

int launch_tail(const TailParams& p, bool any_agc, cudaStream_t st, int64_t* launches);
int64_t tail_tiles(int64_t n);

}  // namespace iq2a
