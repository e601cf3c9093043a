// Launchers of the stage-level / setup kernels (stage.cu) and of the channel bank (channelizer.cu).
#pragma once
#include "common.cuh"

namespace iq2a {

struct HeadParams {
    const void* raw;
    int64_t raw_n0, raw_len;
    int iq_swap, q_neg, decim;
    int64_t mg_begin;          // first channel-rate row to recompute (global index)
    const double* taps;        // all channels' taps, concatenated (device)
    const int64_t* tap_offset; // [C] (device)
    const int* ntaps;          // [C] (device)
    const double* w;           // [C] signed NCO increments (device)
    PhaseModel phase;          // tab rows indexed by global channel
    float2* out;               // [C][out_stride]
    int64_t out_stride;
    int64_t out_mg0;           // decimated index of out[.][0]
    // optional scratch [C][mixed_stride] for the mixed samples 0 .. (last row) * decim: with many channels and long
    // filters every row re-mixing its own window is 26x the work (256 channels, 53 rows: 230 M float64 sincos)
    float2* mixed;
    int64_t mixed_stride;
};

int launch_unpack_mix(const void* d_raw, int64_t n, int codec, int iq_order, double phase, double w,
                      float2* d_out, cudaStream_t st);
int launch_fir_direct(const float2* d_x, int64_t n_out, const double* d_taps, int ntaps, float2* d_y,
                      cudaStream_t st);
int launch_decimate(const float2* d_x, int64_t start, int factor, int64_t n_out, float2* d_y, cudaStream_t st);
int launch_head_direct(const HeadParams& p, int codec, int nrows, int nchan, cudaStream_t st);
// layout 0: slot order of k_channelize (j' = k1*R2 + k2 <-> bin k1 + R1*k2); layout 1: k_channelize2
// (thread j <-> bin (r>>4) + 16*(r&15) + 256*(j>>8), r = j & 255).  `scale` multiplies the table.
int launch_build_g(const double* d_taps, int ntaps, double w, int D, int M, int R1, int qn,
                   const double2* d_wtab, float2* d_gout, int cg, int c_in_group, int layout, double scale,
                   cudaStream_t st);

int launch_build_gpair(const double* d_taps, int ntaps, double w, int D, int qn, const double2* d_wtab,
                       float4* d_gout, int cg, int c_in_group, double scale, const PairGeo& geo, cudaStream_t st);
int launch_build_rot(const double* d_w, int nchan, int D, int ld, float2* d_rot, cudaStream_t st);

int launch_channelize(const ChannelizeParams& p, int m_fft, int cg, int codec, int n_sm, cudaStream_t st);
int channelize_max_group(int m_fft);
bool channelize2_available();
int launch_channelize2_cp(const ChannelizeParams& p, int cg, int n_sm, cudaStream_t st);
int launch_channelize2(const ChannelizeParams& p, int cg, const void* base, int64_t tmap_row0, int64_t rows,
                       int n_sm, cudaStream_t st);
// mirror-pair kernel (symmetric taps of one length for every channel of the launch); false: this bank cannot use it
bool pair_geometry(int ntaps, int D, PairGeo* geo);
int launch_channelize5(const ChannelizeParams& p, int cg, const PairGeo& geo, const void* base, int64_t tmap_row0,
                       int64_t rows, int n_sm, cudaStream_t st);
// many-channel form: forward transforms once per wave of block sets, then every channel group (<= 5 channels each)
int launch_channelize5_many(const ChannelizeParams& p, const PairGeo& geo, const void* base, int64_t tmap_row0, int64_t rows,
                            const SplitParams& sp, int ngroups, int cg_max, int wave_sets, int n_sm, cudaStream_t st, int64_t* launches);
size_t channelize5_scratch_bytes_per_set(const PairGeo& geo);

}  // namespace iq2a
