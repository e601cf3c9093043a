/*
 * iq2a_b200.h -- C ABI of the B200 channelize-and-demodulate library
 * (libiq2a_b200.so, built from iq_to_audio_b200/csrc by __graft_entry__.build()).
 *
 * The reference (rknightion/iq-to-audio) is pure Python and has no FFI; its
 * boundary for this path is the Python stage API in
 * src/iq_to_audio/processing.py and src/iq_to_audio/decoders/.  Each entry point
 * below names the reference interface it stands in for.  All functions return
 * an int status (IQ2A_OK or a negative IQ2A_ERR_*); iq2a_last_error() gives the
 * message for the calling thread.  No callbacks, no torch types, plain pointers
 * and sizes only.  Pointers documented "device" must be CUDA device pointers on
 * the bank's device; everything else is host memory.
 *
 * Threading: one caller thread per bank at a time (the reference runs one
 * pipeline per thread, never concurrently: cli.py:683, interactive/workers.py:378).
 */
#ifndef IQ2A_B200_H
#define IQ2A_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IQ2A_OK            0
#define IQ2A_ERR_INVALID  (-1)  /* bad argument            -> ValueError in the Python host */
#define IQ2A_ERR_STATE    (-2)  /* call ordering / no GPU  -> RuntimeError                  */
#define IQ2A_ERR_CUDA     (-3)  /* CUDA runtime failure    -> RuntimeError                  */
#define IQ2A_ERR_NOMEM    (-4)  /* allocation failure      -> MemoryError                   */

/* input_formats.py:45-94 -- on-disk sample encodings (ffmpeg decode rule, processing.py:143-158) */
#define IQ2A_CODEC_S16 0        /* pcm_s16le: x / 32768            */
#define IQ2A_CODEC_U8  1        /* pcm_u8:    (x - 128) / 128      */
#define IQ2A_CODEC_F32 2        /* pcm_f32le / complex64 passthrough */

/* processing.py:268-279 (IQReader._extract_iq) */
#define IQ2A_ORDER_IQ     0
#define IQ2A_ORDER_QI     1
#define IQ2A_ORDER_IQ_INV 2
#define IQ2A_ORDER_QI_INV 3

/* decoders/__init__.py:9-24 (create_decoder) ; IQ = pass-through slice mode (processing.py:693-695) */
#define IQ2A_MODE_NFM 0
#define IQ2A_MODE_AM  1
#define IQ2A_MODE_USB 2
#define IQ2A_MODE_LSB 3
#define IQ2A_MODE_IQ  4

#define IQ2A_MAX_CHANNELS 256

typedef struct iq2a_bank iq2a_bank;

/* One target: what ProcessingPipeline.run derives per ProcessingConfig (processing.py:882-1002). */
typedef struct iq2a_channel_desc {
    double        freq_offset_hz;  /* target_freq - center_freq (processing.py:883)                    */
    int32_t       mix_sign;        /* +1 / -1: choose_mix_sign or --mix-sign (processing.py:1038-1042) */
    int32_t       mode;            /* IQ2A_MODE_*                                                      */
    const double* taps;            /* design_channel_filter output (processing.py:599-620), real       */
    int32_t       ntaps;
    int32_t       agc_enabled;     /* SSB only (decoders/__init__.py:16-23)                            */
    double        deemph_us;       /* NFM only (decoders/nfm.py:40-47)                                 */
} iq2a_channel_desc;

typedef struct iq2a_bank_config {
    double  sample_rate;     /* input complex sample rate, Hz                                           */
    int32_t decimation;      /* D (processing.py:885-890); common to all channels of a bank             */
    int32_t codec;           /* IQ2A_CODEC_*                                                             */
    int32_t iq_order;        /* IQ2A_ORDER_*                                                             */
    int32_t n_channels;      /* 1..IQ2A_MAX_CHANNELS                                                     */
    int32_t fft_size;        /* per-branch transform length M in {512,1024}; 0 = choose                   */
    int64_t ref_chunk;       /* the reference's chunk size in input samples (tune_chunk_size,
                                processing.py:65-81): NCO phase wrap points (:295), AGC restart points
                                (decoders/ssb.py:72) and statistics windows.  Must be >= 1.             */
    int32_t device;          /* CUDA device ordinal                                                      */
    int32_t reserved;
} iq2a_bank_config;

typedef struct iq2a_bank_info {
    int32_t fft_size;        /* M                                   */
    int32_t overlap_rows;    /* Vd: channel-rate history rows/block */
    int32_t rows_per_block;  /* Ld = M - Vd                         */
    int32_t n_channels;
    int64_t hop;             /* Ld * D input samples per block      */
    int64_t halo;            /* input samples of history a segment start needs (Vd * D) */
    double  fs_channel;
    int32_t kernel_generation; /* 5: mirror-pair kernel (symmetric taps), 4: TMA / cp.async + packed-f32x2 kernel, 1: first
                                  generation only, 0: direct mode (D = 1..2: float64 mixer + direct-form filter) */
    int32_t reserved;
} iq2a_bank_info;

/* Carried per-channel state = exactly the scalars the reference objects hold between chunks. */
typedef struct iq2a_channel_state {
    float  prev_re, prev_im;  /* QuadratureDemod.prev (decoders/nfm.py:15,23)                 */
    float  dc_x, dc_y;        /* DCBlocker._x_prev/_y_prev (decoders/common.py:12-13,28-29)   */
    double deemph_z;          /* DeemphasisFilter.state = lfilter zf[0] (decoders/nfm.py:61)  */
    float  peak;              /* AudioWriter.peak (processing.py:449-451)                     */
    float  reserved;
} iq2a_channel_state;

/* ---- library ------------------------------------------------------------------------- */
const char* iq2a_last_error(void);
int         iq2a_version(void);
int         iq2a_device_count(int32_t* count);
/* pinned host buffers for the streaming path (what IQReader.read_block's bytearray becomes) */
int         iq2a_host_alloc(void** ptr, int64_t bytes);
int         iq2a_host_free(void* ptr);

/* ---- channel bank: ComplexOscillator + OverlapSaveFIR + Decimator + Decoder + AudioWriter.write
 *      arithmetic for C targets at once (processing.py:992-1002 construct, :1088-1147 per chunk) --- */
int  iq2a_bank_create(const iq2a_bank_config* cfg, const iq2a_channel_desc* channels, iq2a_bank** out);
void iq2a_bank_destroy(iq2a_bank* bank);
int  iq2a_bank_info_get(const iq2a_bank* bank, iq2a_bank_info* info);
/* back to stream start: zero FIR history, phase 0, decimator offset 0, fresh decoder state */
int  iq2a_bank_reset(iq2a_bank* bank);
int  iq2a_bank_get_state(const iq2a_bank* bank, iq2a_channel_state* states /* [C] */, int64_t* samples_consumed);
int  iq2a_bank_set_state(iq2a_bank* bank, const iq2a_channel_state* states /* [C] */);

/*
 * Streaming step == one iteration of the reference loop body (processing.py:1070-1154) for all
 * channels: `frames` is the next n_frames of the capture in the bank's codec (host memory; pinned
 * memory from iq2a_host_alloc avoids a staging copy).  One call == one reference chunk: the NCO
 * phase wraps and the SSB AGC restarts at the start of every call.
 * Outputs (host, each [C][out_stride], any may be NULL):
 *   audio     float32 decoder output (Decoder.process()[0], pre-clip)
 *   clipped   float32 what AudioWriter.write pipes to the encoder (clip +-0.99)
 *   baseband  complex64 (re,im pairs) decimated channel samples (Decimator.process output)
 *   n_out     number of channel-rate samples produced per channel (same for all channels)
 *   rms_dbfs  [C] DecoderStats.rms_dbfs of this chunk (decoders/nfm.py:87-89)
 */
int iq2a_bank_process_chunk(iq2a_bank* bank, const void* frames, int64_t n_frames,
                            float* audio, float* clipped, float* baseband, int64_t out_stride,
                            int64_t* n_out, double* rms_dbfs);

/* The same step split in two so that two chunks can be in flight: submit queues the host->device copy
 * (copy stream) and the kernels + device->host copy of the results (compute stream) and returns;
 * collect waits for the OLDEST submitted chunk and delivers its results.  `frames` must stay valid
 * until the chunk is collected.  want: bit 0 audio, bit 1 clipped, bit 2 baseband.
 * Loop: submit(0); for k: submit(k+1); collect(k).  At most two chunks may be in flight. */
int iq2a_bank_submit_chunk(iq2a_bank* bank, const void* frames, int64_t n_frames, int32_t want);
int iq2a_bank_collect_chunk(iq2a_bank* bank, float* audio, float* clipped, float* baseband,
                            int64_t out_stride, int64_t* n_out, double* rms_dbfs);

/* Several iterations of the reference loop body (processing.py:1070-1154) in ONE call: `frames` holds
 * ceil(n_frames / chunk_frames) consecutive reference chunks of chunk_frames frames (the last may be shorter).
 * The NCO phase wraps (:295), the SSB AGC restarts (decoders/ssb.py:67) and DecoderStats are taken per reference
 * chunk exactly as if each had been submitted on its own -- only the kernels see them together, which is what fills
 * the GPU at high decimation (one 4 Mi-sample chunk of a 61.44 MS/s capture is 7 block sets for 296 CTA slots).
 * chunk_frames == 0: the call is one reference chunk (iq2a_bank_submit_chunk).
 * collect: rms_dbfs [C][rms_capacity] and window_rows [rms_capacity] receive one entry per reference chunk of the
 * call (window_rows = the Decimator.process output size of that chunk), n_windows their number. */
int iq2a_bank_submit_chunks(iq2a_bank* bank, const void* frames, int64_t n_frames, int64_t chunk_frames, int32_t want);
int iq2a_bank_collect_chunks(iq2a_bank* bank, float* audio, float* clipped, float* baseband, int64_t out_stride,
                             int64_t* n_out, double* rms_dbfs, int64_t rms_capacity, int64_t* window_rows,
                             int64_t* n_windows);

/*
 * Whole-segment step on data already resident in HBM (bench / time-sharded runs).
 * `dev_frames` holds global sample indices [first_frame, first_frame + n_frames).
 * Produces every channel-rate sample whose input index m*D lies in [seg_begin, seg_end).
 * The reference's chunk grid (multiples of cfg.ref_chunk from sample 0) defines phase wraps,
 * AGC restarts and statistics windows, so a segment reproduces the single-stream result.
 * warmup_rows > 0: the decoder recurrences (de-emphasis / DC blocker / discriminator lag) are
 * started `warmup_rows` channel-rate samples before seg_begin from zero state and those rows are
 * discarded (time-shard start; needs first_frame <= seg_begin - halo - warmup_rows*D unless that
 * is before sample 0).  warmup_rows == 0 continues from the bank's carried state.
 * Outputs are DEVICE pointers [C][out_stride] (any may be NULL); rms_dbfs is host [C][n_windows]
 * or NULL.
 */
int iq2a_bank_process_resident(iq2a_bank* bank, const void* dev_frames, int64_t first_frame,
                               int64_t n_frames, int64_t seg_begin, int64_t seg_end,
                               int32_t warmup_rows, float* dev_audio, float* dev_clipped,
                               float* dev_baseband, int64_t out_stride, int64_t* n_out,
                               double* rms_dbfs, int64_t rms_capacity);
/* same, on a caller-provided CUDA stream (cudaStream_t passed as void*), no host synchronisation */
int iq2a_bank_process_resident_async(iq2a_bank* bank, const void* dev_frames, int64_t first_frame,
                                     int64_t n_frames, int64_t seg_begin, int64_t seg_end,
                                     int32_t warmup_rows, float* dev_audio, float* dev_clipped,
                                     float* dev_baseband, int64_t out_stride, void* cuda_stream);
/* number of kernel launches issued by this bank since creation (bench.py "gpu_launches") */
int iq2a_bank_launch_count(const iq2a_bank* bank, int64_t* launches);
/* Per-kernel device timing for the roofline report: when enabled, CUDA events are recorded on the
 * launching stream around the channel-bank kernel(s), the head fix-up and the tail; get_timing
 * returns the accumulated milliseconds and the number of timed calls since set_timing. */
int iq2a_bank_set_timing(iq2a_bank* bank, int32_t enable);
int iq2a_bank_get_timing(iq2a_bank* bank, double* channelize_ms, double* head_ms, double* tail_ms,
                         int64_t* calls);
/* Multi-GPU hosts: keep `n_sm` multiprocessors free of the persistent channel-bank kernel so that a concurrent
 * collective (the NCCL gather of the previous segment's audio) can be scheduled while it runs.  0 = use all. */
int iq2a_bank_set_sm_reserve(iq2a_bank* bank, int32_t n_sm);
/* debug / tests: copy the bank's G table (complex64 [C][D][M], slot order) to host */
int iq2a_bank_copy_gtable(const iq2a_bank* bank, float* host_out, int64_t n_complex);

/* ---- stage-level entry points (drop-ins for the individual reference classes; host arrays) ---- */
/* IQReader._extract_iq + ComplexOscillator.mix (processing.py:261-297): frames -> complex64 mixed.
 * phase is the oscillator phase at frame 0; w_signed = sign * increment. */
int iq2a_unpack_mix(const void* frames, int64_t n_frames, int32_t codec, int32_t iq_order,
                    double phase, double w_signed, float* out_c64, int32_t device);
/* OverlapSaveFIR.process (processing.py:325-346): causal FIR with carried history, complex64 in/out.
 * history holds the ntaps-1 samples preceding `in` (zeros at stream start). */
int iq2a_fir(const float* in_c64, int64_t n, const float* history_c64, const double* taps,
             int32_t ntaps, float* out_c64, int32_t device);
/* Decimator.process (processing.py:354-360) */
int iq2a_decimate(const float* in_c64, int64_t n, int32_t factor, int64_t offset, float* out_c64,
                  int64_t* n_out, int32_t device);
/* Decoder.process for one channel (decoders/nfm.py:82-97, am.py:25-41, ssb.py:38-61):
 * complex64 channel samples -> float32 audio; state is read and updated. */
int iq2a_demod(int32_t mode, int32_t agc_enabled, double deemph_alpha, const float* in_c64,
               int64_t n, iq2a_channel_state* state, float* audio, double* rms_dbfs, int32_t device);

/* The scalar recurrences on their own, real float32 in/out (decoder building blocks):
 * kind 0 = DeemphasisFilter.process (decoders/nfm.py:49-62; state->deemph_z carried),
 * kind 1 = DCBlocker.process        (decoders/common.py:16-30; state->dc_x, dc_y carried),
 * kind 2 = SSBDecoder._apply_agc    (decoders/ssb.py:67-80; gain restarts at 1.0 every call). */
int iq2a_scan(int32_t kind, double deemph_alpha, const float* in, int64_t n, iq2a_channel_state* state,
              float* out, int32_t device);

/* ---- 48 kHz output stage: what the encode-side ffmpeg subprocess computes (processing.py:399-418:
 *      `-f f32le -ar round(fs_ch) -i - -acodec pcm_s16le -ar 48000`), i.e. libswresample's default
 *      resampler + flt->s16, for C channels at once.  Streaming: process() returns every output
 *      whose filter window is complete, flush() the tail (reflection, like swr_convert(NULL)). ---- */
typedef struct iq2a_resampler iq2a_resampler;
int  iq2a_resampler_create(int32_t in_rate, int32_t out_rate, int32_t n_channels, int32_t device,
                           iq2a_resampler** out);
void iq2a_resampler_destroy(iq2a_resampler* r);
/* audio: host float32 [C][in_stride], n samples per channel; pcm: host int16 [C][out_stride] */
int  iq2a_resampler_process(iq2a_resampler* r, const float* audio, int64_t n, int64_t in_stride,
                            int16_t* pcm, int64_t out_stride, int64_t* n_out);
int  iq2a_resampler_flush(iq2a_resampler* r, int16_t* pcm, int64_t out_stride, int64_t* n_out);
/* upper bound of the outputs a process()/flush() call can return after n_in more input samples */
int  iq2a_resampler_max_outputs(const iq2a_resampler* r, int64_t n_in, int64_t* n_out);

/* ---- spectrum previews: what spectrum.py computes for the interactive front end, in float64 on the device.
 *      Frames are raw interleaved PCM (codec / iq_order as for the bank); complex64 input is codec F32, order IQ.
 *      nfft must be a power of two (the front end offers 65536 ... 524288, interactive/panels.py:238). ---- */
/* compute_psd (spectrum.py:15-45): Hann window over min(n_frames, nfft) samples, zero padded to nfft, fft-shifted
 * dBFS/Hz into psd_db[nfft].  The frequency axis is fftshift(fftfreq(nfft, 1/fs)), left to the caller. */
int  iq2a_psd(const void* raw, int64_t n_frames, int32_t codec, int32_t iq_order, int32_t nfft,
              double sample_rate, double* psd_db, int32_t device);
/* streaming_waterfall (spectrum.py:54-92): push() feeds one chunk of the stream (host memory), every complete
 * nfft-window at stride hop (hop <= 0: nfft/4) is transformed, added to the running mean and appended to the
 * slice list, which is halved by pair averaging whenever it exceeds max_slices (spectrum.py:174-208).
 * result(): avg_psd_db[nfft] float64, times[n_slices] float32 (seconds), matrix[n_slices][nfft] float32;
 * any of them may be NULL.  With no complete window it fails like the reference's ValueError. */
typedef struct iq2a_spectrum iq2a_spectrum;
int  iq2a_spectrum_create(int32_t nfft, int32_t hop, int32_t max_slices, double sample_rate, int32_t codec,
                          int32_t iq_order, int32_t device, iq2a_spectrum** out);
void iq2a_spectrum_destroy(iq2a_spectrum* s);
int  iq2a_spectrum_push(iq2a_spectrum* s, const void* raw, int64_t n_frames);
int  iq2a_spectrum_counts(const iq2a_spectrum* s, int64_t* frames, int32_t* n_slices, int64_t* launches);
int  iq2a_spectrum_result(iq2a_spectrum* s, double* avg_psd_db, float* times, float* matrix);

#ifdef __cplusplus
}
#endif
#endif /* IQ2A_B200_H */
