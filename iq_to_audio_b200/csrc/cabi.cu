// C ABI of libiq2a_b200.so (include/iq2a_b200.h): bank object, device memory, launch
// sequencing.  Host-side logic only; all per-sample arithmetic lives in the kernels.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <new>
#include <vector>

#include "common.cuh"
#include "precise.cuh"
#include "stage.cuh"
#include "tail.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int frame_bytes(int codec) { return codec == CODEC_S16 ? 4 : codec == CODEC_U8 ? 2 : 8; }

// Python's float modulo (processing.py:295 uses `%` on floats): result takes the divisor's sign.
static inline double py_fmod(double a, double b) {
    double r = std::fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

template <typename T>
static int dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) return IQ2A_OK;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? IQ2A_ERR_NOMEM : IQ2A_ERR_CUDA;
    }
    return IQ2A_OK;
}
// channels per multiply-accumulate CTA of the many-channel form: the spectra of a block set are read from L2 once per
// group, so larger groups mean less traffic; 5 is what fits 128 registers without spills (6 spills 412 bytes)
constexpr int kManyGroup = 5;

template <typename T>
static int dev_grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return IQ2A_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 8 + 256;
    int rc = dev_alloc(p, want);
    if (rc != IQ2A_OK) return rc;
    *cap = want;
    return IQ2A_OK;
}

// results of one streamed chunk, in pinned host memory
struct StreamSlot {
    int64_t rows = 0, n_frames = 0;
    int64_t n0 = 0, chunk_frames = 0;   // first input sample and reference-chunk length of the call
    int nwin = 1;                       // reference chunks (statistics windows) in the call
    int want = 0;
    cudaEvent_t copied = nullptr, done = nullptr;
    bool done_valid = false;
    float* h_audio = nullptr;
    float* h_clip = nullptr;
    float2* h_bb = nullptr;
    double* h_ss = nullptr;
    size_t cap = 0;
    size_t ss_cap = 0;
    int ensure(size_t nfl, int C) {
        if (nfl <= cap && h_ss && (size_t)C <= ss_cap) return 0;
        release();
        const size_t want_n = nfl + nfl / 4 + 1024;
        if (cudaHostAlloc((void**)&h_audio, want_n * sizeof(float), cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void**)&h_clip, want_n * sizeof(float), cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void**)&h_bb, want_n * sizeof(float2), cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void**)&h_ss, (size_t)C * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
            set_error("pinned result buffer allocation failed");
            return -4;
        }
        cap = want_n;
        ss_cap = (size_t)C;
        return 0;
    }
    void release() {
        if (h_audio) cudaFreeHost(h_audio);
        if (h_clip) cudaFreeHost(h_clip);
        if (h_bb) cudaFreeHost(h_bb);
        if (h_ss) cudaFreeHost(h_ss);
        h_audio = h_clip = nullptr; h_bb = nullptr; h_ss = nullptr; cap = 0; ss_cap = 0;
    }
};

struct Group {
    int first, count;
    size_t g_off;     // offset (float2 elements) into d_gtab
    size_t g5_off;    // offset (float4 elements) into d_gtab5
    PairGeo pair;     // generation 5 geometry of this group's filter length
    bool pair_ok;     // symmetric taps of one length: the mirror-pair kernel takes the group
    bool skip;        // every channel of the group is on the bit-faithful path (precise.cu computes its samples)
};

}  // namespace iq2a

using namespace iq2a;

struct iq2a_bank {
    iq2a_bank_config cfg{};
    int C = 0, D = 1, M = 0, R1 = 32, vd = 0, ld = 0, n_sm = 148, n_sm_total = 148;
    bool any_agc = false;
    int64_t tail_fused_w = 0;   // > 0: the single-pass tail applies (tail.cuh)
    std::vector<std::vector<double>> taps;
    std::vector<int64_t> tap_off;   // offsets of each channel's taps in d_taps
    std::vector<int> modes;
    std::vector<double> w;          // signed NCO increments (sign * -2 pi f_off / fs)
    std::vector<double> phase;      // streaming: NCO phase at n_pos per channel
    std::vector<std::vector<double>> phase_tab;   // resident: per-chunk start phases (grown lazily)
    // streaming: start sample and per-channel start phase of the most recent reference chunks (oldest first), for the
    // bit-faithful mixer's view of the filter history (precise.cuh: MixExactParams::hist_*)
    std::deque<std::pair<int64_t, std::vector<double>>> phase_hist;
    std::vector<Group> groups;
    int64_t n_pos = 0;              // streaming: input samples consumed
    int64_t launches = 0;
    int64_t launches_v2 = 0;

    float2* d_gtab = nullptr;
    std::vector<int> precise;       // channels on the bit-faithful path (SSB with AGC on)
    std::vector<FirFftPlan> fir_plans;   // transform form of their float64 channel filter (M == 0: direct form)
    int* d_precise = nullptr;
    float2* d_mixed = nullptr;   size_t mixed_cap = 0;
    SeqChunk* d_rec = nullptr;   size_t rec_cap = 0;
    int* d_repaired = nullptr;
    float2* d_gtab2 = nullptr;      // layout/scale of the second-generation kernel (int16, M=512, D%4==0)
    int* d_setctr = nullptr;        // block-set counter of the generation-5 kernel (dynamic scheduling)
    // many-channel form (channelizer5s.cuh): forward transforms once per wave of block sets, groups of <= 5 channels
    bool many_ok = false;
    // low-rate captures (D = 1..2 with 1025+ taps: more history rows than any transform size holds): every channel's
    // samples come from the float64 mixer + direct-form filter of the bit-faithful path; `exact` lists the channels
    // computed that way (direct mode: all, otherwise the SSB+AGC ones)
    bool direct_mode = false;
    std::vector<int> exact;
    SplitGroup* d_split_groups = nullptr;
    double* d_phase_bias = nullptr;
    float4* d_scratch = nullptr;    size_t scratch_cap = 0;
    float4* d_gtab5 = nullptr;      // mirror-pair table of the generation-5 kernel (channelizer5.cuh)
    bool pair_ok = false;           // at least one group on the mirror-pair kernel
    bool v2_ok = false;
    bool cp_ok = false;             // generation 4 with cp.async staging is available (int16, M = 512, any D)
    int kernel_gen = 1;             // 3: warp-specialised kernel, 2: TMA + packed transforms, 1: first generation
    float2* d_tw = nullptr;
    float2* d_rot = nullptr;        // [C][ld] in-block NCO rotation table
    double* d_taps = nullptr;
    int64_t* d_tap_off = nullptr;
    int* d_ntaps = nullptr;
    double* d_w = nullptr;
    TailChan* d_chan = nullptr;
    iq2a_channel_state* d_state = nullptr;
    double* d_phase = nullptr;   size_t phase_cap = 0;
    float2* d_head_mixed = nullptr; size_t head_mixed_cap = 0;   // stream-start fix-up: mixed samples [C][vd * D]
    float2* d_bb = nullptr;      size_t bb_cap = 0;
    float* d_pre = nullptr;      size_t pre_cap = 0;
    float* d_tmp = nullptr;      size_t tmp_cap = 0;
    float* d_audio = nullptr;    size_t audio_cap = 0;
    float* d_clip = nullptr;     size_t clip_cap = 0;
    double2* d_agg = nullptr;    size_t agg_cap = 0;
    double* d_sumsq = nullptr;   size_t sumsq_cap = 0;
    unsigned char* d_ring[2] = {nullptr, nullptr};
    size_t ring_cap[2] = {0, 0};
    int ring_cur = 0;
    int64_t ring_frames = 0;        // frames currently held in d_ring[ring_cur] (all before n_pos)
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    StreamSlot slot[2];
    int64_t submitted = 0, collected = 0;
    int inflight = 0;
    // optional per-kernel timing (bench.py roofline): CUDA events on the launching stream
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // before/after channel bank, after head, after tail
    double t_chan_ms = 0, t_head_ms = 0, t_tail_ms = 0;
    int64_t t_calls = 0;
    bool ev_pending = false;

    ~iq2a_bank() {
        cudaSetDevice(cfg.device);
        void* ptrs[] = {d_split_groups, d_phase_bias, d_scratch, d_setctr, d_gtab5, d_rot, d_precise, d_mixed, d_rec, d_repaired, d_gtab, d_gtab2, d_tw, d_taps, d_tap_off, d_ntaps, d_w, d_chan, d_state, d_phase, d_head_mixed, d_bb,
                        d_pre, d_tmp, d_audio, d_clip, d_agg, d_sumsq, d_ring[0], d_ring[1]};
        for (void* q : ptrs)
            if (q) cudaFree(q);
        for (FirFftPlan& pl : fir_plans) fir_fft_plan_destroy(&pl);
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& sl : slot) {
            sl.release();
            if (sl.copied) cudaEventDestroy(sl.copied);
            if (sl.done) cudaEventDestroy(sl.done);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace iq2a {

static int fresh_state(iq2a_bank* b) {
    std::vector<iq2a_channel_state> st(b->C);
    for (auto& s : st) {
        std::memset(&s, 0, sizeof(s));
        s.prev_re = 1.0f;   // decoders/nfm.py:15
    }
    IQ2A_CUDA_TRY(cudaMemcpyAsync(b->d_state, st.data(), st.size() * sizeof(st[0]), cudaMemcpyHostToDevice, b->stream));
    IQ2A_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return IQ2A_OK;
}

// the reference's per-chunk phase bookkeeping (processing.py:295), extended lazily
static void extend_phase_table(iq2a_bank* b, int64_t nchunks) {
    const double two_pi = 2.0 * M_PI;
    for (int c = 0; c < b->C; ++c) {
        auto& t = b->phase_tab[c];
        if (t.empty()) t.push_back(0.0);
        while ((int64_t)t.size() < nchunks) {
            const double step = b->w[c] * (double)b->cfg.ref_chunk;   // sign*increment*size
            t.push_back(py_fmod(t.back() + step, two_pi));
        }
    }
}

struct CoreArgs {
    const void* d_raw;
    int64_t raw_n0, raw_len;
    int64_t raw_pad;                     // frames readable (finite garbage allowed) beyond raw_len
    int64_t mg_begin, mg_emit, mg_end;   // rows [mg_begin, mg_end) computed, [mg_emit, mg_end) emitted
    bool fresh;
    bool use_hist;                       // streaming call: earlier calls' reference chunks in bank->phase_hist
    // phase / segmentation model
    int64_t seg_origin, seg_len;
    int nseg;
    const double* h_phase;               // host [C][nseg]
    int64_t win0, nwin;                  // statistics windows (0 -> none)
    float* d_audio;
    float* d_clip;
    float* d_bb_out;
    int64_t out_stride;
    cudaStream_t st;
};

static int run_core(iq2a_bank* b, const CoreArgs& a) {
    const int C = b->C;
    const int64_t n_rows = a.mg_end - a.mg_begin;
    if (n_rows <= 0) return IQ2A_OK;
    const int64_t stride = (n_rows + 63) & ~(int64_t)63;
    int rc;
    if ((rc = dev_grow(&b->d_bb, &b->bb_cap, (size_t)C * stride))) return rc;
    if ((rc = dev_grow(&b->d_pre, &b->pre_cap, (size_t)C * stride))) return rc;
    if (b->any_agc && (rc = dev_grow(&b->d_tmp, &b->tmp_cap, (size_t)C * stride))) return rc;
    const int64_t ntiles = tail_tiles(n_rows);
    if ((rc = dev_grow(&b->d_agg, &b->agg_cap, (size_t)C * ntiles))) return rc;
    if ((rc = dev_grow(&b->d_phase, &b->phase_cap, (size_t)C * a.nseg))) return rc;
    IQ2A_CUDA_TRY(cudaMemcpyAsync(b->d_phase, a.h_phase, (size_t)C * a.nseg * sizeof(double),
                                  cudaMemcpyHostToDevice, a.st));
    if (a.nwin > 0) {
        if ((rc = dev_grow(&b->d_sumsq, &b->sumsq_cap, (size_t)C * a.nwin))) return rc;
        IQ2A_CUDA_TRY(cudaMemsetAsync(b->d_sumsq, 0, (size_t)C * a.nwin * sizeof(double), a.st));
    }

    const int order = b->cfg.iq_order;
    const int swap = (order == IQ2A_ORDER_QI || order == IQ2A_ORDER_QI_INV);
    const int neg = (order == IQ2A_ORDER_IQ_INV || order == IQ2A_ORDER_QI_INV);

    if (b->timing) IQ2A_CUDA_TRY(cudaEventRecord(b->ev[0], a.st));
    // ---- channel bank: one launch per channel group -------------------------------------
    // Second-generation kernel (TMA-staged int16, packed transforms) over every row whose input
    // rows are complete in the resident buffer; the first-generation kernel (bounds-checked loads)
    // picks up whatever is left (other codecs, M = 1024, unaligned buffers, a ragged last row).
    bool use_v2 = b->v2_ok;
    int64_t mg_split = a.mg_begin, t_row0 = 0, t_rows = 0;
    const char* t_base = nullptr;
    if (use_v2) {
        const int64_t D = b->D;
        const int64_t need_from = std::max<int64_t>(0, a.mg_begin - b->vd) * D;
        const int64_t n_base = std::max(need_from, ceil_div(a.raw_n0, D) * D);
        t_base = static_cast<const char*>(a.d_raw) + (n_base - a.raw_n0) * 4;
        t_rows = (a.raw_n0 + a.raw_len + a.raw_pad - n_base) / D;
        t_row0 = n_base / D;
        if (a.raw_n0 > need_from || (reinterpret_cast<uintptr_t>(t_base) & 15) || t_rows <= 0) use_v2 = false;
        else mg_split = std::min(a.mg_end, t_row0 + t_rows);
        if (mg_split <= a.mg_begin) use_v2 = false;
    }
    // what the tensor map cannot describe (D % 4 != 0, unaligned or partly resident buffers) goes to the same
    // kernel with cp.async staging, which bounds-checks every frame itself and so takes the whole row range
    const bool use_cp = !use_v2 && b->cp_ok;
    if (use_cp) mg_split = a.mg_end;
    else if (!use_v2) mg_split = a.mg_begin;
    const bool use_many = use_v2 && b->many_ok;
    if (use_many) {
        // every channel group in one pass: forward transforms once per wave of block sets (channelizer5s.cuh)
        ChannelizeParams p{};
        p.raw = a.d_raw;
        p.raw_n0 = a.raw_n0;
        p.raw_len = a.raw_len;
        p.iq_swap = swap;
        p.q_neg = neg;
        p.decim = b->D;
        p.vd = b->vd;
        p.ld = b->ld;
        p.twid = b->d_tw;
        p.out_stride = stride;
        p.phase.seg_len = a.seg_len;
        p.phase.seg0_n = a.seg_origin;
        p.phase.nseg = a.nseg;
        p.mg_begin = a.mg_begin;
        p.mg_end = mg_split;
        p.nblocks = (int)ceil_div(mg_split - a.mg_begin, b->ld);
        const PairGeo& geo = b->groups[0].pair;
        const size_t per_set = channelize5_scratch_bytes_per_set(geo);
        const int nsets = (p.nblocks + 1) / 2;
        const int wave = (int)std::max<size_t>(1, std::min<size_t>((size_t)nsets, ((size_t)96 << 20) / per_set));
        if ((rc = dev_grow(&b->d_scratch, &b->scratch_cap, (size_t)wave * per_set / sizeof(float4)))) return rc;
        SplitParams sp{};
        sp.scratch = b->d_scratch;
        sp.groups = b->d_split_groups;
        sp.gtab5 = b->d_gtab5;
        sp.w = b->d_w;
        sp.phase_bias = b->d_phase_bias;
        sp.rot = b->d_rot;
        sp.phase_tab = b->d_phase;
        sp.out = b->d_bb;
        if ((rc = launch_channelize5_many(p, geo, t_base, t_row0, t_rows, sp, (int)b->groups.size(), kManyGroup, wave, b->n_sm, a.st, &b->launches))) return rc;
        b->launches_v2++;
    }
    for (const Group& g : b->groups) {
        if (g.skip) continue;                        // bit-faithful channels only: precise.cu below computes them
        ChannelizeParams p{};
        p.raw = a.d_raw;
        p.raw_n0 = a.raw_n0;
        p.raw_len = a.raw_len;
        p.iq_swap = swap;
        p.q_neg = neg;
        p.decim = b->D;
        p.vd = b->vd;
        p.ld = b->ld;
        p.nchan = g.count;
        p.twid = b->d_tw;
        p.rot = b->d_rot + (size_t)g.first * b->ld;
        p.out_stride = stride;
        p.phase.tab = b->d_phase + (size_t)g.first * a.nseg;
        p.phase.seg_len = a.seg_len;
        p.phase.seg0_n = a.seg_origin;
        p.phase.nseg = a.nseg;
        for (int i = 0; i < g.count; ++i) p.w[i] = b->w[g.first + i];
        if ((use_v2 || use_cp) && !use_many) {
            p.mg_begin = a.mg_begin;
            p.mg_end = mg_split;
            p.nblocks = (int)ceil_div(mg_split - a.mg_begin, b->ld);
            p.gtab = b->d_gtab2 + g.g_off;
            p.out = b->d_bb + (size_t)g.first * stride;
            if (use_v2 && g.pair_ok) {
                // kappa^{1/2} = e^{-j w (L-1)/2} of the pair identity joins the output rotation
                const double half_len = 0.5 * (double)(b->taps[g.first].size() - 1);
                for (int i = 0; i < g.count; ++i) p.phase_bias[i] = py_fmod(-b->w[g.first + i] * half_len, 2.0 * M_PI);
                p.gtab = reinterpret_cast<const float2*>(b->d_gtab5 + g.g5_off);
                p.set_counter = b->d_setctr;
                IQ2A_CUDA_TRY(cudaMemsetAsync(b->d_setctr, 0, sizeof(int), a.st));
                rc = launch_channelize5(p, g.count, g.pair, t_base, t_row0, t_rows, b->n_sm, a.st);
            } else if (use_v2) rc = launch_channelize2(p, g.count, t_base, t_row0, t_rows, b->n_sm, a.st);
            else rc = launch_channelize2_cp(p, g.count, b->n_sm, a.st);
            if (rc) return rc;
            b->launches++;
            b->launches_v2++;
        }
        if (mg_split < a.mg_end) {
            p.mg_begin = mg_split;
            p.mg_end = a.mg_end;
            p.nblocks = (int)ceil_div(a.mg_end - mg_split, b->ld);
            p.gtab = b->d_gtab + g.g_off;
            p.out = b->d_bb + (size_t)g.first * stride + (mg_split - a.mg_begin);
            if ((rc = launch_channelize(p, b->M, g.count, b->cfg.codec, b->n_sm, a.st))) return rc;
            b->launches++;
        }
    }
    if (b->timing) IQ2A_CUDA_TRY(cudaEventRecord(b->ev[1], a.st));
    // ---- head fix-up: stream start rows recomputed in float64 ----------------------------
    if (a.mg_begin < b->vd && !b->direct_mode) {
        const int64_t last = std::min<int64_t>(a.mg_end, b->vd);
        HeadParams h{};
        h.raw = a.d_raw;
        h.raw_n0 = a.raw_n0;
        h.raw_len = a.raw_len;
        h.iq_swap = swap;
        h.q_neg = neg;
        h.decim = b->D;
        h.mg_begin = a.mg_begin;
        h.taps = b->d_taps;
        h.tap_offset = b->d_tap_off;
        h.ntaps = b->d_ntaps;
        h.w = b->d_w;
        h.phase.tab = b->d_phase;
        h.phase.seg_len = a.seg_len;
        h.phase.seg0_n = a.seg_origin;
        h.phase.nseg = a.nseg;
        h.out = b->d_bb;
        h.out_stride = stride;
        h.out_mg0 = a.mg_begin;
        // every row of a channel reads the same mixed samples: mix them once when there is enough work to matter
        // (C x rows x taps) and the scratch stays small
        const int64_t n_in = (last - 1) * (int64_t)b->D + 1;
        if ((int64_t)C * (last - a.mg_begin) >= 64 && (size_t)C * (size_t)n_in * sizeof(float2) <= ((size_t)256 << 20)) {
            if ((rc = dev_grow(&b->d_head_mixed, &b->head_mixed_cap, (size_t)C * (size_t)n_in))) return rc;
            h.mixed = b->d_head_mixed;
            h.mixed_stride = n_in;
            b->launches++;
        }
        if ((rc = launch_head_direct(h, b->cfg.codec, (int)(last - a.mg_begin), C, a.st))) return rc;
        b->launches++;
    }
    // ---- bit-faithful channel samples for SSB+AGC channels (precise.cu) ------------------------
    if (!b->exact.empty()) {
        const int D = b->D, Q = b->vd - 1;
        // rows per pass: ~300 blocks of the transform-form filter (two CTAs per SM), i.e. 0.5 GB of mixed samples
        // rows per pass: one round of the transform-form filter's CTAs (two per SM, 1024 - Q rows each) when it is in
        // use -- a second, mostly empty round costs as much as a full one -- else ~0.5 GB of mixed samples
        int64_t batch = std::max<int64_t>(1024, (64LL << 20) / D) & ~(int64_t)31;
        if (!b->fir_plans.empty() && b->fir_plans[0].reg && b->fir_plans[0].M > 0)
            batch = (int64_t)2 * b->n_sm_total * (b->fir_plans[0].M - Q);
        for (size_t pi = 0; pi < b->exact.size(); ++pi) {
            const int c = b->exact[pi];
            for (int64_t r0 = 0; r0 < n_rows; r0 += batch) {
                const int64_t r1 = std::min(n_rows, r0 + batch);
                const int64_t rows_pad = (r1 - r0 + 31) & ~(int64_t)31;
                MixExactParams m{};
                m.raw = a.d_raw;
                m.raw_n0 = a.raw_n0;
                m.raw_len = a.raw_len;
                m.iq_swap = swap;
                m.q_neg = neg;
                m.n0 = (a.mg_begin + r0 - Q) * (int64_t)D;
                m.count = (rows_pad + Q) * (int64_t)D;
                m.phase.tab = b->d_phase;
                m.phase.seg_len = a.seg_len;
                m.phase.seg0_n = a.seg_origin;
                m.phase.nseg = a.nseg;
                m.chan = c;
                m.w = b->w[c];
                m.nhist = 0;
                if (a.use_hist) {
                    // the newest kMixHist reference chunks that started before this call
                    size_t n_prev = 0;
                    while (n_prev < b->phase_hist.size() && b->phase_hist[n_prev].first < a.seg_origin) ++n_prev;
                    for (size_t i = n_prev > (size_t)kMixHist ? n_prev - kMixHist : 0; i < n_prev; ++i) {
                        m.hist_start[m.nhist] = b->phase_hist[i].first;
                        m.hist_phase[m.nhist] = b->phase_hist[i].second[c];
                        ++m.nhist;
                    }
                }
                if ((rc = dev_grow(&b->d_mixed, &b->mixed_cap, (size_t)m.count))) return rc;
                m.mixed = b->d_mixed;
                if ((rc = launch_mix_exact(m, b->cfg.codec, a.st))) return rc;
                if (pi < b->fir_plans.size() && b->fir_plans[pi].M > 0 && b->fir_plans[pi].reg)
                    rc = launch_fir_fft64r(b->fir_plans[pi], b->d_mixed, r1 - r0, b->d_bb + (size_t)c * stride + r0, a.st);
                else if (pi < b->fir_plans.size() && b->fir_plans[pi].M > 0)
                    rc = launch_fir_fft64(b->fir_plans[pi], b->d_mixed, r1 - r0, b->d_bb + (size_t)c * stride + r0, a.st);
                else
                    rc = launch_fir_decim_f64(b->d_mixed, b->d_taps + b->tap_off[c], (int)b->taps[c].size(), D, Q,
                                              r1 - r0, b->d_bb + (size_t)c * stride + r0, a.st);
                if (rc) return rc;
                b->launches += 2;
            }
        }
    }
    if (b->timing) IQ2A_CUDA_TRY(cudaEventRecord(b->ev[2], a.st));
    // ---- channel-rate tail ------------------------------------------------------------------
    TailParams t{};
    t.bb = b->d_bb;
    t.bb_stride = stride;
    t.pre = b->d_pre;
    t.tmp = b->d_tmp;
    t.work_stride = stride;
    t.audio = a.d_audio;
    t.clipped = a.d_clip;
    t.out_stride = a.out_stride;
    t.agg = b->d_agg;
    t.ntiles = ntiles;
    t.sumsq = a.nwin > 0 ? b->d_sumsq : nullptr;
    t.nwin = a.nwin;
    t.win0 = a.win0;
    t.state = b->d_state;
    t.chan = b->d_chan;
    t.nchan = C;
    t.n = n_rows;
    t.n_skip = a.mg_emit - a.mg_begin;
    t.mg0 = a.mg_begin;
    t.decim = b->D;
    t.seg_origin = a.seg_origin;
    t.seg_len = a.seg_len;
    t.fresh = a.fresh ? 1 : 0;
    t.dc_radius = 0.995;                          // decoders/common.py:9
    t.agc_target = std::pow(10.0, -12.0 / 20.0);  // decoders/ssb.py:21,33
    t.agc_decay = 0.001;                          // decoders/ssb.py:22
    t.fused_w = b->tail_fused_w;
    if ((rc = launch_tail(t, b->any_agc, a.st, &b->launches))) return rc;
    if (!b->precise.empty()) {
        if ((rc = dev_grow(&b->d_tmp, &b->tmp_cap, (size_t)C * stride))) return rc;
        SeqParams q{};
        q.pre = b->d_pre;
        q.tmp = b->d_tmp;
        q.work_stride = stride;
        q.audio = a.d_audio;
        q.clipped = a.d_clip;
        q.out_stride = a.out_stride;
        q.sumsq = t.sumsq;
        q.nwin = a.nwin;
        q.win_chunk0 = a.win0;
        q.state = b->d_state;
        q.chan_idx = b->d_precise;
        q.nprecise = (int)b->precise.size();
        q.chunk0 = 0;
        q.seg_origin = a.seg_origin;
        q.seg_len = a.seg_len;
        q.mg0 = a.mg_begin;
        q.n = n_rows;
        q.n_skip = a.mg_emit - a.mg_begin;
        q.decim = b->D;
        q.fresh = a.fresh ? 1 : 0;
        q.agc_target = t.agc_target;
        q.agc_decay = t.agc_decay;
        q.repaired = b->d_repaired;
        const int64_t last_n = (a.mg_end - 1) * (int64_t)b->D;
        q.nchunks = (int)(std::max<int64_t>(0, last_n - a.seg_origin) / a.seg_len + 1);
        if ((rc = dev_grow(&b->d_rec, &b->rec_cap, (size_t)q.nprecise * q.nchunks))) return rc;
        q.rec = b->d_rec;
        if ((rc = launch_seq_tail(q, a.st, &b->launches))) return rc;
    }
    if (b->timing) {
        IQ2A_CUDA_TRY(cudaEventRecord(b->ev[3], a.st));
        b->ev_pending = true;
    }

    if (a.d_bb_out) {
        const int64_t n_emit = a.mg_end - a.mg_emit;
        IQ2A_CUDA_TRY(cudaMemcpy2DAsync(a.d_bb_out, (size_t)a.out_stride * sizeof(float2),
                                        b->d_bb + (a.mg_emit - a.mg_begin), (size_t)stride * sizeof(float2),
                                        (size_t)n_emit * sizeof(float2), C, cudaMemcpyDeviceToDevice, a.st));
    }
    return IQ2A_OK;
}

// fold the last recorded events into the running totals (call after the stream is idle)
static void collect_timing(iq2a_bank* b) {
    if (!b->timing || !b->ev_pending) return;
    float c = 0, h = 0, t = 0;
    if (cudaEventElapsedTime(&c, b->ev[0], b->ev[1]) == cudaSuccess &&
        cudaEventElapsedTime(&h, b->ev[1], b->ev[2]) == cudaSuccess &&
        cudaEventElapsedTime(&t, b->ev[2], b->ev[3]) == cudaSuccess) {
        b->t_chan_ms += c;
        b->t_head_ms += h;
        b->t_tail_ms += t;
        b->t_calls++;
    }
    b->ev_pending = false;
}

static void stats_to_dbfs(const double* sumsq, const int64_t* counts, int64_t n, double* out) {
    // decoders/nfm.py:87-88: rms = sqrt(mean(x^2) + 1e-18); dbfs = 20 log10(rms + 1e-12)
    for (int64_t i = 0; i < n; ++i) {
        const double mean = counts[i] > 0 ? sumsq[i] / (double)counts[i] : 0.0;
        out[i] = 20.0 * std::log10(std::sqrt(mean + 1e-18) + 1e-12);
    }
}

}  // namespace iq2a

// =========================================================================================
// extern "C"
// =========================================================================================
extern "C" {

const char* iq2a_last_error(void) { return g_err; }
int iq2a_version(void) { return 100; }

int iq2a_device_count(int32_t* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s", cudaGetErrorString(e));
        if (count) *count = 0;
        return IQ2A_ERR_STATE;
    }
    if (count) *count = n;
    return IQ2A_OK;
}

int iq2a_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) { set_error("bad host_alloc arguments"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return IQ2A_OK;
}
int iq2a_host_free(void* ptr) {
    if (ptr) IQ2A_CUDA_TRY(cudaFreeHost(ptr));
    return IQ2A_OK;
}

int iq2a_bank_create(const iq2a_bank_config* cfg, const iq2a_channel_desc* ch, iq2a_bank** out) {
    if (!cfg || !ch || !out) { set_error("null argument"); return IQ2A_ERR_INVALID; }
    *out = nullptr;
    if (cfg->n_channels < 1 || cfg->n_channels > IQ2A_MAX_CHANNELS) { set_error("n_channels must be in [1, %d]", IQ2A_MAX_CHANNELS); return IQ2A_ERR_INVALID; }
    if (!(cfg->sample_rate > 0)) { set_error("sample_rate must be positive"); return IQ2A_ERR_INVALID; }
    if (cfg->decimation < 1) { set_error("decimation must be >= 1"); return IQ2A_ERR_INVALID; }
    if (cfg->ref_chunk < 1) { set_error("ref_chunk must be >= 1"); return IQ2A_ERR_INVALID; }
    if (cfg->codec < 0 || cfg->codec > 2) { set_error("unknown codec %d", cfg->codec); return IQ2A_ERR_INVALID; }
    if (cfg->iq_order < 0 || cfg->iq_order > 3) { set_error("Unsupported iq_order %d", cfg->iq_order); return IQ2A_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (the B200 path has no CPU fallback)"); return IQ2A_ERR_STATE; }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("device %d out of range", cfg->device); return IQ2A_ERR_INVALID; }
    int nt_max = 0;
    for (int c = 0; c < cfg->n_channels; ++c) {
        if (!ch[c].taps || ch[c].ntaps < 1) { set_error("channel %d: taps missing", c); return IQ2A_ERR_INVALID; }
        if (ch[c].mix_sign != 1 && ch[c].mix_sign != -1) { set_error("channel %d: mix_sign must be +1 or -1", c); return IQ2A_ERR_INVALID; }
        if (ch[c].mode < 0 || ch[c].mode > IQ2A_MODE_IQ) { set_error("channel %d: Unsupported demod mode %d", c, ch[c].mode); return IQ2A_ERR_INVALID; }
        nt_max = std::max(nt_max, ch[c].ntaps);
    }
    const int D = cfg->decimation;
    const int vd = (nt_max - 1 + D - 1) / D + 1;      // plan.py: overlap_rows
    int M = cfg->fft_size;
    bool direct = false;
    if (M == 0) {
        // cost per new sample ~ (5 log2 M + 8 C') / (1 - vd/M); 1024 only when it clearly wins -- and never for int16
        // input whose filter fits 512: the TMA / packed-f32x2 kernels (generations 4, 5) exist for M = 512 only and
        // are 2-3x faster than the first-generation kernel that would take M = 1024
        const double cm = std::min(cfg->n_channels, 6);
        auto cost = [&](int m) { return m <= vd + 16 ? 1e300 : (5.0 * std::log2((double)m) + 8.0 * cm) / (1.0 - (double)vd / m); };
        M = (cost(1024) < 0.95 * cost(512)) ? 1024 : 512;
        const char* env = std::getenv("IQ2A_CHANNELIZER");
        if (cfg->codec == IQ2A_CODEC_S16 && vd + 16 < 512 && !(env && std::strcmp(env, "v1") == 0)) M = 512;
        // no transform size holds the history (sample rates up to ~2 x fs_ch: D = 1..2 with the 1025+ taps the reference
        // always designs): direct mode.  The reference handles every rate (OverlapSaveFIR has no such limit).
        if (cost(M) >= 1e300) { direct = true; M = 512; }
    }
    if (M != 512 && M != 1024) { set_error("fft_size must be 512 or 1024"); return IQ2A_ERR_INVALID; }
    if (!direct && vd + 16 >= M) { set_error("channel filter too long for fft_size %d (%d history rows)", M, vd); return IQ2A_ERR_INVALID; }

    iq2a_bank* b = new (std::nothrow) iq2a_bank();
    if (!b) { set_error("out of host memory"); return IQ2A_ERR_NOMEM; }
    b->cfg = *cfg;
    b->C = cfg->n_channels;
    b->D = D;
    b->M = M;
    b->R1 = 32;
    b->vd = vd;
    b->ld = direct ? 256 : M - vd;
    b->direct_mode = direct;
    int rc = IQ2A_OK;
    auto fail = [&](int code) { delete b; return code; };
    if (cudaSetDevice(cfg->device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", cfg->device); return fail(IQ2A_ERR_CUDA); }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); return fail(IQ2A_ERR_CUDA); }
    b->n_sm = b->n_sm_total = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(IQ2A_ERR_CUDA); }

    const int C = b->C;
    b->taps.resize(C);
    b->modes.resize(C);
    b->w.resize(C);
    b->phase.assign(C, 0.0);
    b->phase_tab.assign(C, {});
    std::vector<TailChan> tc(C);
    std::vector<int64_t> toff(C);
    std::vector<int> tn(C);
    std::vector<double> all_taps;
    for (int c = 0; c < C; ++c) {
        b->taps[c].assign(ch[c].taps, ch[c].taps + ch[c].ntaps);
        b->modes[c] = ch[c].mode;
        // processing.py:287 and :293 -- increment = -2*pi*f/fs ; the ramp uses sign*increment
        const double inc = -2.0 * M_PI * ch[c].freq_offset_hz / cfg->sample_rate;
        b->w[c] = (double)ch[c].mix_sign * inc;
        tc[c].mode = ch[c].mode;
        tc[c].agc = ch[c].agc_enabled ? 1 : 0;
        // decoders/nfm.py:40-47
        const double fs_ch = cfg->sample_rate / D;
        const double tau = std::max(ch[c].deemph_us * 1e-6, 1e-6);
        tc[c].alpha = std::exp(-1.0 / (fs_ch * tau));
        tc[c].beta = 1.0 - tc[c].alpha;
        tc[c].precise = 0;
        tc[c].pad = 0;
        if ((ch[c].mode == IQ2A_MODE_USB || ch[c].mode == IQ2A_MODE_LSB) && ch[c].agc_enabled) {
            const char* env = std::getenv("IQ2A_PRECISE_SSB");
            const bool off = env && std::strcmp(env, "0") == 0;
            const size_t qp = (size_t)((vd - 1 + 8) & ~7);            // k_fir_decim_f64: 128 rows + padded history, 16 lanes
            const size_t fir_smem = (128 + qp) * 16 * 16 + qp * 16 * 8;
            (void)fir_smem;                                           // longer histories take k_fir_rows_f64
            if (!off) {
                tc[c].precise = 1;
                b->precise.push_back(c);
            } else {
                b->any_agc = true;       // float64-scan AGC (not bit-faithful)
            }
        }
        toff[c] = (int64_t)all_taps.size();
        tn[c] = ch[c].ntaps;
        all_taps.insert(all_taps.end(), b->taps[c].begin(), b->taps[c].end());
    }
    if (direct) for (int c = 0; c < C; ++c) b->exact.push_back(c);
    else b->exact = b->precise;
    // channel groups: runs of consecutive channels with the same filter length on the same path (fast / bit-faithful),
    // each run in groups of <= gmax channels -- so that a group can take the mirror-pair kernel with its own geometry
    // and a group of bit-faithful channels is not computed twice; too many runs: plain groups of consecutive channels
    // 25+ channels of one filter on the int16 / TMA path (5+ launches of the fused kernel): the many-channel form, whose multiply-accumulate kernel
    // takes groups of <= 5 channels (IQ2A_MANY=0 keeps the fused kernel)
    bool many = false;
    {
        const char* env = std::getenv("IQ2A_CHANNELIZER");
        const char* em = std::getenv("IQ2A_MANY");
        many = C >= 25 && M == 512 && cfg->codec == IQ2A_CODEC_S16 && D % 4 == 0 && channelize2_available() && !env &&
               !(em && std::strcmp(em, "0") == 0) && b->precise.empty();
        for (int c = 1; c < C && many; ++c) many = tn[c] == tn[0];
        PairGeo probe{};
        many = many && pair_geometry(tn[0], D, &probe);
    }
    const int gmax = many ? kManyGroup : channelize_max_group(M);
    size_t g_total = 0;
    {
        std::vector<std::pair<int, int>> runs;                 // (first, count)
        for (int c = 0; c < C; ++c) {
            if (!runs.empty() && tn[c] == tn[c - 1] && tc[c].precise == tc[c - 1].precise) runs.back().second++;
            else runs.push_back({c, 1});
        }
        if ((int)runs.size() > 8) runs.assign(1, {0, C});
        if (direct) runs.clear();
        for (const auto& run : runs) {
            const int ng = (run.second + gmax - 1) / gmax;
            for (int g = 0, first = run.first; g < ng; ++g) {
                const int count = run.second / ng + (g < run.second % ng ? 1 : 0);
                Group grp{};
                grp.first = first;
                grp.count = count;
                grp.g_off = g_total;
                grp.skip = true;
                for (int i = 0; i < count; ++i) grp.skip = grp.skip && tc[first + i].precise;
                b->groups.push_back(grp);
                g_total += (size_t)D * count * M;
                first += count;
            }
        }
    }
    // device tables
    std::vector<double2> wtab(M);
    std::vector<float2> twf(M);
    for (int t = 0; t < M; ++t) {
        const double ang = -2.0 * M_PI * (double)t / (double)M;
        wtab[t] = make_double2(std::cos(ang), std::sin(ang));
        twf[t] = make_float2((float)wtab[t].x, (float)wtab[t].y);
    }
    double2* d_wtab = nullptr;
    if ((rc = dev_alloc(&b->d_gtab, std::max<size_t>(g_total, 1))) || (rc = dev_alloc(&b->d_tw, (size_t)M)) ||
        (rc = dev_alloc(&b->d_taps, all_taps.size())) || (rc = dev_alloc(&b->d_tap_off, (size_t)C)) ||
        (rc = dev_alloc(&b->d_ntaps, (size_t)C)) || (rc = dev_alloc(&b->d_w, (size_t)C)) ||
        (rc = dev_alloc(&b->d_chan, (size_t)C)) || (rc = dev_alloc(&b->d_state, (size_t)C)) ||
        (rc = dev_alloc(&d_wtab, (size_t)M)))
        return fail(rc);
    bool ok = true;
    ok &= cudaMemcpy(b->d_tw, twf.data(), M * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(d_wtab, wtab.data(), M * sizeof(double2), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(b->d_taps, all_taps.data(), all_taps.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(b->d_tap_off, toff.data(), C * sizeof(int64_t), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(b->d_ntaps, tn.data(), C * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(b->d_w, b->w.data(), C * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(b->d_chan, tc.data(), C * sizeof(TailChan), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!b->precise.empty()) {
        if ((rc = dev_alloc(&b->d_precise, b->precise.size())) || (rc = dev_alloc(&b->d_repaired, (size_t)1))) { cudaFree(d_wtab); return fail(rc); }
        ok &= cudaMemcpy(b->d_precise, b->precise.data(), b->precise.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemset(b->d_repaired, 0, sizeof(int)) == cudaSuccess;
    }
    if (!b->exact.empty()) {
        // float64 channel filter of these channels: "direct" = precise.cu's direct form (2 * ntaps DFMAs per sample),
        // "fft" = its first transform form (1.3e-4 of the samples round to the neighbouring complex64), default = the
        // register-pass transform form with repair of every sample near a rounding boundary (precise_fft.cu): the
        // direct form's result, bit for bit, at a fraction of its cost
        const char* env = std::getenv("IQ2A_PRECISE_FIR");
        const bool want_direct = env && std::strcmp(env, "direct") == 0;
        const bool want_old_fft = env && std::strcmp(env, "fft") == 0;
        if (!want_direct) {
            b->fir_plans.resize(b->exact.size());
            for (size_t pi = 0; pi < b->exact.size(); ++pi) {
                const int c = b->exact[pi];
                rc = want_old_fft ? fir_fft_plan_create(&b->fir_plans[pi], b->d_taps + toff[c], tn[c], D, vd - 1, b->stream)
                                  : fir_fftr_plan_create(&b->fir_plans[pi], b->d_taps + toff[c], tn[c], D, vd - 1, b->stream);
                if (rc) { cudaFree(d_wtab); return fail(rc); }
                b->launches += 2;
            }
        }
    }
    if (!ok) { cudaFree(d_wtab); set_error("table upload failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(IQ2A_ERR_CUDA); }
    {
        // single-pass tail: every channel a constant-pole recurrence (NFM de-emphasis; AM / SSB-without-AGC DC blocker)
        const char* env = std::getenv("IQ2A_TAIL");
        int64_t w = (env && std::strcmp(env, "v1") == 0) ? 0 : 1;
        for (int c = 0; c < C && w > 0; ++c) {
            const int m = tc[c].mode;
            int64_t rows = 0;
            if (tc[c].precise || b->any_agc) rows = 0;
            else if (m == IQ2A_MODE_NFM) rows = tail_memory_rows(tc[c].alpha);
            else if (m == IQ2A_MODE_AM || m == IQ2A_MODE_USB || m == IQ2A_MODE_LSB) rows = tail_memory_rows(0.995);
            w = rows > 0 ? std::max(w, rows) : 0;
        }
        b->tail_fused_w = w;
    }
    {
        const char* env = std::getenv("IQ2A_CHANNELIZER");
        const bool force_v1 = env && std::strcmp(env, "v1") == 0;
        const bool force_v4 = env && std::strcmp(env, "v4") == 0;
        const bool bulk = !force_v1 && M == 512 && cfg->codec == IQ2A_CODEC_S16;
        const char* stg = std::getenv("IQ2A_STAGING");               // "cp": cp.async staging even where TMA applies
        b->v2_ok = bulk && D % 4 == 0 && channelize2_available() && !(stg && std::strcmp(stg, "cp") == 0);
        b->cp_ok = bulk;
        // generation 5 (mirror pairs), per group: every channel the same number of taps, each filter exactly symmetric
        // (firwin's are, processing.py:613-619), and a rotation that fits the staging geometry
        size_t g5_total = 0;
        for (Group& g : b->groups) {
            bool sym = b->v2_ok && !force_v4 && !g.skip;
            for (int i = 0; i < g.count && sym; ++i) {
                const int c = g.first + i;
                sym = tn[c] == tn[g.first];
                const std::vector<double>& h = b->taps[c];
                for (size_t k = 0, n = h.size(); k < n / 2 && sym; ++k) sym = h[k] == h[n - 1 - k];
            }
            g.pair_ok = sym && pair_geometry(tn[g.first], D, &g.pair);
            if (g.pair_ok) {
                g.g5_off = g5_total;
                g5_total += (size_t)pair_table_entries(g.pair) * g.count * 256;
                b->pair_ok = true;
            }
        }
        b->kernel_gen = direct ? 0 : (b->pair_ok ? 5 : ((b->v2_ok || b->cp_ok) ? 4 : 1));
        b->many_ok = many;
        for (const Group& g : b->groups) b->many_ok = b->many_ok && g.pair_ok;
        if (b->many_ok) {
            std::vector<SplitGroup> sg;
            for (const Group& g : b->groups) sg.push_back(SplitGroup{g.first, g.count, g.g5_off});
            std::vector<double> bias(C);
            const double half_len = 0.5 * (double)(tn[0] - 1);
            for (int c = 0; c < C; ++c) bias[c] = py_fmod(-b->w[c] * half_len, 2.0 * M_PI);
            if ((rc = dev_alloc(&b->d_split_groups, sg.size())) || (rc = dev_alloc(&b->d_phase_bias, (size_t)C))) { cudaFree(d_wtab); return fail(rc); }
            if (cudaMemcpy(b->d_split_groups, sg.data(), sg.size() * sizeof(SplitGroup), cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(b->d_phase_bias, bias.data(), C * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFree(d_wtab);
                set_error("table upload failed");
                return fail(IQ2A_ERR_CUDA);
            }
        }
        if ((b->v2_ok || b->cp_ok) && (rc = dev_alloc(&b->d_gtab2, g_total))) { cudaFree(d_wtab); return fail(rc); }
        if (b->pair_ok && ((rc = dev_alloc(&b->d_gtab5, g5_total)) || (rc = dev_alloc(&b->d_setctr, (size_t)1)))) { cudaFree(d_wtab); return fail(rc); }
    }
    for (const Group& g : b->groups)
        for (int i = 0; i < g.count; ++i) {
            const int c = g.first + i;
            rc = launch_build_g(b->d_taps + toff[c], tn[c], b->w[c], D, M, b->R1, vd, d_wtab,
                                b->d_gtab + g.g_off, g.count, i, 0, 1.0, b->stream);
            if (!rc && (b->v2_ok || b->cp_ok))
                rc = launch_build_g(b->d_taps + toff[c], tn[c], b->w[c], D, M, b->R1, vd, d_wtab,
                                    b->d_gtab2 + g.g_off, g.count, i, 2, 1.0 / 32768.0, b->stream);
            if (!rc && g.pair_ok)
                rc = launch_build_gpair(b->d_taps + toff[c], tn[c], b->w[c], D, vd, d_wtab, b->d_gtab5 + g.g5_off,
                                        g.count, i, 1.0 / 32768.0, g.pair, b->stream);
            if (rc) { cudaFree(d_wtab); return fail(rc); }
            b->launches += 1 + ((b->v2_ok || b->cp_ok) ? 1 : 0) + (g.pair_ok ? 1 : 0);
        }
    b->tap_off = toff;
    if ((rc = dev_alloc(&b->d_rot, (size_t)C * b->ld)) || (rc = launch_build_rot(b->d_w, C, D, b->ld, b->d_rot, b->stream))) {
        cudaFree(d_wtab);
        return fail(rc);
    }
    b->launches++;
    cudaError_t e = cudaStreamSynchronize(b->stream);
    cudaFree(d_wtab);
    if (e != cudaSuccess) { set_error("G-table build failed: %s", cudaGetErrorString(e)); return fail(IQ2A_ERR_CUDA); }
    if ((rc = fresh_state(b))) return fail(rc);
    *out = b;
    return IQ2A_OK;
}

void iq2a_bank_destroy(iq2a_bank* bank) { delete bank; }

int iq2a_bank_info_get(const iq2a_bank* b, iq2a_bank_info* info) {
    if (!b || !info) { set_error("null argument"); return IQ2A_ERR_INVALID; }
    info->fft_size = b->M;
    info->overlap_rows = b->vd;
    info->rows_per_block = b->ld;
    info->n_channels = b->C;
    info->hop = (int64_t)b->ld * b->D;
    info->halo = (int64_t)b->vd * b->D;
    info->fs_channel = b->cfg.sample_rate / b->D;
    info->kernel_generation = b->kernel_gen;
    info->reserved = 0;
    return IQ2A_OK;
}

int iq2a_bank_reset(iq2a_bank* b) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    if (b->inflight != 0) { set_error("reset while chunks are in flight"); return IQ2A_ERR_STATE; }
    IQ2A_CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (b->copy_stream) IQ2A_CUDA_TRY(cudaStreamSynchronize(b->copy_stream));
    b->n_pos = 0;
    b->ring_frames = 0;
    for (auto& sl : b->slot) sl.done_valid = false;
    std::fill(b->phase.begin(), b->phase.end(), 0.0);
    b->phase_hist.clear();
    return fresh_state(b);
}

int iq2a_bank_get_state(const iq2a_bank* b, iq2a_channel_state* states, int64_t* consumed) {
    if (!b || !states) { set_error("null argument"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    IQ2A_CUDA_TRY(cudaStreamSynchronize(b->stream));
    IQ2A_CUDA_TRY(cudaMemcpy(states, b->d_state, b->C * sizeof(iq2a_channel_state), cudaMemcpyDeviceToHost));
    if (consumed) *consumed = b->n_pos;
    return IQ2A_OK;
}

int iq2a_bank_set_state(iq2a_bank* b, const iq2a_channel_state* states) {
    if (!b || !states) { set_error("null argument"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    IQ2A_CUDA_TRY(cudaMemcpy(b->d_state, states, b->C * sizeof(iq2a_channel_state), cudaMemcpyHostToDevice));
    return IQ2A_OK;
}

int iq2a_bank_launch_count(const iq2a_bank* b, int64_t* launches) {
    if (!b || !launches) { set_error("null argument"); return IQ2A_ERR_INVALID; }
    *launches = b->launches;
    return IQ2A_OK;
}

int iq2a_bank_set_sm_reserve(iq2a_bank* b, int32_t n_sm) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    if (n_sm < 0 || n_sm >= b->n_sm_total) { set_error("sm reserve %d out of range [0, %d)", n_sm, b->n_sm_total); return IQ2A_ERR_INVALID; }
    b->n_sm = b->n_sm_total - n_sm;
    return IQ2A_OK;
}

int iq2a_bank_set_timing(iq2a_bank* b, int32_t enable) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    if (enable && !b->ev[0])
        for (auto& e : b->ev) IQ2A_CUDA_TRY(cudaEventCreate(&e));
    b->timing = enable != 0;
    b->t_chan_ms = b->t_head_ms = b->t_tail_ms = 0;
    b->t_calls = 0;
    b->ev_pending = false;
    return IQ2A_OK;
}

int iq2a_bank_get_timing(iq2a_bank* b, double* channelize_ms, double* head_ms, double* tail_ms, int64_t* calls) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    if (b->ev_pending) {
        IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
        IQ2A_CUDA_TRY(cudaEventSynchronize(b->ev[3]));
        collect_timing(b);
    }
    if (channelize_ms) *channelize_ms = b->t_chan_ms;
    if (head_ms) *head_ms = b->t_head_ms;
    if (tail_ms) *tail_ms = b->t_tail_ms;
    if (calls) *calls = b->t_calls;
    return IQ2A_OK;
}

int iq2a_bank_copy_gtable(const iq2a_bank* b, float* host_out, int64_t n_complex) {
    if (!b || !host_out) { set_error("null argument"); return IQ2A_ERR_INVALID; }
    const size_t per = (size_t)b->D * b->M;
    if ((size_t)n_complex < per * b->C) { set_error("gtable buffer too small"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    // device layout is [group][D][cg][M]; hand back [C][D][M]
    std::vector<float2> tmp;
    for (const Group& g : b->groups) {
        tmp.resize(per * g.count);
        IQ2A_CUDA_TRY(cudaMemcpy(tmp.data(), b->d_gtab + g.g_off, tmp.size() * sizeof(float2), cudaMemcpyDeviceToHost));
        for (int i = 0; i < g.count; ++i)
            for (int p = 0; p < b->D; ++p)
                std::memcpy(host_out + 2 * (((size_t)(g.first + i) * b->D + p) * b->M),
                            tmp.data() + ((size_t)p * g.count + i) * b->M, b->M * sizeof(float2));
    }
    return IQ2A_OK;
}

// ---- streaming: two chunks in flight.  submit() queues the H2D copy on the copy stream and the
// kernels + D2H of the results on the compute stream; collect() waits for the oldest chunk and hands
// its results over.  H2D of chunk k+1 overlaps the kernels and the D2H of chunk k.
int iq2a_bank_submit_chunk(iq2a_bank* b, const void* frames, int64_t n_frames, int32_t want) {
    return iq2a_bank_submit_chunks(b, frames, n_frames, 0, want);
}

int iq2a_bank_submit_chunks(iq2a_bank* b, const void* frames, int64_t n_frames, int64_t chunk_frames, int32_t want) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    if (n_frames < 0 || (n_frames > 0 && !frames) || chunk_frames < 0) { set_error("bad frame buffer"); return IQ2A_ERR_INVALID; }
    if (b->inflight >= 2) { set_error("two chunks already in flight: collect one first"); return IQ2A_ERR_STATE; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    const int C = b->C, D = b->D;
    const int fb = frame_bytes(b->cfg.codec);
    const int si = (int)(b->submitted & 1);
    StreamSlot& sl = b->slot[si];
    if (!sl.copied) {
        IQ2A_CUDA_TRY(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
        IQ2A_CUDA_TRY(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    if (!b->copy_stream) IQ2A_CUDA_TRY(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    const int64_t mg_begin = ceil_div(b->n_pos, D);
    const int64_t mg_end = ceil_div(b->n_pos + n_frames, D);
    const int64_t rows = mg_end - mg_begin;
    // reference chunks covered by this call: the NCO phase wraps, the AGC restarts and the statistics windows begin at
    // every multiple of chunk_frames from the call's first sample (chunk_frames == 0: the whole call is one chunk)
    const int64_t cf = chunk_frames > 0 ? chunk_frames : std::max<int64_t>(n_frames, 1);
    const int nseg = (int)std::max<int64_t>(1, ceil_div(n_frames, cf));
    sl.rows = rows;
    sl.n_frames = n_frames;
    sl.n0 = b->n_pos;
    sl.chunk_frames = cf;
    sl.nwin = nseg;
    sl.want = want;
    b->submitted++;
    b->inflight++;
    if (n_frames == 0) return IQ2A_OK;                 // every stage returns its (empty) input

    // stream buffer: [history | new chunk]; history = the (vd+1)*D frames before n_pos
    const int64_t want_hist = std::min<int64_t>((int64_t)(b->vd + 1) * D, b->n_pos);
    int64_t keep = std::min(want_hist, b->ring_frames);
    // first ring frame on a 4-frame boundary of the global index (16 B for int16 -> TMA-addressable)
    keep = std::max<int64_t>(0, b->n_pos - (((b->n_pos - keep) + 3) & ~(int64_t)3));
    const int nxt = b->ring_cur ^ 1;
    // the buffer about to be overwritten was read by the chunk before the previous one
    StreamSlot& prev2 = b->slot[si];   // same parity == chunk k-2, already collected (inflight < 2) -> its kernels are done
    (void)prev2;
    // the chunk still in flight (k-1) reads ring[ring_cur]; we only read its tail (history) -> no hazard
    // D extra frames after the chunk: the bulk kernel reads whole rows of D frames (values beyond the
    // data only ever meet zero taps, but must be readable)
    const size_t need_bytes = (size_t)(keep + n_frames + D) * fb + 16;
    if (need_bytes > b->ring_cap[nxt]) {
        // growing frees the old buffer: make sure nothing queued still uses it
        IQ2A_CUDA_TRY(cudaStreamSynchronize(b->stream));
        IQ2A_CUDA_TRY(cudaStreamSynchronize(b->copy_stream));
    }
    int rc = dev_grow(&b->d_ring[nxt], &b->ring_cap[nxt], need_bytes);
    if (rc) return rc;
    // ring[nxt] was last read by chunk k-2's kernels: order the copy stream after them
    if (b->slot[si].done_valid) IQ2A_CUDA_TRY(cudaStreamWaitEvent(b->copy_stream, b->slot[si].done, 0));
    if (keep > 0)
        IQ2A_CUDA_TRY(cudaMemcpyAsync(b->d_ring[nxt], b->d_ring[b->ring_cur] + (size_t)(b->ring_frames - keep) * fb,
                                      (size_t)keep * fb, cudaMemcpyDeviceToDevice, b->copy_stream));
    IQ2A_CUDA_TRY(cudaMemcpyAsync(b->d_ring[nxt] + (size_t)keep * fb, frames, (size_t)n_frames * fb,
                                  cudaMemcpyHostToDevice, b->copy_stream));
    // the row of slack after the data only ever meets zero taps, but in floating point a stale value
    // would still leave a rounding-level trace in the last block: keep it zero -> bit-reproducible
    IQ2A_CUDA_TRY(cudaMemsetAsync(b->d_ring[nxt] + (size_t)(keep + n_frames) * fb, 0, (size_t)D * fb, b->copy_stream));
    IQ2A_CUDA_TRY(cudaEventRecord(sl.copied, b->copy_stream));
    IQ2A_CUDA_TRY(cudaStreamWaitEvent(b->stream, sl.copied, 0));
    b->ring_cur = nxt;
    b->ring_frames = keep + n_frames;

    const size_t r1 = (size_t)std::max<int64_t>(rows, 1);
    if (C * r1 > b->audio_cap || C * r1 > b->clip_cap) IQ2A_CUDA_TRY(cudaStreamSynchronize(b->stream));
    if ((rc = dev_grow(&b->d_audio, &b->audio_cap, C * r1))) return rc;
    if ((rc = dev_grow(&b->d_clip, &b->clip_cap, C * r1))) return rc;

    CoreArgs a{};
    a.d_raw = b->d_ring[nxt];
    a.raw_n0 = b->n_pos - keep;
    a.raw_len = keep + n_frames;
    a.raw_pad = (b->cfg.codec == IQ2A_CODEC_S16) ? D : 0;
    a.mg_begin = a.mg_emit = mg_begin;
    a.mg_end = mg_end;
    a.fresh = false;
    // the reference's phase carry (processing.py:295) chunk by chunk: start phase of every reference chunk of the call
    std::vector<double> hp((size_t)C * nseg);
    for (int c = 0; c < C; ++c) {
        double ph = b->phase[c];
        for (int k = 0; k < nseg; ++k) {
            hp[(size_t)c * nseg + k] = ph;
            const int64_t len = std::min<int64_t>(cf, n_frames - (int64_t)k * cf);
            ph = py_fmod(ph + b->w[c] * (double)len, 2.0 * M_PI);
        }
        b->phase[c] = ph;
    }
    for (int k = 0; k < nseg; ++k) {
        std::vector<double> ph(C);
        for (int c = 0; c < C; ++c) ph[c] = hp[(size_t)c * nseg + k];
        b->phase_hist.emplace_back(b->n_pos + (int64_t)k * cf, std::move(ph));
    }
    while ((int)b->phase_hist.size() > 2 * kMixHist) b->phase_hist.pop_front();
    a.use_hist = true;
    a.seg_origin = b->n_pos;
    a.seg_len = cf;
    a.nseg = nseg;
    a.h_phase = hp.data();            // staged by the runtime before cudaMemcpyAsync returns (pageable source)
    a.win0 = 0;
    a.nwin = rows > 0 ? nseg : 0;
    a.d_audio = b->d_audio;
    a.d_clip = b->d_clip;
    a.d_bb_out = nullptr;
    a.out_stride = (int64_t)r1;
    a.st = b->stream;
    if (rows > 0 && (rc = run_core(b, a))) return rc;

    b->n_pos += n_frames;

    if (rows > 0) {
        // results -> pinned slot buffers (D2H on the compute stream, behind the kernels)
        const size_t nfl = (size_t)C * rows;
        if ((rc = sl.ensure(nfl, C * nseg))) return rc;
        if (want & 1)
            IQ2A_CUDA_TRY(cudaMemcpyAsync(sl.h_audio, b->d_audio, nfl * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
        if (want & 2)
            IQ2A_CUDA_TRY(cudaMemcpyAsync(sl.h_clip, b->d_clip, nfl * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
        if (want & 4) {
            const int64_t stride = (rows + 63) & ~(int64_t)63;
            IQ2A_CUDA_TRY(cudaMemcpy2DAsync(sl.h_bb, rows * sizeof(float2), b->d_bb, stride * sizeof(float2),
                                            rows * sizeof(float2), C, cudaMemcpyDeviceToHost, b->stream));
        }
        IQ2A_CUDA_TRY(cudaMemcpyAsync(sl.h_ss, b->d_sumsq, (size_t)C * nseg * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    }
    IQ2A_CUDA_TRY(cudaEventRecord(sl.done, b->stream));
    sl.done_valid = true;
    return IQ2A_OK;
}

int iq2a_bank_collect_chunk(iq2a_bank* b, float* audio, float* clipped, float* baseband, int64_t out_stride,
                            int64_t* n_out, double* rms_dbfs) {
    if (b && b->inflight > 0 && b->slot[(int)(b->collected & 1)].nwin > 1 && rms_dbfs) {
        set_error("the oldest call covers %d reference chunks: collect it with iq2a_bank_collect_chunks", b->slot[(int)(b->collected & 1)].nwin);
        return IQ2A_ERR_STATE;
    }
    return iq2a_bank_collect_chunks(b, audio, clipped, baseband, out_stride, n_out, rms_dbfs, 1, nullptr, nullptr);
}

int iq2a_bank_collect_chunks(iq2a_bank* b, float* audio, float* clipped, float* baseband, int64_t out_stride,
                             int64_t* n_out, double* rms_dbfs, int64_t rms_capacity, int64_t* window_rows,
                             int64_t* n_windows) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    if (b->inflight <= 0) { set_error("no chunk in flight"); return IQ2A_ERR_STATE; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    StreamSlot& sl = b->slot[(int)(b->collected & 1)];
    const int C = b->C;
    const int64_t rows = sl.rows;
    const int nwin = sl.n_frames > 0 ? sl.nwin : 0;
    if ((rms_dbfs || window_rows) && rms_capacity < nwin) { set_error("statistics buffers hold %lld windows, the call has %d", (long long)rms_capacity, nwin); return IQ2A_ERR_INVALID; }
    if (rows > out_stride && (audio || clipped || baseband)) { set_error("output stride %lld < %lld rows", (long long)out_stride, (long long)rows); return IQ2A_ERR_INVALID; }
    b->collected++;
    b->inflight--;
    if (n_out) *n_out = rows;
    if (n_windows) *n_windows = nwin;
    if (sl.n_frames == 0) return IQ2A_OK;
    IQ2A_CUDA_TRY(cudaEventSynchronize(sl.done));
    collect_timing(b);
    if (rows > 0) {
        for (int c = 0; c < C; ++c) {
            if (audio && (sl.want & 1)) std::memcpy(audio + (size_t)c * out_stride, sl.h_audio + (size_t)c * rows, rows * sizeof(float));
            if (clipped && (sl.want & 2)) std::memcpy(clipped + (size_t)c * out_stride, sl.h_clip + (size_t)c * rows, rows * sizeof(float));
            if (baseband && (sl.want & 4)) std::memcpy(baseband + 2 * (size_t)c * out_stride, sl.h_bb + (size_t)c * rows, rows * sizeof(float2));
        }
    }
    // rows of every reference chunk of the call (Decimator.process output sizes) and its DecoderStats.rms_dbfs
    std::vector<int64_t> cnt(std::max(nwin, 1), 0);
    for (int k = 0; k < nwin; ++k) {
        const int64_t lo = sl.n0 + (int64_t)k * sl.chunk_frames, hi = std::min(lo + sl.chunk_frames, sl.n0 + sl.n_frames);
        cnt[k] = ceil_div(hi, b->D) - ceil_div(lo, b->D);
        if (window_rows) window_rows[k] = cnt[k];
    }
    if (rms_dbfs) {
        std::vector<double> ss((size_t)C * std::max(nwin, 1), 0.0);
        if (rows > 0) std::memcpy(ss.data(), sl.h_ss, (size_t)C * nwin * sizeof(double));
        for (int c = 0; c < C; ++c) stats_to_dbfs(ss.data() + (size_t)c * nwin, cnt.data(), nwin, rms_dbfs + (size_t)c * rms_capacity);
    }
    return IQ2A_OK;
}

int iq2a_bank_process_chunk(iq2a_bank* b, const void* frames, int64_t n_frames, float* audio, float* clipped,
                            float* baseband, int64_t out_stride, int64_t* n_out, double* rms_dbfs) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    if (b->inflight != 0) { set_error("process_chunk while submitted chunks are in flight"); return IQ2A_ERR_STATE; }
    const int want = (audio ? 1 : 0) | (clipped ? 2 : 0) | (baseband ? 4 : 0);
    int rc = iq2a_bank_submit_chunk(b, frames, n_frames, want);
    if (rc) { if (b->inflight > 0) { b->inflight--; b->collected++; } return rc; }
    return iq2a_bank_collect_chunk(b, audio, clipped, baseband, out_stride, n_out, rms_dbfs);
}

static int resident_common(iq2a_bank* b, const void* dev_frames, int64_t first_frame, int64_t n_frames,
                           int64_t seg_begin, int64_t seg_end, int32_t warmup_rows, float* dev_audio,
                           float* dev_clipped, float* dev_baseband, int64_t out_stride, cudaStream_t st,
                           int64_t* n_out, int64_t* win0_out, int64_t* nwin_out, int64_t* mg_emit_out) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    if (!dev_frames || n_frames <= 0) { set_error("no frames"); return IQ2A_ERR_INVALID; }
    if (seg_begin < 0 || seg_end < seg_begin || warmup_rows < 0) { set_error("bad segment [%lld, %lld)", (long long)seg_begin, (long long)seg_end); return IQ2A_ERR_INVALID; }
    if (seg_end > first_frame + n_frames) { set_error("segment end %lld beyond the resident frames", (long long)seg_end); return IQ2A_ERR_INVALID; }
    const int D = b->D;
    const int64_t mg_emit = ceil_div(seg_begin, D);
    const int64_t mg_end = ceil_div(seg_end, D);
    const int64_t mg_begin = std::max<int64_t>(0, mg_emit - warmup_rows);
    // history the first computed row needs
    const int64_t need_from = std::max<int64_t>(0, (mg_begin - b->vd) * (int64_t)D);
    if (first_frame > need_from) {
        set_error("segment needs input history from sample %lld but the resident frames start at %lld",
                  (long long)need_from, (long long)first_frame);
        return IQ2A_ERR_INVALID;
    }
    if (n_out) *n_out = mg_end - mg_emit;
    if (mg_end - mg_emit > out_stride && (dev_audio || dev_clipped || dev_baseband)) { set_error("output stride too small"); return IQ2A_ERR_INVALID; }
    const int64_t chunk = b->cfg.ref_chunk;
    const int64_t k_lo = need_from / chunk;
    const int64_t k_hi = (std::max<int64_t>(seg_end, 1) - 1) / chunk;
    extend_phase_table(b, k_hi + 1);
    const int nseg = (int)(k_hi - k_lo + 1);
    std::vector<double> hp((size_t)b->C * nseg);
    for (int c = 0; c < b->C; ++c)
        for (int k = 0; k < nseg; ++k) hp[(size_t)c * nseg + k] = b->phase_tab[c][k_lo + k];

    CoreArgs a{};
    a.d_raw = dev_frames;
    a.raw_n0 = first_frame;
    a.raw_len = n_frames;
    a.raw_pad = 0;
    a.mg_begin = mg_begin;
    a.mg_emit = mg_emit;
    a.mg_end = mg_end;
    a.fresh = (warmup_rows > 0) || (seg_begin == 0);
    a.seg_origin = k_lo * chunk;
    a.seg_len = chunk;
    a.nseg = nseg;
    a.h_phase = hp.data();
    a.win0 = seg_begin / chunk - k_lo;
    a.nwin = k_hi - seg_begin / chunk + 1;
    a.d_audio = dev_audio;
    a.d_clip = dev_clipped;
    a.d_bb_out = dev_baseband;
    a.out_stride = out_stride;
    a.st = st;
    if (win0_out) *win0_out = seg_begin / chunk;
    if (nwin_out) *nwin_out = a.nwin;
    if (mg_emit_out) *mg_emit_out = mg_emit;
    // hp must outlive the async H2D copy: run_core copies from pageable memory, which the
    // runtime stages before returning, so the vector may go out of scope afterwards.
    return run_core(b, a);
}

int iq2a_bank_process_resident(iq2a_bank* b, const void* dev_frames, int64_t first_frame, int64_t n_frames,
                               int64_t seg_begin, int64_t seg_end, int32_t warmup_rows, float* dev_audio,
                               float* dev_clipped, float* dev_baseband, int64_t out_stride, int64_t* n_out,
                               double* rms_dbfs, int64_t rms_capacity) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    int64_t win0 = 0, nwin = 0, mg_emit = 0;
    int rc = resident_common(b, dev_frames, first_frame, n_frames, seg_begin, seg_end, warmup_rows, dev_audio,
                             dev_clipped, dev_baseband, out_stride, b->stream, n_out, &win0, &nwin, &mg_emit);
    if (rc) return rc;
    if (rms_dbfs && nwin > 0) {
        if (rms_capacity < nwin) { set_error("rms buffer holds %lld windows, need %lld", (long long)rms_capacity, (long long)nwin); return IQ2A_ERR_INVALID; }
        std::vector<double> ss((size_t)b->C * nwin);
        IQ2A_CUDA_TRY(cudaMemcpyAsync(ss.data(), b->d_sumsq, ss.size() * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        IQ2A_CUDA_TRY(cudaStreamSynchronize(b->stream));
        const int64_t chunk = b->cfg.ref_chunk;
        const int D = b->D;
        std::vector<int64_t> cnt(nwin);
        for (int64_t k = 0; k < nwin; ++k) {
            const int64_t lo = std::max((win0 + k) * chunk, seg_begin), hi = std::min((win0 + k + 1) * chunk, seg_end);
            cnt[k] = hi > lo ? ceil_div(hi, D) - ceil_div(lo, D) : 0;
        }
        for (int c = 0; c < b->C; ++c) stats_to_dbfs(ss.data() + (size_t)c * nwin, cnt.data(), nwin, rms_dbfs + (size_t)c * rms_capacity);
    }
    IQ2A_CUDA_TRY(cudaStreamSynchronize(b->stream));
    collect_timing(b);
    return IQ2A_OK;
}

int iq2a_bank_process_resident_async(iq2a_bank* b, const void* dev_frames, int64_t first_frame, int64_t n_frames,
                                     int64_t seg_begin, int64_t seg_end, int32_t warmup_rows, float* dev_audio,
                                     float* dev_clipped, float* dev_baseband, int64_t out_stride, void* cuda_stream) {
    if (!b) { set_error("null bank"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(b->cfg.device));
    return resident_common(b, dev_frames, first_frame, n_frames, seg_begin, seg_end, warmup_rows, dev_audio,
                           dev_clipped, dev_baseband, out_stride, (cudaStream_t)cuda_stream, nullptr, nullptr,
                           nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------
// stage-level entry points (host arrays in, host arrays out)
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return IQ2A_ERR_NOMEM; }
        return IQ2A_OK;
    }
};

static int stage_device(int32_t device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (the B200 path has no CPU fallback)"); return IQ2A_ERR_STATE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(device));
    return IQ2A_OK;
}

int iq2a_unpack_mix(const void* frames, int64_t n_frames, int32_t codec, int32_t iq_order, double phase,
                    double w_signed, float* out_c64, int32_t device) {
    if (n_frames < 0 || (n_frames && (!frames || !out_c64))) { set_error("bad arguments"); return IQ2A_ERR_INVALID; }
    if (codec < 0 || codec > 2) { set_error("unknown codec %d", codec); return IQ2A_ERR_INVALID; }
    if (iq_order < 0 || iq_order > 3) { set_error("Unsupported iq_order %d", iq_order); return IQ2A_ERR_INVALID; }
    if (n_frames == 0) return IQ2A_OK;
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf in, out;
    const size_t fb = frame_bytes(codec);
    if ((rc = in.alloc(n_frames * fb)) || (rc = out.alloc(n_frames * sizeof(float2)))) return rc;
    IQ2A_CUDA_TRY(cudaMemcpy(in.p, frames, n_frames * fb, cudaMemcpyHostToDevice));
    if ((rc = launch_unpack_mix(in.p, n_frames, codec, iq_order, phase, w_signed, (float2*)out.p, 0))) return rc;
    IQ2A_CUDA_TRY(cudaMemcpy(out_c64, out.p, n_frames * sizeof(float2), cudaMemcpyDeviceToHost));
    return IQ2A_OK;
}

int iq2a_fir(const float* in_c64, int64_t n, const float* history_c64, const double* taps, int32_t ntaps,
             float* out_c64, int32_t device) {
    if (n < 0 || ntaps < 1 || !taps || (n && (!in_c64 || !out_c64))) { set_error("bad arguments"); return IQ2A_ERR_INVALID; }
    if (n == 0) return IQ2A_OK;
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf x, y, h;
    const int64_t nh = ntaps - 1;
    if ((rc = x.alloc((n + nh) * sizeof(float2))) || (rc = y.alloc(n * sizeof(float2))) || (rc = h.alloc(ntaps * sizeof(double)))) return rc;
    if (nh) {
        if (history_c64) IQ2A_CUDA_TRY(cudaMemcpy(x.p, history_c64, nh * sizeof(float2), cudaMemcpyHostToDevice));
        else IQ2A_CUDA_TRY(cudaMemset(x.p, 0, nh * sizeof(float2)));
    }
    IQ2A_CUDA_TRY(cudaMemcpy((float2*)x.p + nh, in_c64, n * sizeof(float2), cudaMemcpyHostToDevice));
    IQ2A_CUDA_TRY(cudaMemcpy(h.p, taps, ntaps * sizeof(double), cudaMemcpyHostToDevice));
    if ((rc = launch_fir_direct((const float2*)x.p, n, (const double*)h.p, ntaps, (float2*)y.p, 0))) return rc;
    IQ2A_CUDA_TRY(cudaMemcpy(out_c64, y.p, n * sizeof(float2), cudaMemcpyDeviceToHost));
    return IQ2A_OK;
}

int iq2a_decimate(const float* in_c64, int64_t n, int32_t factor, int64_t offset, float* out_c64, int64_t* n_out,
                  int32_t device) {
    if (n < 0 || factor < 1 || offset < 0) { set_error("bad arguments"); return IQ2A_ERR_INVALID; }
    const int64_t start = ((-offset) % factor + factor) % factor;     // processing.py:357
    const int64_t cnt = n > start ? ceil_div(n - start, factor) : 0;
    if (n_out) *n_out = cnt;
    if (cnt == 0) return IQ2A_OK;
    if (!in_c64 || !out_c64) { set_error("null buffers"); return IQ2A_ERR_INVALID; }
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf x, y;
    if ((rc = x.alloc(n * sizeof(float2))) || (rc = y.alloc(cnt * sizeof(float2)))) return rc;
    IQ2A_CUDA_TRY(cudaMemcpy(x.p, in_c64, n * sizeof(float2), cudaMemcpyHostToDevice));
    if ((rc = launch_decimate((const float2*)x.p, start, factor, cnt, (float2*)y.p, 0))) return rc;
    IQ2A_CUDA_TRY(cudaMemcpy(out_c64, y.p, cnt * sizeof(float2), cudaMemcpyDeviceToHost));
    return IQ2A_OK;
}

int iq2a_demod(int32_t mode, int32_t agc_enabled, double deemph_alpha, const float* in_c64, int64_t n,
               iq2a_channel_state* state, float* audio, double* rms_dbfs, int32_t device) {
    if (mode < 0 || mode > IQ2A_MODE_LSB) { set_error("Unsupported demod mode %d", mode); return IQ2A_ERR_INVALID; }
    if (n < 0 || !state || (n && (!in_c64 || !audio))) { set_error("bad arguments"); return IQ2A_ERR_INVALID; }
    if (n == 0) return IQ2A_OK;
    int rc = stage_device(device);
    if (rc) return rc;
    const int64_t stride = (n + 63) & ~(int64_t)63;
    const int64_t ntiles = tail_tiles(n);
    DevBuf bb, pre, tmp, aud, agg, ss, st, chn;
    if ((rc = bb.alloc(stride * sizeof(float2))) || (rc = pre.alloc(stride * sizeof(float))) ||
        (rc = tmp.alloc(stride * sizeof(float))) || (rc = aud.alloc(stride * sizeof(float))) ||
        (rc = agg.alloc(ntiles * sizeof(double2))) || (rc = ss.alloc(sizeof(double))) ||
        (rc = st.alloc(sizeof(iq2a_channel_state))) || (rc = chn.alloc(sizeof(TailChan))))
        return rc;
    TailChan tc{mode, agc_enabled ? 1 : 0, deemph_alpha, 1.0 - deemph_alpha, 0, 0};
    IQ2A_CUDA_TRY(cudaMemcpy(bb.p, in_c64, n * sizeof(float2), cudaMemcpyHostToDevice));
    IQ2A_CUDA_TRY(cudaMemcpy(st.p, state, sizeof(*state), cudaMemcpyHostToDevice));
    IQ2A_CUDA_TRY(cudaMemcpy(chn.p, &tc, sizeof(tc), cudaMemcpyHostToDevice));
    IQ2A_CUDA_TRY(cudaMemset(ss.p, 0, sizeof(double)));
    TailParams t{};
    t.bb = (const float2*)bb.p;  t.bb_stride = stride;
    t.pre = (float*)pre.p;       t.tmp = (float*)tmp.p;   t.work_stride = stride;
    t.audio = (float*)aud.p;     t.clipped = nullptr;     t.out_stride = stride;
    t.agg = (double2*)agg.p;     t.ntiles = ntiles;
    t.sumsq = (double*)ss.p;     t.nwin = 1;              t.win0 = 0;
    t.state = (iq2a_channel_state*)st.p;
    t.chan = (const TailChan*)chn.p;
    t.nchan = 1;
    t.n = n;  t.n_skip = 0;  t.mg0 = 0;  t.decim = 1;
    t.seg_origin = 0;  t.seg_len = std::max<int64_t>(n, 1);   // one call == one chunk: AGC restarts at row 0
    t.fresh = 0;
    t.dc_radius = 0.995;  t.agc_target = std::pow(10.0, -12.0 / 20.0);  t.agc_decay = 0.001;
    const bool agc = (mode == IQ2A_MODE_USB || mode == IQ2A_MODE_LSB) && agc_enabled;
    if ((rc = launch_tail(t, agc, 0, nullptr))) return rc;
    double sumsq = 0.0;
    IQ2A_CUDA_TRY(cudaMemcpy(audio, aud.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    IQ2A_CUDA_TRY(cudaMemcpy(state, st.p, sizeof(*state), cudaMemcpyDeviceToHost));
    IQ2A_CUDA_TRY(cudaMemcpy(&sumsq, ss.p, sizeof(double), cudaMemcpyDeviceToHost));
    if (rms_dbfs) stats_to_dbfs(&sumsq, &n, 1, rms_dbfs);
    return IQ2A_OK;
}

int iq2a_scan(int32_t kind, double deemph_alpha, const float* in, int64_t n, iq2a_channel_state* state,
              float* out, int32_t device) {
    if (kind < 0 || kind > 2) { set_error("unknown recurrence kind %d", kind); return IQ2A_ERR_INVALID; }
    if (n < 0 || !state || (n && (!in || !out))) { set_error("bad arguments"); return IQ2A_ERR_INVALID; }
    if (n == 0) return IQ2A_OK;
    int rc = stage_device(device);
    if (rc) return rc;
    const int64_t stride = (n + 63) & ~(int64_t)63;
    const int64_t ntiles = tail_tiles(n);
    DevBuf pre, aud, agg, st, chn;
    if ((rc = pre.alloc(stride * sizeof(float))) || (rc = aud.alloc(stride * sizeof(float))) ||
        (rc = agg.alloc(ntiles * sizeof(double2))) || (rc = st.alloc(sizeof(iq2a_channel_state))) ||
        (rc = chn.alloc(sizeof(TailChan))))
        return rc;
    TailChan tc{MODE_RAW_DEEMPH + kind, 0, deemph_alpha, 1.0 - deemph_alpha, 0, 0};
    IQ2A_CUDA_TRY(cudaMemcpy(pre.p, in, n * sizeof(float), cudaMemcpyHostToDevice));
    IQ2A_CUDA_TRY(cudaMemcpy(st.p, state, sizeof(*state), cudaMemcpyHostToDevice));
    IQ2A_CUDA_TRY(cudaMemcpy(chn.p, &tc, sizeof(tc), cudaMemcpyHostToDevice));
    TailParams t{};
    t.pre = (float*)pre.p;  t.tmp = nullptr;  t.work_stride = stride;
    t.audio = (float*)aud.p;  t.out_stride = stride;
    t.agg = (double2*)agg.p;  t.ntiles = ntiles;
    t.state = (iq2a_channel_state*)st.p;
    t.chan = (const TailChan*)chn.p;
    t.nchan = 1;
    t.n = n;  t.mg0 = 0;  t.decim = 1;
    t.seg_origin = 0;  t.seg_len = std::max<int64_t>(n, 1);
    t.skip_pre = 1;
    t.dc_radius = 0.995;  t.agc_target = std::pow(10.0, -12.0 / 20.0);  t.agc_decay = 0.001;
    const float peak_in = state->peak;
    if ((rc = launch_tail(t, false, 0, nullptr))) return rc;
    IQ2A_CUDA_TRY(cudaMemcpy(out, aud.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    IQ2A_CUDA_TRY(cudaMemcpy(state, st.p, sizeof(*state), cudaMemcpyDeviceToHost));
    state->peak = peak_in;      // the peak belongs to the writer, not to a bare recurrence
    return IQ2A_OK;
}

}  // extern "C"
