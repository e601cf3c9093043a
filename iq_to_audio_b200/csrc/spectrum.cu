// Spectrum / waterfall preview on the GPU.
//
// The reference computes its previews on the CPU in float64 (src/iq_to_audio/spectrum.py):
//   compute_psd          :15-45    one Hann-windowed, zero-padded FFT of the first <= nfft samples
//   streaming_waterfall  :54-92    PSD (dB) of every nfft-window at stride hop; their mean; a capped slice list
//   _sliding_windows     :95-128   how windows are cut out of a chunked stream (and their start indices)
//   _SlidingFFT.psd      :159-171  window -> complex128 FFT -> fftshift -> |X|^2/scale -> 10 log10(. + 1e-18)
//   _WaterfallAggregator :174-208  pairwise halving of the slice list whenever it exceeds max_slices
// The interactive front end asks for nfft = 65536 ... 524288 (interactive/panels.py:238), so the large transform
// is the normal case.  Layout of the device version:
//   * float64 throughout, like the reference.  nfft = N1*N2 (four-step): k_cols transforms 16 (8) adjacent
//     columns n2 over n1 in shared memory, applies W_N^(n2 k1) and writes Y[k1][n2]; k_rows transforms 16 (8)
//     adjacent rows k1 over n2 and emits dB bins k1 + N1 k2, already fft-shifted.  nfft <= 8192 runs in one
//     CTA per window (k_small).  In-place radix-2^2 DIF passes, natural-order input, bit-reversed read-out.
//   * k_accum adds the dB rows of a batch to the running float64 sum in window order and appends float32 copies
//     to the slice matrix; k_halve is the aggregator's pair averaging ((a + b)/2 in float64, rounded to float32).
//   * the host object keeps the reference's pending-tail and start-index bookkeeping exactly, including the way
//     the start index slips back by the pending length on every chunk after a yield (spectrum.py:108-110,:125).
#include <cmath>
#include <vector>

#include "common.cuh"
#include "fft64.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

constexpr double kPsdEps = 1e-18;      // _NUMPY_EPS, spectrum.py:12
constexpr int kSmallMax = 8192;        // largest transform done by one CTA (128 KiB of shared memory)

// np.hanning(n): 0.5 + 0.5 cos(pi (1 - n + 2 i) / (n - 1)); n == 1 -> 1.0
__host__ __device__ inline double hann_at(int i, int n) {
    if (n == 1) return 1.0;
    return 0.5 + 0.5 * cos(3.14159265358979323846 * (double)(1 - n + 2 * i) / (double)(n - 1));
}

__global__ void k_spec_hann(double* __restrict__ w, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = hann_at(i, n);
}

struct SpecWin {
    const void* raw;     // device frames; window w starts at frame first + w*hop
    int64_t first;
    int64_t hop;
    int n_use;           // samples taken per window (<= nfft); the rest of the transform is zero padding
    int nfft, log2n;
    int iq_swap, q_neg;
    const double* win;   // [n_use]
    const double2* tw;   // [nfft]
    double scale;        // n_use * fs * win_power + eps
};

template <int FMT>
__device__ __forceinline__ double2 load_windowed(const SpecWin& p, int w, int i) {
    if (i >= p.n_use) return make_double2(0.0, 0.0);
    using R = typename RawT<FMT>::type;
    const float2 v = raw_to_c64<FMT>(reinterpret_cast<const R*>(p.raw)[p.first + (int64_t)w * p.hop + i], p.iq_swap, p.q_neg);
    const double g = p.win[i];
    return make_double2((double)v.x * g, (double)v.y * g);
}

__device__ __forceinline__ double to_db(double2 x, double scale) {
    return 10.0 * log10((x.x * x.x + x.y * x.y) / scale + kPsdEps);
}

// one CTA per window, nfft <= kSmallMax
template <int FMT>
__global__ void __launch_bounds__(512) k_spec_small(SpecWin p, double* __restrict__ db) {
    extern __shared__ double2 s_fft[];
    const int w = blockIdx.x;
    for (int i = threadIdx.x; i < p.nfft; i += blockDim.x) s_fft[i] = load_windowed<FMT>(p, w, i);
    __syncthreads();
    fft_dif_shared<1>(s_fft, p.nfft, p.tw, p.nfft);
    double* row = db + (size_t)w * p.nfft;
    const int half = p.nfft >> 1;
    for (int i = threadIdx.x; i < p.nfft; i += blockDim.x) {
        const int k = (i + half) & (p.nfft - 1);                       // fftshift
        row[i] = to_db(s_fft[bitrev_n(k, p.log2n)], p.scale);
    }
}

// four-step, first half: NC columns n2 of the N1 x N2 view x[N2 n1 + n2]; out Y[k1][n2] * W_N^(n2 k1)
template <int FMT, int NC>
__global__ void __launch_bounds__(512) k_spec_cols(SpecWin p, int n1, int n2, double2* __restrict__ y) {
    extern __shared__ double2 s_fft[];
    const int w = blockIdx.y, c0 = blockIdx.x * NC;
    const int log2n1 = 31 - __clz(n1);
    for (int t = threadIdx.x; t < n1 * NC; t += blockDim.x) {
        const int c = t % NC, r = t / NC;
        s_fft[t] = load_windowed<FMT>(p, w, r * n2 + c0 + c);
    }
    __syncthreads();
    fft_dif_shared<NC>(s_fft, n1, p.tw, p.nfft);
    double2* yw = y + (size_t)w * p.nfft;
    for (int t = threadIdx.x; t < n1 * NC; t += blockDim.x) {
        const int c = t % NC, k1 = t / NC;
        const double2 v = s_fft[(size_t)bitrev_n(k1, log2n1) * NC + c];
        yw[(size_t)k1 * n2 + c0 + c] = cmul(v, p.tw[(c0 + c) * k1]);
    }
}

// four-step, second half: NC rows k1, transform over n2, emit bins k = k1 + N1 k2 (fft-shifted) in dB
template <int NC>
__global__ void __launch_bounds__(512) k_spec_rows(SpecWin p, int n1, int n2, const double2* __restrict__ y,
                                                   double* __restrict__ db) {
    extern __shared__ double2 s_fft[];
    const int w = blockIdx.y, r0 = blockIdx.x * NC;
    const int log2n2 = 31 - __clz(n2);
    const double2* yw = y + (size_t)w * p.nfft;
    for (int t = threadIdx.x; t < n2 * NC; t += blockDim.x) {
        const int c = t / n2, j = t - c * n2;                            // coalesced along the row
        s_fft[(size_t)j * NC + c] = yw[(size_t)(r0 + c) * n2 + j];
    }
    __syncthreads();
    fft_dif_shared<NC>(s_fft, n2, p.tw, p.nfft);
    double* row = db + (size_t)w * p.nfft;
    const int half = p.nfft >> 1;
    for (int t = threadIdx.x; t < n2 * NC; t += blockDim.x) {
        const int c = t % NC, k2 = t / NC;
        const int k = r0 + c + n1 * k2;
        row[(k + half) & (p.nfft - 1)] = to_db(s_fft[(size_t)bitrev_n(k2, log2n2) * NC + c], p.scale);
    }
}

// psd_sum += rows (in window order) and float32 copies appended to the slice matrix (spectrum.py:76-81, :183)
__global__ void k_spec_accum(const double* __restrict__ db, int rows, int nfft, double* __restrict__ sum,
                             float* __restrict__ slices) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfft) return;
    double acc = sum[i];
    for (int r = 0; r < rows; ++r) {
        const double v = db[(size_t)r * nfft + i];
        acc += v;
        if (slices) slices[(size_t)r * nfft + i] = (float)v;
    }
    sum[i] = acc;
}

// _WaterfallAggregator._maybe_reduce (spectrum.py:193-208): adjacent pairs averaged in float64, odd tail kept
__global__ void k_spec_halve(const float* __restrict__ src, int len, int nfft, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= nfft) return;
    const float a = src[(size_t)(2 * r) * nfft + i];
    if (2 * r + 1 < len) {
        const float b = src[(size_t)(2 * r + 1) * nfft + i];
        dst[(size_t)r * nfft + i] = (float)(((double)a + (double)b) / 2.0);
    } else {
        dst[(size_t)r * nfft + i] = a;
    }
}

struct SpecPlan {
    int nfft = 0, log2n = 0, n1 = 0, n2 = 0, nc = 0;   // n1 == 0: single-CTA transform
    double2* tw = nullptr;
    int device = 0;
};

static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

static int plan_init(SpecPlan& pl, int nfft, int device, cudaStream_t st) {
    pl.nfft = nfft;
    pl.device = device;
    pl.log2n = 0;
    while ((1 << pl.log2n) < nfft) ++pl.log2n;
    if (nfft > kSmallMax) {
        pl.n1 = 1 << ((pl.log2n + 1) / 2);
        pl.n2 = nfft / pl.n1;
        pl.nc = pl.n1 > 512 ? 8 : 16;
    }
    IQ2A_CUDA_TRY(cudaMalloc(&pl.tw, (size_t)nfft * sizeof(double2)));
    k_fft64_twiddle<<<(nfft + 255) / 256, 256, 0, st>>>(pl.tw, nfft);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

template <int FMT>
static int launch_windows_fmt(const SpecPlan& pl, SpecWin p, int nwin, double2* scratch, double* db, cudaStream_t st) {
    if (pl.n1 == 0) {
        const size_t smem = (size_t)pl.nfft * sizeof(double2);
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_spec_small<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int nt = pl.nfft >= 2048 ? 512 : (pl.nfft >= 512 ? 128 : 32);
        k_spec_small<FMT><<<nwin, nt, smem, st>>>(p, db);
    } else if (pl.nc == 16) {
        const size_t sa = (size_t)pl.n1 * 16 * sizeof(double2), sb = (size_t)pl.n2 * 16 * sizeof(double2);
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_spec_cols<FMT, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa));
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_spec_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
        k_spec_cols<FMT, 16><<<dim3(pl.n2 / 16, nwin), 512, sa, st>>>(p, pl.n1, pl.n2, scratch);
        k_spec_rows<16><<<dim3(pl.n1 / 16, nwin), 512, sb, st>>>(p, pl.n1, pl.n2, scratch, db);
    } else {
        const size_t sa = (size_t)pl.n1 * 8 * sizeof(double2), sb = (size_t)pl.n2 * 8 * sizeof(double2);
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_spec_cols<FMT, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa));
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_spec_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
        k_spec_cols<FMT, 8><<<dim3(pl.n2 / 8, nwin), 512, sa, st>>>(p, pl.n1, pl.n2, scratch);
        k_spec_rows<8><<<dim3(pl.n1 / 8, nwin), 512, sb, st>>>(p, pl.n1, pl.n2, scratch, db);
    }
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

static int launch_windows(const SpecPlan& pl, int codec, const SpecWin& p, int nwin, double2* scratch, double* db,
                          cudaStream_t st) {
    switch (codec) {
        case CODEC_S16: return launch_windows_fmt<CODEC_S16>(pl, p, nwin, scratch, db, st);
        case CODEC_U8: return launch_windows_fmt<CODEC_U8>(pl, p, nwin, scratch, db, st);
        default: return launch_windows_fmt<CODEC_F32>(pl, p, nwin, scratch, db, st);
    }
}

static int frame_bytes(int codec) { return codec == CODEC_S16 ? 4 : (codec == CODEC_U8 ? 2 : 8); }

static int check_common(int codec, int order, int nfft, double fs, int device) {
    if (codec < 0 || codec > 2 || order < 0 || order > 3) { set_error("bad codec/order"); return IQ2A_ERR_INVALID; }
    if (!is_pow2(nfft) || nfft < 2 || nfft > (1 << 20)) {
        set_error("nfft must be a power of two in [2, 2^20] (got %d)", nfft);
        return IQ2A_ERR_INVALID;
    }
    if (!(fs > 0.0)) { set_error("sample rate must be positive"); return IQ2A_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (the B200 path has no CPU fallback)");
        return IQ2A_ERR_STATE;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return IQ2A_ERR_INVALID; }
    return IQ2A_OK;
}

// Hann power the way the reference forms it: np.sum(window**2) / n (pairwise summation differs from this plain
// loop only in the last bits; the result enters a 10 log10, far below the parity tolerance).
static double hann_power(int n) {
    long double acc = 0.0L;
    for (int i = 0; i < n; ++i) {
        const double w = hann_at(i, n);
        acc += (long double)(w * w);
    }
    return (double)(acc / n);
}

}  // namespace iq2a

using namespace iq2a;

struct iq2a_spectrum {
    SpecPlan plan;
    int hop = 0, max_slices = 0, codec = 0, swap = 0, qneg = 0, device = 0;
    double fs = 0.0, scale = 0.0;
    double* win = nullptr;
    double* sum = nullptr;
    float* slices[2] = {nullptr, nullptr};
    int cur = 0, n_slices = 0;
    std::vector<double> times;
    double* db = nullptr;
    double2* scratch = nullptr;
    int batch_cap = 0;
    char* raw[2] = {nullptr, nullptr};
    int64_t raw_cap = 0;          // frames
    int rcur = 0;
    int64_t pending = 0, offset = 0, frames = 0, launches = 0;
    cudaStream_t st = nullptr;
    ~iq2a_spectrum() {
        cudaSetDevice(device);
        if (plan.tw) cudaFree(plan.tw);
        if (win) cudaFree(win);
        if (sum) cudaFree(sum);
        for (auto* p : slices) if (p) cudaFree(p);
        if (db) cudaFree(db);
        if (scratch) cudaFree(scratch);
        for (auto* p : raw) if (p) cudaFree(p);
        if (st) cudaStreamDestroy(st);
    }
};

namespace iq2a {

static int spectrum_windows(iq2a_spectrum* s, int64_t first, int64_t nwin, int64_t block_offset) {
    const int nfft = s->plan.nfft;
    int64_t done = 0;
    while (done < nwin) {
        const int room = s->max_slices + 1 - s->n_slices;
        const int t = (int)std::min<int64_t>(std::min<int64_t>(nwin - done, room), s->batch_cap);
        SpecWin p{};
        p.raw = s->raw[s->rcur];
        p.first = first + done * s->hop;
        p.hop = s->hop;
        p.n_use = nfft;
        p.nfft = nfft;
        p.log2n = s->plan.log2n;
        p.iq_swap = s->swap;
        p.q_neg = s->qneg;
        p.win = s->win;
        p.tw = s->plan.tw;
        p.scale = s->scale;
        int rc = launch_windows(s->plan, s->codec, p, t, s->scratch, s->db, s->st);
        if (rc != IQ2A_OK) return rc;
        k_spec_accum<<<(nfft + 255) / 256, 256, 0, s->st>>>(s->db, t, nfft, s->sum,
                                                           s->slices[s->cur] + (size_t)s->n_slices * nfft);
        IQ2A_CUDA_TRY(cudaGetLastError());
        s->launches += s->plan.n1 ? 3 : 2;
        for (int r = 0; r < t; ++r)
            s->times.push_back((double)(block_offset + (done + r) * s->hop) / s->fs);
        s->n_slices += t;
        s->frames += t;
        done += t;
        while (s->n_slices > s->max_slices) {                     // one pass unless max_slices == 1
            const int len = s->n_slices, out = (len + 1) / 2;
            k_spec_halve<<<dim3((nfft + 255) / 256, out), 256, 0, s->st>>>(s->slices[s->cur], len, nfft,
                                                                           s->slices[s->cur ^ 1]);
            IQ2A_CUDA_TRY(cudaGetLastError());
            ++s->launches;
            s->cur ^= 1;
            std::vector<double> nt;
            nt.reserve(out);
            for (int r = 0; r < len; r += 2) nt.push_back(s->times[r]);
            s->times.swap(nt);
            s->n_slices = out;
        }
    }
    return IQ2A_OK;
}

}  // namespace iq2a

extern "C" {

int iq2a_psd(const void* raw, int64_t n_frames, int32_t codec, int32_t order, int32_t nfft, double sample_rate,
             double* psd_db, int32_t device) {
    if (!raw || !psd_db) { set_error("null buffer"); return IQ2A_ERR_INVALID; }
    if (n_frames <= 0) { set_error("Cannot compute PSD for an empty signal."); return IQ2A_ERR_INVALID; }
    int rc = check_common(codec, order, nfft, sample_rate, device);
    if (rc != IQ2A_OK) return rc;
    IQ2A_CUDA_TRY(cudaSetDevice(device));
    const int n_use = (int)std::min<int64_t>(n_frames, nfft);
    const int bpf = frame_bytes(codec);
    SpecPlan pl;
    cudaStream_t st = nullptr;
    void* d_raw = nullptr;
    double *d_win = nullptr, *d_db = nullptr;
    double2* d_scr = nullptr;
    auto cleanup = [&]() {
        if (pl.tw) cudaFree(pl.tw);
        if (d_raw) cudaFree(d_raw);
        if (d_win) cudaFree(d_win);
        if (d_db) cudaFree(d_db);
        if (d_scr) cudaFree(d_scr);
    };
    rc = plan_init(pl, nfft, device, st);
    if (rc != IQ2A_OK) { cleanup(); return rc; }
    cudaError_t e = cudaMalloc(&d_raw, (size_t)n_use * bpf);
    if (e == cudaSuccess) e = cudaMalloc(&d_win, (size_t)n_use * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_db, (size_t)nfft * sizeof(double));
    if (e == cudaSuccess && pl.n1) e = cudaMalloc(&d_scr, (size_t)nfft * sizeof(double2));
    if (e == cudaSuccess) e = cudaMemcpy(d_raw, raw, (size_t)n_use * bpf, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("psd setup failed: %s", cudaGetErrorString(e)); cleanup(); return IQ2A_ERR_CUDA; }
    k_spec_hann<<<(n_use + 255) / 256, 256, 0, st>>>(d_win, n_use);
    SpecWin p{};
    p.raw = d_raw;
    p.first = 0;
    p.hop = 0;
    p.n_use = n_use;
    p.nfft = nfft;
    p.log2n = pl.log2n;
    p.iq_swap = order & 1;
    p.q_neg = (order >> 1) & 1;
    p.win = d_win;
    p.tw = pl.tw;
    p.scale = (double)n_use * sample_rate * hann_power(n_use) + kPsdEps;      // spectrum.py:41
    rc = launch_windows(pl, codec, p, 1, d_scr, d_db, st);
    if (rc == IQ2A_OK) {
        e = cudaMemcpy(psd_db, d_db, (size_t)nfft * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("psd copy failed: %s", cudaGetErrorString(e)); rc = IQ2A_ERR_CUDA; }
    }
    cleanup();
    return rc;
}

int iq2a_spectrum_create(int32_t nfft, int32_t hop, int32_t max_slices, double sample_rate, int32_t codec,
                         int32_t order, int32_t device, iq2a_spectrum** out) {
    if (!out) { set_error("null out pointer"); return IQ2A_ERR_INVALID; }
    *out = nullptr;
    int rc = check_common(codec, order, nfft, sample_rate, device);
    if (rc != IQ2A_OK) return rc;
    IQ2A_CUDA_TRY(cudaSetDevice(device));
    iq2a_spectrum* s = new iq2a_spectrum();
    s->device = device;
    s->hop = hop > 0 ? hop : std::max(1, nfft / 4);                // spectrum.py:69
    s->max_slices = std::max(1, max_slices);                        // spectrum.py:178
    s->codec = codec;
    s->swap = order & 1;
    s->qneg = (order >> 1) & 1;
    s->fs = sample_rate;
    s->scale = (double)nfft * sample_rate * hann_power(nfft) + kPsdEps;   // spectrum.py:168
    auto fail = [&](int code) { delete s; return code; };
    if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream"); return fail(IQ2A_ERR_CUDA); }
    rc = plan_init(s->plan, nfft, device, s->st);
    if (rc != IQ2A_OK) return fail(rc);
    // batch: bounded by the slice room and by ~256 MiB of transform scratch
    const int64_t per_win = (int64_t)nfft * (sizeof(double) + (s->plan.n1 ? sizeof(double2) : 0));
    s->batch_cap = (int)std::max<int64_t>(1, std::min<int64_t>(s->max_slices + 1, (256ll << 20) / per_win));
    cudaError_t e = cudaMalloc(&s->win, (size_t)nfft * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->sum, (size_t)nfft * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->db, (size_t)s->batch_cap * nfft * sizeof(double));
    if (e == cudaSuccess && s->plan.n1) e = cudaMalloc(&s->scratch, (size_t)s->batch_cap * nfft * sizeof(double2));
    for (int i = 0; i < 2 && e == cudaSuccess; ++i)
        e = cudaMalloc(&s->slices[i], (size_t)(s->max_slices + 1) * nfft * sizeof(float));
    if (e != cudaSuccess) { set_error("spectrum allocation failed: %s", cudaGetErrorString(e)); return fail(IQ2A_ERR_NOMEM); }
    k_spec_hann<<<(nfft + 255) / 256, 256, 0, s->st>>>(s->win, nfft);
    if (cudaMemsetAsync(s->sum, 0, (size_t)nfft * sizeof(double), s->st) != cudaSuccess ||
        cudaStreamSynchronize(s->st) != cudaSuccess) { set_error("spectrum init failed"); return fail(IQ2A_ERR_CUDA); }
    *out = s;
    return IQ2A_OK;
}

void iq2a_spectrum_destroy(iq2a_spectrum* s) { delete s; }

int iq2a_spectrum_push(iq2a_spectrum* s, const void* raw, int64_t n_frames) {
    if (!s) { set_error("null spectrum"); return IQ2A_ERR_INVALID; }
    if (n_frames < 0 || (n_frames > 0 && !raw)) { set_error("bad chunk"); return IQ2A_ERR_INVALID; }
    if (n_frames == 0) return IQ2A_OK;                               // spectrum.py:106-107
    IQ2A_CUDA_TRY(cudaSetDevice(s->device));
    const int bpf = frame_bytes(s->codec);
    const int nfft = s->plan.nfft;
    const int64_t total = s->pending + n_frames;
    if (total > s->raw_cap) {
        const int64_t cap = total + total / 4 + nfft;
        char* nb[2] = {nullptr, nullptr};
        for (int i = 0; i < 2; ++i)
            if (cudaMalloc(&nb[i], (size_t)cap * bpf) != cudaSuccess) {
                if (nb[0]) cudaFree(nb[0]);
                set_error("spectrum input buffer allocation failed");
                return IQ2A_ERR_NOMEM;
            }
        if (s->pending)
            IQ2A_CUDA_TRY(cudaMemcpyAsync(nb[0], s->raw[s->rcur], (size_t)s->pending * bpf, cudaMemcpyDeviceToDevice, s->st));
        IQ2A_CUDA_TRY(cudaStreamSynchronize(s->st));
        for (auto* p : s->raw) if (p) cudaFree(p);
        s->raw[0] = nb[0];
        s->raw[1] = nb[1];
        s->rcur = 0;
        s->raw_cap = cap;
    }
    IQ2A_CUDA_TRY(cudaMemcpyAsync(s->raw[s->rcur] + (size_t)s->pending * bpf, raw, (size_t)n_frames * bpf,
                                  cudaMemcpyHostToDevice, s->st));
    // _sliding_windows (spectrum.py:95-128), with its offset bookkeeping kept as is
    if (s->pending) s->offset -= s->pending;
    if (total < nfft) {
        s->pending = total;
        s->offset += total;
        IQ2A_CUDA_TRY(cudaStreamSynchronize(s->st));
        return IQ2A_OK;
    }
    const int64_t nwin = (total - nfft) / s->hop + 1;
    int rc = spectrum_windows(s, 0, nwin, s->offset);
    if (rc != IQ2A_OK) return rc;
    const int64_t start = nwin * s->hop;
    int64_t keep = total > start ? total - start : 0;               // pending = block[start:]
    s->offset += total - keep;                                       // spectrum.py:125
    if (keep > nfft) keep = nfft;                                    // spectrum.py:126-127
    const int64_t from = total - keep;
    if (keep)
        IQ2A_CUDA_TRY(cudaMemcpyAsync(s->raw[s->rcur ^ 1], s->raw[s->rcur] + (size_t)from * bpf, (size_t)keep * bpf,
                                      cudaMemcpyDeviceToDevice, s->st));
    s->rcur ^= 1;
    s->pending = keep;
    IQ2A_CUDA_TRY(cudaStreamSynchronize(s->st));                      // the caller may reuse `raw` now
    return IQ2A_OK;
}

int iq2a_spectrum_counts(const iq2a_spectrum* s, int64_t* frames, int32_t* n_slices, int64_t* launches) {
    if (!s) { set_error("null spectrum"); return IQ2A_ERR_INVALID; }
    if (frames) *frames = s->frames;
    if (n_slices) *n_slices = s->n_slices;
    if (launches) *launches = s->launches;
    return IQ2A_OK;
}

int iq2a_spectrum_result(iq2a_spectrum* s, double* avg_psd_db, float* times, float* matrix) {
    if (!s) { set_error("null spectrum"); return IQ2A_ERR_INVALID; }
    if (s->frames == 0) {
        set_error("Input did not contain enough samples for one FFT frame.");      // spectrum.py:86-87
        return IQ2A_ERR_INVALID;
    }
    IQ2A_CUDA_TRY(cudaSetDevice(s->device));
    const int nfft = s->plan.nfft;
    if (avg_psd_db) {
        IQ2A_CUDA_TRY(cudaMemcpyAsync(avg_psd_db, s->sum, (size_t)nfft * sizeof(double), cudaMemcpyDeviceToHost, s->st));
        IQ2A_CUDA_TRY(cudaStreamSynchronize(s->st));
        for (int i = 0; i < nfft; ++i) avg_psd_db[i] /= (double)s->frames;         // spectrum.py:89
    }
    if (times)
        for (int r = 0; r < s->n_slices; ++r) times[r] = (float)s->times[r];
    if (matrix) {
        IQ2A_CUDA_TRY(cudaMemcpyAsync(matrix, s->slices[s->cur], (size_t)s->n_slices * nfft * sizeof(float),
                                      cudaMemcpyDeviceToHost, s->st));
        IQ2A_CUDA_TRY(cudaStreamSynchronize(s->st));
    }
    return IQ2A_OK;
}

}  // extern "C"
