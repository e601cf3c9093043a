// 48 kHz output stage (K14): libswresample's default polyphase resampler + flt -> s16, on the GPU.
//
// The reference pipes float32 audio to `ffmpeg ... -ar 48000 -acodec pcm_s16le`
// (src/iq_to_audio/processing.py:399-418); the arithmetic is libswresample's (resample.c, default
// options: filter_size 32, phase_shift 10, linear_interp, exact_rational, Kaiser beta 9, cutoff 0.97).
// oracle/swr_model.py restates it and is pinned against the real library; this file is the device
// version of that model:
//   * filter bank built on the host in float64 (windowed sinc, every phase normalised by the tap sum
//     of phase 0), stored as float32 [phase_count][filter_length];
//   * output k sits at t_k = k*dst_incr/(src_incr*phase_count) input samples; its window starts at
//     floor(t_k) - center; value = v1 + (v2 - v1)*frac/src_incr, v2 from the next phase (phase 0 of
//     the next sample at the wrap);
//   * stream edges by reflection (x[-n] = x[n]; x[N+j] = x[N-1-j] at flush), total count by the
//     library's leftover rule;  flt -> s16 = clip(lrintf(x * 32768)).
// One thread per (output, channel); float64 accumulation.  The data rate here is tiny (48 kS/s per
// channel), the point of this stage is removing the second ffmpeg subprocess, not throughput.
#include <cmath>
#include <numeric>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

struct ResampleGeom {
    int fl, pc, center;
    int64_t src_incr, dst_incr;
};

__global__ void k_resample48(const float* __restrict__ buf, int64_t buf_stride, int64_t buf_base,
                             int64_t n_total, int reflect_end, const float* __restrict__ bank, ResampleGeom g,
                             int64_t k0, int64_t count, int16_t* __restrict__ out, int64_t out_stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= count) return;
    const int64_t k = k0 + i;
    // 128-bit product: k * dst_incr stays < 2^63 for k < 4e12 (dst_incr < 2^21)
    const int64_t tot = k * g.dst_incr;
    const int64_t idx = tot / g.src_incr, frac = tot - idx * g.src_incr;
    const int64_t samp = idx / g.pc;
    const int ph = (int)(idx - samp * g.pc);
    const int64_t start = samp - g.center;
    const bool wrap = ph == g.pc - 1;
    const float* h1 = bank + (size_t)ph * g.fl;
    const float* h2 = bank + (size_t)(wrap ? 0 : ph + 1) * g.fl;
    const float* x = buf + (size_t)c * buf_stride;
    auto sample = [&](int64_t j) -> double {
        if (j < 0) j = -j;                                     // invert_initial_buffer: x[-n] = x[n]
        if (reflect_end && j >= n_total) j = 2 * n_total - 1 - j;   // resample_flush: x[N+j] = x[N-1-j]
        if (j < buf_base || j >= n_total) return 0.0;
        return (double)x[j - buf_base];
    };
    double v1 = 0.0, v2 = 0.0;
    for (int t = 0; t < g.fl; ++t) {
        v1 = fma(sample(start + t), (double)h1[t], v1);
        v2 = fma(sample(start + t + (wrap ? 1 : 0)), (double)h2[t], v2);
    }
    const double val = v1 + (v2 - v1) * (double)frac / (double)g.src_incr;
    const float f = (float)val;
    float s = rintf(f * 32768.0f);
    s = fminf(fmaxf(s, -32768.0f), 32767.0f);
    out[(size_t)c * out_stride + i] = (int16_t)s;
}

static double bessel_i0(double x) {
    const double q = x * x / 4.0;
    double term = 1.0, total = 1.0;
    for (int k = 1; k < 64; ++k) {
        term = term * q / ((double)k * k);
        total += term;
    }
    return total;
}

}  // namespace iq2a

using namespace iq2a;

struct iq2a_resampler {
    int in_rate = 0, out_rate = 0, C = 0, device = 0;
    bool passthrough = false;
    ResampleGeom g{};
    float* d_bank = nullptr;
    float* d_buf = nullptr;      // [C][cap] pending input, element 0 <-> global index base
    int64_t cap = 0, base = 0, n_total = 0, k_next = 0;
    int16_t* d_out = nullptr;
    int64_t out_cap = 0;
    cudaStream_t stream = nullptr;
    ~iq2a_resampler() {
        cudaSetDevice(device);
        if (d_bank) cudaFree(d_bank);
        if (d_buf) cudaFree(d_buf);
        if (d_out) cudaFree(d_out);
        if (stream) cudaStreamDestroy(stream);
    }
    int64_t start_of(int64_t k) const {
        const __int128 tot = (__int128)k * g.dst_incr;
        const int64_t idx = (int64_t)(tot / g.src_incr);
        return idx / g.pc - g.center;
    }
    // number of outputs k (from 0) whose window [start_k, start_k + fl) ends at or before `limit`
    int64_t count_upto(int64_t limit) const {
        int64_t lo = 0, hi = (int64_t)((double)(limit + g.fl + 4) * out_rate / in_rate) + 8;
        while (lo < hi) {                                      // first k with start_k + fl > limit
            const int64_t mid = lo + (hi - lo) / 2;
            if (start_of(mid) + g.fl <= limit) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
};

namespace iq2a {

static int run_outputs(iq2a_resampler* r, int64_t k_end, bool reflect_end, int16_t* pcm, int64_t out_stride,
                       int64_t* n_out) {
    const int64_t count = std::max<int64_t>(0, k_end - r->k_next);
    if (n_out) *n_out = count;
    if (count == 0) return IQ2A_OK;
    if (!pcm || out_stride < count) { set_error("pcm buffer too small: %lld outputs", (long long)count); return IQ2A_ERR_INVALID; }
    if ((int64_t)r->C * count > r->out_cap) {
        if (r->d_out) cudaFree(r->d_out);
        r->out_cap = (int64_t)r->C * count * 5 / 4 + 1024;
        IQ2A_CUDA_TRY(cudaMalloc(&r->d_out, (size_t)r->out_cap * sizeof(int16_t)));
    }
    const dim3 grid((unsigned)((count + 127) / 128), r->C);
    k_resample48<<<grid, 128, 0, r->stream>>>(r->d_buf, r->cap, r->base, r->n_total, reflect_end ? 1 : 0, r->d_bank,
                                               r->g, r->k_next, count, r->d_out, count);
    IQ2A_CUDA_TRY(cudaGetLastError());
    IQ2A_CUDA_TRY(cudaMemcpy2DAsync(pcm, out_stride * sizeof(int16_t), r->d_out, count * sizeof(int16_t),
                                    count * sizeof(int16_t), r->C, cudaMemcpyDeviceToHost, r->stream));
    IQ2A_CUDA_TRY(cudaStreamSynchronize(r->stream));
    r->k_next = k_end;
    return IQ2A_OK;
}

}  // namespace iq2a

extern "C" {

int iq2a_resampler_create(int32_t in_rate, int32_t out_rate, int32_t n_channels, int32_t device, iq2a_resampler** out) {
    if (!out || in_rate < 1 || out_rate < 1 || n_channels < 1) { set_error("bad resampler arguments"); return IQ2A_ERR_INVALID; }
    *out = nullptr;
    if (out_rate >= in_rate) { set_error("the output stage only downsamples (channel rate %d, output %d); equal rates need no resampler", in_rate, out_rate); return IQ2A_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (the B200 path has no CPU fallback)"); return IQ2A_ERR_STATE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(device));
    iq2a_resampler* r = new iq2a_resampler();
    r->in_rate = in_rate; r->out_rate = out_rate; r->C = n_channels; r->device = device;
    r->passthrough = in_rate == out_rate;
    // resample_init (libswresample/resample.c), default options
    const double cutoff = 0.97, beta = 9.0;
    const int filter_size = 32, phase_shift = 10;
    const double factor = std::min((double)out_rate * cutoff / in_rate, 1.0);
    int pc = 1 << phase_shift;
    int fl = std::max((int)std::ceil(filter_size / factor), 1);
    if (fl > 1) fl = (fl + 1) & ~1;
    const int64_t gc = std::gcd((int64_t)out_rate, (int64_t)in_rate);
    if (out_rate / gc <= pc) pc = (int)(out_rate / gc);           // exact_rational
    r->g.fl = fl; r->g.pc = pc; r->g.center = (fl - 1) / 2;
    int64_t num = out_rate, den = (int64_t)in_rate * pc;
    const int64_t g2 = std::gcd(num, den);
    num /= g2; den /= g2;
    while (den < (1 << 20) && num < (1 << 20)) { den *= 2; num *= 2; }
    r->g.src_incr = num; r->g.dst_incr = den;
    // the filter bank depends on the two rates only: 1024 x 66 Kaiser-windowed sinc values take ~20 ms to build, and
    // a run opens one resampler per target, so the last one built is kept
    static std::mutex cache_mu;
    static std::vector<float> cache_bank;
    static int cache_in = 0, cache_out = 0;
    std::vector<float> bankf;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        if (cache_in == in_rate && cache_out == out_rate && cache_bank.size() == (size_t)pc * fl) bankf = cache_bank;
    }
    if (bankf.empty()) {
        std::vector<double> bank((size_t)pc * fl);
        double norm0 = 0.0;
        for (int ph = 0; ph < pc; ++ph)
            for (int i = 0; i < fl; ++i) {
                const double t = (double)(i - r->g.center) - (double)ph / pc;
                const double x = M_PI * t * factor;
                double y = x == 0.0 ? 1.0 : std::sin(x) / x;
                const double w = 2.0 * std::fabs(t) / fl;
                y *= bessel_i0(beta * std::sqrt(std::max(1.0 - w * w, 0.0)));
                bank[(size_t)ph * fl + i] = y;
                if (ph == 0) norm0 += y;
            }
        bankf.resize(bank.size());
        for (size_t i = 0; i < bank.size(); ++i) bankf[i] = (float)(bank[i] / norm0);
        std::lock_guard<std::mutex> lk(cache_mu);
        cache_bank = bankf;
        cache_in = in_rate;
        cache_out = out_rate;
    }
    if (cudaMalloc(&r->d_bank, bankf.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(r->d_bank, bankf.data(), bankf.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("resampler setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete r;
        return IQ2A_ERR_CUDA;
    }
    *out = r;
    return IQ2A_OK;
}

void iq2a_resampler_destroy(iq2a_resampler* r) { delete r; }

int iq2a_resampler_process(iq2a_resampler* r, const float* audio, int64_t n, int64_t in_stride, int16_t* pcm,
                           int64_t out_stride, int64_t* n_out) {
    if (!r) { set_error("null resampler"); return IQ2A_ERR_INVALID; }
    if (n < 0 || (n > 0 && !audio) || in_stride < n) { set_error("bad audio buffer"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(r->device));
    if (n_out) *n_out = 0;
    if (n == 0) return IQ2A_OK;
    // append: make room for [base, n_total + n)
    const int64_t have = r->n_total - r->base;
    if (have + n > r->cap) {
        const int64_t ncap = (have + n) * 3 / 2 + 4096;
        float* nb = nullptr;
        IQ2A_CUDA_TRY(cudaMalloc(&nb, (size_t)r->C * ncap * sizeof(float)));
        if (have > 0)
            IQ2A_CUDA_TRY(cudaMemcpy2DAsync(nb, ncap * sizeof(float), r->d_buf, r->cap * sizeof(float), have * sizeof(float),
                                            r->C, cudaMemcpyDeviceToDevice, r->stream));
        IQ2A_CUDA_TRY(cudaStreamSynchronize(r->stream));
        if (r->d_buf) cudaFree(r->d_buf);
        r->d_buf = nb;
        r->cap = ncap;
    }
    IQ2A_CUDA_TRY(cudaMemcpy2DAsync(r->d_buf + have, r->cap * sizeof(float), audio, in_stride * sizeof(float),
                                    n * sizeof(float), r->C, cudaMemcpyHostToDevice, r->stream));
    r->n_total += n;
    const int64_t k_end = r->count_upto(r->n_total);
    int rc = run_outputs(r, k_end, false, pcm, out_stride, n_out);
    if (rc) return rc;
    // drop input no later window needs: everything before the next window start (keep x[0..fl+2) while the
    // reflected prefix can still be referenced, i.e. while the next window starts before sample 0)
    const int64_t next_start = r->start_of(r->k_next);
    const int64_t new_base = std::max<int64_t>(r->base, std::min<int64_t>(next_start, r->n_total));
    if (new_base > r->base && next_start > r->g.fl + 2) {
        const int64_t keep = r->n_total - new_base;
        if (keep > 0) {
            // forward overlapping move within each row: stage through a temporary only when ranges overlap
            float* tmp = nullptr;
            IQ2A_CUDA_TRY(cudaMalloc(&tmp, (size_t)r->C * keep * sizeof(float)));
            IQ2A_CUDA_TRY(cudaMemcpy2DAsync(tmp, keep * sizeof(float), r->d_buf + (new_base - r->base), r->cap * sizeof(float),
                                            keep * sizeof(float), r->C, cudaMemcpyDeviceToDevice, r->stream));
            IQ2A_CUDA_TRY(cudaMemcpy2DAsync(r->d_buf, r->cap * sizeof(float), tmp, keep * sizeof(float), keep * sizeof(float),
                                            r->C, cudaMemcpyDeviceToDevice, r->stream));
            IQ2A_CUDA_TRY(cudaStreamSynchronize(r->stream));
            cudaFree(tmp);
        }
        r->base = new_base;
    }
    return IQ2A_OK;
}

int iq2a_resampler_flush(iq2a_resampler* r, int16_t* pcm, int64_t out_stride, int64_t* n_out) {
    if (!r) { set_error("null resampler"); return IQ2A_ERR_INVALID; }
    IQ2A_CUDA_TRY(cudaSetDevice(r->device));
    if (n_out) *n_out = 0;
    if (r->n_total == 0) return IQ2A_OK;
    // resample_flush: reflect (min(leftover, fl) + 1) / 2 samples, leftover = what the regular calls left
    const int64_t leftover = r->n_total - r->start_of(r->k_next);
    const int64_t refl = (std::min<int64_t>(leftover, r->g.fl) + 1) / 2;
    const int64_t k_end = r->count_upto(r->n_total + refl);
    return run_outputs(r, k_end, true, pcm, out_stride, n_out);
}

int iq2a_resampler_max_outputs(const iq2a_resampler* r, int64_t n_in, int64_t* n_out) {
    if (!r || !n_out) { set_error("null argument"); return IQ2A_ERR_INVALID; }
    *n_out = (int64_t)((double)(n_in + r->g.fl + 64) * r->out_rate / r->in_rate) + 64;
    return IQ2A_OK;
}

}  // extern "C"
