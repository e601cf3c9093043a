// Polyphase overlap-save channel bank: raw PCM frames -> complex64 channel-rate samples
// for up to kMaxGroup targets per launch.  One kernel, no intermediate global traffic.
//
// Replaces, for all targets at once, the reference's per-target
//   ComplexOscillator.mix -> OverlapSaveFIR.process -> Decimator.process
// (src/iq_to_audio/processing.py:289-297, :325-346, :354-360) together with the
// sample unpack / IQ order fix that ffmpeg + IQReader._extract_iq do (:143-158, :261-279).
//
// Math (iq_to_audio_b200/plan.py has the derivation and a numpy model):
//   s_c[m] = e^{j phi_c(mD)} * IFFT_M( sum_p G[c,p,:] * FFT_M(x_p) )[m],   x_p[m] = x[mD + p]
// A CTA owns a "block set" of BT consecutive overlap-save blocks.  It walks the D
// polyphase branches in tiles of P = LW/BT branches; a tile is LW independent M-point
// transforms (BT blocks x P branches), one per "lane slot":
//   pass 1  global PCM -> registers (coalesced: consecutive slots = consecutive branches),
//           R1-point DIF in registers, twiddle, store to the shared tile [row][slot]
//   pass 2  R2-point DIF in registers, store back in place -> X_p[j'] in slot order
//   MAC     thread <-> spectrum slot j': acc[c][b] += G[p][c][j'] * X[b][p][j']  (G from L2,
//           each G element reused for the BT blocks of the set)
// After the last tile the accumulators go through the inverse two passes (one M-point
// inverse transform per block and channel), the overlap rows are dropped, the NCO
// rotation is applied at the channel rate and complex64 samples are written.
//
// Layout choices: the tile row stride is LW+1 float2 so that both the pass stores
// (consecutive slots) and the MAC loads (consecutive rows) are bank-conflict free;
// PCM for the next tile is prefetched into registers before the MAC phase.
#pragma once
#include "common.cuh"
#include "fft_regs.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

constexpr int kThreads = 512;

template <int M>
struct Geo;
template <>
struct Geo<512> {
    static constexpr int R1 = 32, R2 = 16, LW = 32;
};
template <>
struct Geo<1024> {
    static constexpr int R1 = 32, R2 = 32, LW = 16;
};

// Inverse M-point transforms of the output spectra held in `tile` ([slot j'][YS] float2, j' = k1*R2 + k2
// <-> bin k1 + R1*k2), overlap rows dropped, NCO rotation at the channel rate, complex64 store.
// Shared by both channel-bank kernels.  Ends with the tile free for reuse (caller syncs).
template <int M, int CG, int BT = kBlocksPerSet, int NT = kThreads>
__device__ __forceinline__ void inverse_and_store(float2* tile, const float2* tw, const ChannelizeParams& p, int blk0) {
    using G = Geo<M>;
    constexpr int R1 = G::R1, R2 = G::R2;
    constexpr int NS = CG * BT;
    constexpr int YS = NS | 1;
    const int tid = threadIdx.x;
    const int D = p.decim;
    __syncthreads();
    // inverse pass A: for each k1, R2-point inverse over k2, then W_M^{-m2*k1}
    for (int idx = tid; idx < R1 * NS; idx += NT) {
        const int sy = idx % NS, k1 = idx / NS;
        float2 v[R2];
        float2* base = tile + (k1 * R2) * YS + sy;
#pragma unroll
        for (int i = 0; i < R2; ++i) v[i] = base[i * YS];
        dif<R2, -1>(v);
        static_for<R2>([&](auto mc) {
            constexpr int m2 = decltype(mc)::value;
            float2 x = v[bitrev<R2>(m2)];
            if constexpr (m2 != 0) x = cmul_conj(x, tw[(m2 * k1) & (M - 1)]);
            base[m2 * YS] = x;
        });
    }
    __syncthreads();
    // inverse pass B: for each m2, R1-point inverse over k1 -> y[R2*m1 + m2]
    for (int idx = tid; idx < R2 * NS; idx += NT) {
        const int sy = idx % NS, m2 = idx / NS;
        float2 v[R1];
        float2* base = tile + m2 * YS + sy;
#pragma unroll
        for (int i = 0; i < R1; ++i) v[i] = base[(i * R2) * YS];
        dif<R1, -1>(v);
        static_for<R1>([&](auto mc) {
            constexpr int m1 = decltype(mc)::value;
            base[(m1 * R2) * YS] = v[bitrev<R1>(m1)];
        });
    }
    __syncthreads();
    // ---------------- drop the overlap rows, rotate by the NCO, store ------------------------
    const int ld = p.ld;
    // NCO rotation: phi(mD) is linear in m inside a block, so e^{j phi} = P[b][c] * Q[c][r] with
    //   P = phasor at the block's first row, evaluated with the reference's exact chunk-wise phase bookkeeping,
    //   Q[c][r] = e^{j w_c D r}, a per-channel table built once per bank (float64 -> float32).
    // (A block that straddles a reference chunk boundary inherits the boundary's ~1e-9 rad wrap rounding.)
    __shared__ float2 s_base[kMaxGroup * kBlocksPerSet];
    if (tid < NS) {
        const int b = tid / CG, c = tid % CG;
        const int64_t mg_b = p.mg_begin + (int64_t)(blk0 + b) * ld;
        s_base[tid] = phasor_f32(nco_phase(p.phase, c, p.w[c], mg_b * (int64_t)D) + p.phase_bias[c]);
    }
    __syncthreads();
    // one (block, channel) spectrum at a time, rows strided over the threads: no division by the run-time row count,
    // 32-bit indices inside the row loop, the block / range tests hoisted out of it
#pragma unroll 1
    for (int sy = 0; sy < NS; ++sy) {
        const int b = sy / CG, c = sy % CG;
        const int blk = blk0 + b;
        if (blk >= p.nblocks || c >= p.nchan) continue;
        const int64_t row0 = (int64_t)blk * ld;                               // first row of the block, from mg_begin
        const int64_t left = p.mg_end - p.mg_begin - row0;
        const int rows = left < (int64_t)ld ? (int)left : ld;
        const float2 base = s_base[sy];
        const float2* __restrict__ rot = p.rot + (size_t)c * ld;
        const float2* __restrict__ col = tile + p.vd * YS + sy;
        float2* __restrict__ out = p.out + (size_t)c * p.out_stride + row0;
        for (int r = tid; r < rows; r += NT) {
            const float2 lo = cmul(base, __ldg(rot + r));
            out[r] = cmul(col[r * YS], lo);
        }
    }
}

template <int M, int CG, int FMT>
__global__ void __launch_bounds__(kThreads, 1) k_channelize(const ChannelizeParams p) {
    using G = Geo<M>;
    constexpr int R1 = G::R1, R2 = G::R2, LW = G::LW;
    constexpr int BT = kBlocksPerSet;
    constexpr int P = LW / BT;          // branches per tile
    constexpr int NRG = kThreads / LW;  // row groups
    constexpr int RS = LW + 1;          // tile row stride (float2)
    constexpr int JPT = M / kThreads;   // spectrum slots per thread in the MAC phase
    constexpr int NS = CG * BT;         // inverse transforms per block set
    constexpr int YS = NS | 1;          // odd row stride of the output-spectrum tile
    static_assert(R1 * R2 == M && NRG >= 1 && JPT >= 1, "geometry");
    static_assert(R2 % NRG == 0 || NRG % R2 == 0, "pass-1 mapping");
    using raw_t = typename RawT<FMT>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tile = reinterpret_cast<float2*>(smem_raw);                 // [M][max(RS, YS)]
    constexpr int TS = (RS > YS ? RS : YS);
    float2* tw = tile + (size_t)M * TS;                                 // [M] W_M^t

    const int tid = threadIdx.x;
    const int slot = tid % LW;
    const int rg = tid / LW;
    const int b_slot = slot / P;       // block within the set handled by this slot
    const int pl = slot % P;           // branch within the tile

    for (int i = tid; i < M; i += kThreads) tw[i] = p.twid[i];
    __syncthreads();

    const int D = p.decim;
    const int ntiles = (D + P - 1) / P;
    const raw_t* __restrict__ rawp = reinterpret_cast<const raw_t*>(p.raw);
    const int nsets = (p.nblocks + BT - 1) / BT;

    for (int set = blockIdx.x; set < nsets; set += gridDim.x) {
        const int blk0 = set * BT;
        float2 acc[JPT][CG][BT];
#pragma unroll
        for (int i = 0; i < JPT; ++i)
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[i][c][b] = make_float2(0.f, 0.f);

        // frame index (relative to p.raw) of row 0, branch 0 of this slot's block
        const bool blk_ok = (blk0 + b_slot) < p.nblocks;
        const int64_t row0 = p.mg_begin + (int64_t)(blk0 + b_slot) * p.ld - p.vd;
        const int64_t f_base = row0 * D - p.raw_n0;

        raw_t rawv[R1];
        auto fetch = [&](int t) {
            const int pb = t * P + pl;
            // pass 1 of this thread: sub-transform m2 (rows m = R2*m1 + m2)
            const int m2 = rg % R2;
            const int64_t f0 = f_base + pb + (int64_t)m2 * D;
            const int64_t step = (int64_t)R2 * D;
            const bool live = blk_ok && pb < D && rg < R2;
            const bool inside = live && f0 >= 0 && (f0 + step * (R1 - 1)) < p.raw_len;
            if (inside) {
#pragma unroll
                for (int i = 0; i < R1; ++i) rawv[i] = __ldg(rawp + (f0 + step * i));
            } else {
#pragma unroll
                for (int i = 0; i < R1; ++i) {
                    const int64_t f = f0 + step * i;
                    rawv[i] = (live && f >= 0 && f < p.raw_len) ? __ldg(rawp + f) : raw_zero<FMT>();
                }
            }
        };

        fetch(0);
        for (int t = 0; t < ntiles; ++t) {
            // ---------------- pass 1: R1-point DIF over m1, twiddle W_M^{m2*k1} ----------------
            if (rg < R2) {
                const int m2 = rg;
                float2 v[R1];
#pragma unroll
                for (int i = 0; i < R1; ++i) v[i] = raw_to_c64<FMT>(rawv[i], p.iq_swap, p.q_neg);
                dif<R1, +1>(v);
                static_for<R1>([&](auto kc) {
                    constexpr int k1 = decltype(kc)::value;
                    float2 x = v[bitrev<R1>(k1)];
                    if constexpr (k1 != 0) x = cmul(x, tw[(m2 * k1) & (M - 1)]);
                    tile[(k1 * R2 + m2) * RS + slot] = x;
                });
            }
            __syncthreads();
            if (t + 1 < ntiles) fetch(t + 1);
            // ---------------- pass 2: R2-point DIF over m2 (in place) --------------------------
            for (int k1 = rg; k1 < R1; k1 += NRG) {
                float2 v[R2];
                float2* base = tile + (k1 * R2) * RS + slot;
#pragma unroll
                for (int i = 0; i < R2; ++i) v[i] = base[i * RS];
                dif<R2, +1>(v);
                static_for<R2>([&](auto kc) {
                    constexpr int k2 = decltype(kc)::value;
                    base[k2 * RS] = v[bitrev<R2>(k2)];
                });
            }
            __syncthreads();
            // ---------------- multiply-accumulate against the channel spectra -------------------
#pragma unroll
            for (int q = 0; q < P; ++q) {
                const int pb = t * P + q;
                if (pb < D) {
#pragma unroll
                    for (int i = 0; i < JPT; ++i) {
                        const int j = tid + i * kThreads;
                        float2 g[CG];
#pragma unroll
                        for (int c = 0; c < CG; ++c) g[c] = __ldg(p.gtab + ((size_t)pb * CG + c) * M + j);
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            const float2 x = tile[j * RS + b * P + q];
#pragma unroll
                            for (int c = 0; c < CG; ++c) {
                                acc[i][c][b].x = fmaf(g[c].x, x.x, acc[i][c][b].x);
                                acc[i][c][b].x = fmaf(-g[c].y, x.y, acc[i][c][b].x);
                                acc[i][c][b].y = fmaf(g[c].x, x.y, acc[i][c][b].y);
                                acc[i][c][b].y = fmaf(g[c].y, x.x, acc[i][c][b].y);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---------------- output spectra -> shared, inverse transform ---------------------------
#pragma unroll
        for (int i = 0; i < JPT; ++i) {
            const int j = tid + i * kThreads;
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
                for (int c = 0; c < CG; ++c) tile[j * YS + b * CG + c] = acc[i][c][b];
        }
        inverse_and_store<M, CG>(tile, tw, p, blk0);
        __syncthreads();
    }
}


template <int M, int CG>
static size_t smem_bytes() {
    constexpr int RS = Geo<M>::LW + 1;
    constexpr int YS = (CG * kBlocksPerSet) | 1;
    constexpr int TS = RS > YS ? RS : YS;
    return ((size_t)M * TS + M) * sizeof(float2);
}

template <int M, int CG, int FMT>
static int launch_one(const ChannelizeParams& p, int n_sm, cudaStream_t st) {
    auto kern = k_channelize<M, CG, FMT>;
    const size_t smem = smem_bytes<M, CG>();
    static bool configured = false;
    if (!configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int nsets = (p.nblocks + kBlocksPerSet - 1) / kBlocksPerSet;
    const int grid = nsets < n_sm ? nsets : n_sm;   // persistent: one CTA per SM, block sets strided
    kern<<<grid, kThreads, smem, st>>>(p);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

template <int M, int CG>
static int launch_fmt(const ChannelizeParams& p, int codec, int n_sm, cudaStream_t st) {
    switch (codec) {
        case CODEC_S16: return launch_one<M, CG, CODEC_S16>(p, n_sm, st);
        case CODEC_U8: return launch_one<M, CG, CODEC_U8>(p, n_sm, st);
        case CODEC_F32: return launch_one<M, CG, CODEC_F32>(p, n_sm, st);
    }
    set_error("unknown codec %d", codec);
    return IQ2A_ERR_INVALID;
}

}  // namespace iq2a
