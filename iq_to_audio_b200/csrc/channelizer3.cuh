// Channel-bank kernel, third generation: the arithmetic of k_channelize2 (channelizer2.cuh), but the
// two halves of the work run CONCURRENTLY on different warps instead of alternating behind CTA-wide
// barriers.  The second ncu capture showed why: the transform passes (FADD2 / LOP3 / LDS-STS bound,
// FMA pipe mostly idle) and the multiply-accumulate phase (91 % FFMA, FMA pipe saturated) each had
// the SM to themselves for ~45 % of the time, so neither the issue slots nor the FMA pipe were ever
// more than ~45 % busy.
//
//   warps 0..7   "transform group"  per tile: TMA-staged int16 -> packed 16x16-point DIF -> tile T[s]
//   warps 8..15  "MAC group"        per tile: last radix-2 stage + multiply-accumulate from T[s]
//
// T and the PCM staging buffer are double-buffered (tile = 16 lane slots: 4 blocks x 4 branches);
// the groups hand tiles over through mbarriers (T_full / T_empty), the transform group synchronises
// internally with a named barrier, TMA loads run two tiles ahead.  Registers are re-partitioned with
// setmaxnreg (transform warps 96, MAC warps 160: the MAC threads own two spectrum bins each, i.e.
// 2 x CG x 4 complex accumulators) and given back for the common inverse-transform epilogue.
#pragma once
#include "channelizer2.cuh"

namespace iq2a {

constexpr int kT3Slots = 16;                        // 4 blocks x 4 branches
constexpr int kT3RS = 17;                           // tile row stride (float4)
constexpr int kT3TileBytes = 256 * kT3RS * 16;      // 69 632
constexpr int kT3BoxRows = 136;                     // 136 * 16 B = 17 * 128 B
constexpr int kT3Boxes = 4;                         // 544 rows >= 512 + 3 rotation rows
constexpr int kT3StageBlock = kT3BoxRows * kT3Boxes * 16;   // 8 704 B per block
constexpr int kT3StageBytes = 4 * kT3StageBlock;    // 34 816 B per stage
constexpr size_t kSmem3 = 2 * (size_t)kT3TileBytes + 2 * (size_t)kT3StageBytes + 512 * sizeof(float2) +
                          256 * sizeof(float4) + 64;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait with a watchdog: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(kThreads, 1)
k_channelize3(const ChannelizeParams p, const __grid_constant__ CUtensorMap tmap, const int64_t tmap_row0) {
    constexpr int BT = kBlocksPerSet, P = 4, RS = kT3RS;
    constexpr int NS = CG * BT, YS = NS | 1;
    static_assert(BT == 4 && (size_t)512 * YS * sizeof(float2) <= 2 * (size_t)kT3TileBytes, "layout");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* Tbuf[2] = {reinterpret_cast<float4*>(smem_raw), reinterpret_cast<float4*>(smem_raw + kT3TileBytes)};
    unsigned char* stage[2] = {smem_raw + 2 * kT3TileBytes, smem_raw + 2 * kT3TileBytes + kT3StageBytes};
    float2* tw512 = reinterpret_cast<float2*>(smem_raw + 2 * kT3TileBytes + 2 * kT3StageBytes);
    float4* tw256b = reinterpret_cast<float4*>(tw512 + 512);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw256b + 256);
    uint64_t* raw_full = bars;        // [2]
    uint64_t* t_full = bars + 2;      // [2]
    uint64_t* t_empty = bars + 4;     // [2]

    const int tid = threadIdx.x;
    const bool is_mac = tid >= 256;

    for (int i = tid; i < 512; i += kThreads) tw512[i] = p.twid[i];
    for (int i = tid; i < 256; i += kThreads) {
        const float2 w = p.twid[2 * i];
        tw256b[i] = make_float4(w.x, w.x, w.y, w.y);
    }
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(raw_full + s, 1);
            mbar_init(t_full + s, 256);
            mbar_init(t_empty + s, 256);
        }
    }
    __syncthreads();

    const int D = p.decim;
    const int ntiles = (D + P - 1) / P;
    const int nsets = (p.nblocks + BT - 1) / BT;
    uint32_t gt = 0;                                   // running tile counter (buffer = gt & 1, use = gt >> 1)

    for (int set = blockIdx.x; set < nsets; set += gridDim.x) {
        const int blk0 = set * BT;
        float2* ytile = reinterpret_cast<float2*>(smem_raw);

        if (!is_mac) {
            // =========================== transform group ===========================================
            asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
            const int slot = tid & 15, hw = tid >> 4;
            const int b_slot = slot >> 2, pl = slot & 3;
            // half-warps of one warp take sub-transforms m2 and m2+2: their PCM rows then sit 16 banks apart
            const int m2 = ((hw & 1) << 1) | ((hw >> 1) & 1) | (hw & ~3);
            auto issue = [&](int t, int s) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(raw_full + s, kT3StageBytes);
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    const int64_t row0 = p.mg_begin + (int64_t)(blk0 + b) * p.ld - p.vd;
                    const int rt = (int)(row0 - tmap_row0) - b;
#pragma unroll
                    for (int i = 0; i < kT3Boxes; ++i)
                        tma_load_2d(stage[s] + b * kT3StageBlock + i * kT3BoxRows * 16, &tmap, t * P, rt + i * kT3BoxRows,
                                    raw_full + s);
                }
            };
            if (tid == 0) {
                issue(0, gt & 1);
                if (ntiles > 1) issue(1, (gt + 1) & 1);
            }
            uint32_t g = gt;
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int s = g & 1;
                const uint32_t par = (g >> 1) & 1;
                float4* T = Tbuf[s];
                mbar_wait_wd(raw_full + s, par);
                mbar_wait_wd(t_empty + s, par ^ 1);
                {
                    const uint32_t* st = reinterpret_cast<const uint32_t*>(stage[s] + b_slot * kT3StageBlock) + pl;
                    pk_t re[16], im[16];
#pragma unroll
                    for (int m1 = 0; m1 < 16; ++m1) {
                        const int row = 32 * m1 + 2 * m2 + b_slot;
                        uint32_t w0 = st[row * 4], w1 = st[(row + 1) * 4];
                        if (p.iq_swap) {
                            w0 = __funnelshift_l(w0, w0, 16);
                            w1 = __funnelshift_l(w1, w1, 16);
                        }
                        const pk_t ui = pk_make(__uint_as_float((w0 & 0xffffu) ^ 0x4B008000u),
                                                __uint_as_float((w1 & 0xffffu) ^ 0x4B008000u));
                        const pk_t uq = pk_make(__uint_as_float(__byte_perm(w0, 0x4B00u, 0x5432) ^ 0x8000u),
                                                __uint_as_float(__byte_perm(w1, 0x4B00u, 0x5432) ^ 0x8000u));
                        re[m1] = pk_add(ui, pk_bc(-8421376.0f));
                        im[m1] = p.q_neg ? pk_sub(pk_bc(8421376.0f), uq) : pk_add(uq, pk_bc(-8421376.0f));
                    }
                    pk_dif<16>(re, im);
                    pk_t* dst = reinterpret_cast<pk_t*>(T) + 2 * (m2 * RS + slot);
                    static_for<16>([&](auto kc) {
                        constexpr int k1 = decltype(kc)::value;
                        pk_t xr = re[bitrev<16>(k1)], xi = im[bitrev<16>(k1)];
                        if constexpr (k1 != 0) {
                            const float4 w = tw256b[(m2 * k1) & 255];
                            const pk_t wr = pk_make(w.x, w.y), wi = pk_make(w.z, w.w);
                            const pk_t nr = pk_sub(pk_mul(xr, wr), pk_mul(xi, wi));
                            const pk_t ni = pk_fma(xr, wi, pk_mul(xi, wr));
                            xr = nr;
                            xi = ni;
                        }
                        sts64(dst + 2 * (k1 * 16) * RS, xr);
                        sts64(dst + 2 * (k1 * 16) * RS + 1, xi);
                    });
                }
                named_bar_sync(1, 256);
                if (tid == 0 && t + 2 < ntiles) issue(t + 2, s);       // stage[s] is consumed: refill two tiles ahead
                {
                    const int k1 = hw;
                    ulonglong2* base = reinterpret_cast<ulonglong2*>(T) + (k1 * 16) * RS + slot;
                    pk_t re[16], im[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const ulonglong2 v = base[i * RS];
                        re[i] = v.x;
                        im[i] = v.y;
                    }
                    pk_dif<16>(re, im);
                    pk_t* base64 = reinterpret_cast<pk_t*>(base);
                    static_for<16>([&](auto kc) {
                        constexpr int k2 = decltype(kc)::value;
                        sts64(base64 + 2 * k2 * RS, re[bitrev<16>(k2)]);
                        sts64(base64 + 2 * k2 * RS + 1, im[bitrev<16>(k2)]);
                    });
                }
                mbar_arrive(t_full + s);
            }
            __syncthreads();                                          // (A) all tiles transformed and consumed
            __syncthreads();                                          // (B) output spectra are in shared memory
            asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");     // after the MAC warps gave theirs back
        } else {
            // =========================== MAC group ================================================
            asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
            const int r = tid - 256;                                  // row of T == bin within the 256-point halves
            const int kq = (r >> 4) + 16 * (r & 15);
            const float2 wc = p.twid[kq];                             // W_512^{k'}
            float2 acc[2][CG][BT];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int c = 0; c < CG; ++c)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[h][c][b] = make_float2(0.f, 0.f);
            // G layout [p][c][j], j = r (bin k') and r + 256 (bin k' + 256): slot_to_bin_v2
            const float2* __restrict__ gp = p.gtab + r;
            auto gload = [&](float2 (&gg)[2][CG], int pb) {
                const int pn = min(pb, D - 1);
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    gg[0][c] = __ldg(gp + ((size_t)pn * CG + c) * 512);
                    gg[1][c] = __ldg(gp + ((size_t)pn * CG + c) * 512 + 256);
                }
            };
            float2 ga[2][CG], gb[2][CG];
            gload(ga, 0);
            uint32_t g = gt;
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int s = g & 1;
                const uint32_t par = (g >> 1) & 1;
                const float4* trow = Tbuf[s] + r * RS;
                mbar_wait_wd(t_full + s, par);
                auto mac_step = [&](int q, const float2 (&gg)[2][CG]) {
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        const float4 v = trow[b * P + q];             // (E.re, O.re, E.im, O.im)
                        const float pr = fmaf(wc.x, v.y, -wc.y * v.w);
                        const float pi = fmaf(wc.x, v.w, wc.y * v.y);
                        const float x0r = v.x + pr, x0i = v.z + pi, x1r = v.x - pr, x1i = v.z - pi;
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            acc[0][c][b].x = fmaf(gg[0][c].x, x0r, acc[0][c][b].x);
                            acc[0][c][b].x = fmaf(-gg[0][c].y, x0i, acc[0][c][b].x);
                            acc[0][c][b].y = fmaf(gg[0][c].x, x0i, acc[0][c][b].y);
                            acc[0][c][b].y = fmaf(gg[0][c].y, x0r, acc[0][c][b].y);
                            acc[1][c][b].x = fmaf(gg[1][c].x, x1r, acc[1][c][b].x);
                            acc[1][c][b].x = fmaf(-gg[1][c].y, x1i, acc[1][c][b].x);
                            acc[1][c][b].y = fmaf(gg[1][c].x, x1i, acc[1][c][b].y);
                            acc[1][c][b].y = fmaf(gg[1][c].y, x1r, acc[1][c][b].y);
                        }
                    }
                };
                if (t * P + P <= D) {
                    static_for<P>([&](auto qc) {
                        constexpr int q = decltype(qc)::value;
                        if constexpr (q % 2 == 0) { gload(gb, t * P + q + 1); mac_step(q, ga); }
                        else { gload(ga, t * P + q + 1); mac_step(q, gb); }
                    });
                } else {
                    for (int q = 0; q < P && t * P + q < D; ++q) {
                        gload(gb, t * P + q + 1);
                        mac_step(q, ga);
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            ga[0][c] = gb[0][c];
                            ga[1][c] = gb[1][c];
                        }
                    }
                }
                mbar_arrive(t_empty + s);
            }
            __syncthreads();                                          // (A)
            // output spectra -> shared in the layout of the shared inverse (bin k -> slot (k&31)*16 + (k>>5))
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kbin = kq + 256 * h;
                const int yrow = (kbin & 31) * 16 + (kbin >> 5);
#pragma unroll
                for (int b = 0; b < BT; ++b)
#pragma unroll
                    for (int c = 0; c < CG; ++c) ytile[yrow * YS + b * CG + c] = acc[h][c][b];
            }
            asm volatile("setmaxnreg.dec.sync.aligned.u32 128;");
            __syncthreads();                                          // (B)
        }
        gt += ntiles;
        inverse_and_store<512, CG>(ytile, tw512, p, blk0);
        __syncthreads();
    }
}

template <int CG>
static int launch_channelize3_cg(const ChannelizeParams& p, const CUtensorMap& tmap, int64_t tmap_row0, int n_sm,
                                 cudaStream_t st) {
    auto kern = k_channelize3<CG>;
    static bool configured = false;
    if (!configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem3));
        configured = true;
    }
    const int nsets = (p.nblocks + kBlocksPerSet - 1) / kBlocksPerSet;
    const int grid = nsets < n_sm ? nsets : n_sm;
    kern<<<grid, kThreads, kSmem3, st>>>(p, tmap, tmap_row0);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
