// Channel-bank kernel, second generation (int16 PCM, M = 512, D % 4 == 0): same math and same
// output as k_channelize (channelizer.cuh), restructured around what the first ncu capture showed
// (issue-bound at 41 % issue utilisation, long-scoreboard stalls on G and on spilled registers):
//
//  * PCM tiles arrive by TMA (cp.async.bulk.tensor.2d + mbarrier): 4 blocks x 512 rows x 8 branches
//    of raw int16 frames land in shared memory while the previous tile is in its pass-2 / MAC
//    phase.  No register prefetch -> no spills; out-of-range rows (stream start, tail) are
//    zero-filled by the TMA unit, so the kernel has no bounds checks at all.
//  * The 512-point branch transform is computed as two 256-point transforms (even rows E, odd rows
//    O) that live in the two halves of packed f32x2 registers: 16-point DIF x 16-point DIF entirely
//    in FADD2/FMUL2/FFMA2 (fft_pk.cuh), 128-bit shared-memory accesses; the last radix-2 stage
//    X[k] = E[k] + W^k O[k], X[k+256] = E[k] - W^k O[k] is fused into the MAC phase (4 FFMA).
//  * int16 -> float conversion only (I2F); the 1/32768 scale is folded into the G table.
//  * G is fetched two MAC steps ahead (the first two requests of a tile are issued before the barrier).
//
// STG = 1 replaces the TMA tensor copy by per-thread 4-byte cp.async copies (zero-filled out of range, completion
// on the same mbarrier): any D, any 4-byte aligned buffer, rows that are only partly resident -- the cases the
// tensor map cannot describe (row pitch D*4 B must be a multiple of 16 B for TMA).
//
// Shared memory: T[256][33] float4 (E.re,O.re,E.im,O.im) 135 168 B | PCM staging 4 x 16 512 B |
// twiddles W_512^t 4 KB | mbarrier.  Block b's PCM rows are
// stored rotated by b rows so that the four blocks of a warp's 32 slots hit different banks.
#pragma once
#include <cuda.h>

#include "channelizer.cuh"
#include "fft_pk.cuh"

namespace iq2a {

constexpr int kStageRows = 516;                    // 512 + rotation slack, 3 TMA boxes of 172 rows
constexpr int kBoxRows = 172;
constexpr int kStageBytes = kStageRows * 32;       // per block: rows of 8 frames x 4 B
constexpr int kTileBytes2 = 256 * 33 * 16;
// BT blocks per set: 4 -> 512 threads, one CTA per SM; 2 -> 256 threads and ~111 KB of shared memory, so
// TWO CTAs share an SM and the hardware overlaps one CTA's FFMA-bound MAC phase with the other's
// ALU/LSU-bound transform passes (each G element is then reused for 2 blocks instead of 4).
template <int BT>
struct Geo2 {
    static constexpr int NT = 128 * BT;            // threads
    static constexpr int LW = 8 * BT;              // lane slots per tile (BT blocks x 8 branches)
    static constexpr int RS = LW + 1;              // tile row stride (float4)
    static constexpr int BPT = 512 / NT;           // spectrum bins per thread in the MAC phase
    static constexpr size_t tile_bytes = (size_t)256 * RS * 16;
    static constexpr size_t smem = tile_bytes + (size_t)BT * kStageBytes + 512 * sizeof(float2) + 16;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared store that the compiler cannot fuse with its neighbour into a 128-bit store: the fused form
// needs its four source registers contiguous and costs four MOVs per store (seen in the ncu source view)
__device__ __forceinline__ void sts64(void* dst, unsigned long long v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(smem_u32(dst)), "l"(v) : "memory");
}
// The same with a shared-window address and a compile-time byte offset.  ptxas fuses the re / im stores of one tile
// element ([a + o], [a + o + 8]) into one STS.128 and lines the four source registers up with MOVs.  Both ways of
// avoiding that were measured on B200 and lost (cfg2, 60 s: fused 2.17 ms; two STS.64 kept apart by a run-time
// address gap 2.41 ms -- a half-warp then covers every other 8 bytes, a two-way bank conflict; real and imaginary
// parts in separate halves of the tile row, conflict-free STS.64 and no MOVs, 2.19 ms), so the fused form stays.
template <int OFF>
__device__ __forceinline__ void sts64_at(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared.b64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "l"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ ulonglong2 lds128_at(uint32_t addr) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2+%3];" : "=l"(v.x), "=l"(v.y) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pk_make_u(uint32_t lo, uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}


__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 4-byte asynchronous copy, zero-filled when src_bytes == 0
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// this thread's earlier cp.async copies arrive on the mbarrier when they land (count pre-charged at init)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// spectrum slot of thread j in the MAC phase -> DFT bin (must match k_build_g layout 1)
__host__ __device__ inline int slot_to_bin_v2(int j) {
    const int r = j & 255;
    return (r >> 4) + 16 * (r & 15) + 256 * (j >> 8);
}

template <int CG, int BT, int STG = 0>
__global__ void __launch_bounds__(Geo2<BT>::NT, 4 / BT)
k_channelize2(const ChannelizeParams p, const __grid_constant__ CUtensorMap tmap, const int64_t tmap_row0) {
    using G2 = Geo2<BT>;
    constexpr int P = 8, RS = G2::RS, NT = G2::NT, LW = G2::LW, BPT = G2::BPT;
    constexpr int NS = CG * BT, YS = NS | 1;
    static_assert((BT == 4 || BT == 2) && (size_t)512 * YS * sizeof(float2) <= G2::tile_bytes, "layout");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* T = reinterpret_cast<float4*>(smem_raw);
    unsigned char* stage = smem_raw + G2::tile_bytes;
    float2* tw512 = reinterpret_cast<float2*>(stage + BT * kStageBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw512 + 512);

    const int tid = threadIdx.x;
    const int slot = tid % LW, rg = tid / LW;          // 16 row groups either way
    const int b_slot = slot >> 3, pl = slot & 7;

    for (int i = tid; i < 512; i += NT) tw512[i] = p.twid[i];
    if (tid == 0) mbar_init(bar, STG == 0 ? 1 : NT);
    __syncthreads();

    const int D = p.decim;
    const int ntiles = (D + P - 1) / P;
    const int nsets = (p.nblocks + BT - 1) / BT;
    uint32_t parity = 0;

    // MAC-phase identity of this thread: row r of T; with 512 threads one half of the spectrum each
    // (bin k' or k'+256), with 256 threads both halves
    const int r_mac = tid & 255, hb0 = tid >> 8;
    const int kq = (r_mac >> 4) + 16 * (r_mac & 15);          // bin within the 256-point halves
    const float2 wq = p.twid[kq];                             // W_512^{k'}
    const float2 wc = (BPT == 1 && hb0) ? make_float2(-wq.x, -wq.y) : wq;
    const float2* __restrict__ gp = p.gtab + (BPT == 1 ? tid : r_mac);

    // int16 -> float without the (quarter-rate) I2F unit: flip the sign bit of both halves of a frame word (offset
    // binary), then PRMT each half under the exponent bytes 0x4B00: the float 8388608 + 32768 + s, and the bias
    // comes off with one packed add per E/O pair.  Q negation: flip the other 15 bits as well, which is the
    // offset-binary form of -q - 1, and take one less off.  The exponent bytes come from a register that already
    // holds such a float (the previous row's; a run-time seed for the first), never from a constant: with a
    // constant ptxas keeps the PRMT selector in a uniform register and spends a MOV per PRMT bringing it over.
    const uint32_t qmask = p.q_neg ? 0x7fffu : 0x8000u;
    const uint32_t xmask = p.iq_swap ? (0x80000000u | qmask) : ((qmask << 16) | 0x8000u);
    const pk_t bias_i = pk_bc(-8421376.0f), bias_q = pk_bc(p.q_neg ? -8421375.0f : -8421376.0f);
    const uint32_t exp_seed = 0x4B000000u + ((uint32_t)p.iq_swap >> 8);

    for (int set = blockIdx.x; set < nsets; set += gridDim.x) {
        const int blk0 = set * BT;
        // BPT == 1: acc[0][c][b] = (re, im) of this thread's bin.  BPT == 2: the two bins k', k'+256 share packed
        // registers: acc[0][c][b] = (re_k', re_k'+256), acc[1][c][b] = (im_k', im_k'+256) -> FFMA2 multiply-accumulate.
        float2 acc[BPT][CG][BT];
#pragma unroll
        for (int h = 0; h < BPT; ++h)
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[h][c][b] = make_float2(0.f, 0.f);

        auto issue = [&](int t) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, BT * kStageBytes);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                const int64_t row0 = p.mg_begin + (int64_t)(blk0 + b) * p.ld - p.vd;
                const int rt = (int)(row0 - tmap_row0) - b;   // rotation: block b stored b rows lower
#pragma unroll
                for (int i = 0; i < kStageRows / kBoxRows; ++i)
                    tma_load_2d(stage + b * kStageBytes + i * kBoxRows * 32, &tmap, t * P, rt + i * kBoxRows, bar);
            }
        };
        // STG == 1: every thread copies one frame column of 32-row sweeps; same staging layout as the TMA boxes
        auto issue_cp = [&](int t) {
            const int col = tid & 7, jr = tid >> 3;
            const bool col_ok = t * P + col < D;
            const uint32_t* raw = reinterpret_cast<const uint32_t*>(p.raw);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                const int64_t row0 = p.mg_begin + (int64_t)(blk0 + b) * p.ld - p.vd;
                int64_t f = (row0 - b + jr) * (int64_t)D + t * P + col - p.raw_n0;       // frame offset into p.raw
                uint32_t dst = smem_u32(stage + b * kStageBytes + jr * 32 + col * 4);
                for (int j = jr; j < kStageRows; j += NT / 8) {
                    const bool ok = col_ok && f >= 0 && f < p.raw_len;
                    cp_async4(dst, raw + (ok ? f : 0), ok ? 4 : 0);
                    f += (int64_t)(NT / 8) * D;
                    dst += (NT / 8) * 32;
                }
            }
            cp_async_arrive(bar);
        };
        if constexpr (STG == 0) {
            if (tid == 0) issue(0);
        } else {
            issue_cp(0);
        }

        for (int t = 0; t < ntiles; ++t) {
            mbar_wait(bar, parity);
            parity ^= 1;
            // ------------- pass 1: two 16-point DIFs (even / odd rows) per thread, packed ---------------
            {
                const int m2 = rg;
                const uint32_t* st = reinterpret_cast<const uint32_t*>(stage + b_slot * kStageBytes) + pl;
                pk_t re[16], im[16];
                auto unpack = [&](auto swapped) {
                    constexpr uint32_t sel_lo = 0x7610u, sel_hi = 0x7632u;      // half of a | bytes 2, 3 of b
                    constexpr uint32_t sel_i = decltype(swapped)::value ? sel_hi : sel_lo;
                    constexpr uint32_t sel_q = decltype(swapped)::value ? sel_lo : sel_hi;
                    uint32_t ex = exp_seed;
#pragma unroll
                    for (int m1 = 0; m1 < 16; ++m1) {
                        const int row = 32 * m1 + 2 * m2 + b_slot;
                        const uint32_t w0 = st[row * 8] ^ xmask, w1 = st[(row + 1) * 8] ^ xmask;
                        const uint32_t i0 = __byte_perm(w0, ex, sel_i), i1 = __byte_perm(w1, ex, sel_i);
                        const uint32_t q0 = __byte_perm(w0, ex, sel_q), q1 = __byte_perm(w1, ex, sel_q);
                        ex = i0;
                        re[m1] = pk_add(pk_make_u(i0, i1), bias_i);
                        im[m1] = pk_add(pk_make_u(q0, q1), bias_q);
                    }
                };
                if (p.iq_swap) unpack(std::true_type{});
                else unpack(std::false_type{});
                pk_dif<16>(re, im);
                const uint32_t dst = smem_u32(T) + (m2 * RS + slot) * 16;
                static_for<16>([&](auto kc) {
                    constexpr int k1 = decltype(kc)::value;
                    if constexpr (k1 != 0) {
                        // W_256^{m2 k1} = W_512^{2 m2 k1}: an 8-byte load that the two row groups of a warp
                        // share; ptxas turns the (w, w) pairs into scalar-broadcast operands of the packed ops
                        const pk_t xr = re[bitrev<16>(k1)], xi = im[bitrev<16>(k1)];
                        const float2 w = tw512[(2 * m2 * k1) & 511];
                        const pk_t wr = pk_make(w.x, w.x), wi = pk_make(w.y, w.y);
                        re[bitrev<16>(k1)] = pk_sub(pk_mul(xr, wr), pk_mul(xi, wi));
                        im[bitrev<16>(k1)] = pk_fma(xr, wi, pk_mul(xi, wr));
                    }
                });
                // the tile is free once every warp has left the previous multiply-accumulate phase; waiting for
                // that here, with this tile's pass 1 already computed, absorbs the skew between warps
                __syncthreads();
                static_for<16>([&](auto kc) {
                    constexpr int k1 = decltype(kc)::value;
                    sts64_at<16 * (k1 * 16) * RS>(dst, re[bitrev<16>(k1)]);
                    sts64_at<16 * (k1 * 16) * RS + 8>(dst, im[bitrev<16>(k1)]);
                });
            }
            __syncthreads();
            // (issuing the next tile's copy one barrier earlier, before this thread's stores, was slower: 2.26 vs
            // 2.17 ms on cfg2 -- the whole CTA then waits for thread 0 at the barrier below)
            if constexpr (STG == 0) {
                if (tid == 0 && t + 1 < ntiles) issue(t + 1);
            } else {
                if (t + 1 < ntiles) issue_cp(t + 1);
            }
            // ------------- pass 2: 16-point DIF over m2, in place ----------------------------------------
            {
                const int k1 = rg;
                const uint32_t col = smem_u32(T) + ((k1 * 16) * RS + slot) * 16;
                pk_t re[16], im[16];
                static_for<16>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const ulonglong2 v = lds128_at<16 * i * RS>(col);
                    re[i] = v.x;
                    im[i] = v.y;
                });
                pk_dif<16>(re, im);
                static_for<16>([&](auto kc) {
                    constexpr int k2 = decltype(kc)::value;
                    sts64_at<16 * k2 * RS>(col, re[bitrev<16>(k2)]);
                    sts64_at<16 * k2 * RS + 8>(col, im[bitrev<16>(k2)]);
                });
            }
            // G for the first two steps of this tile is requested before the barrier so that the L2 round
            // trip overlaps the wait; inside the loop the table is fetched two steps ahead.
            // BPT == 2 reads table layout 2: per (p, c, r) one float4 (g_k'.x, g_k'+256.x, g_k'.y, g_k'+256.y), i.e.
            // g[0][c] = packed real parts, g[1][c] = packed imaginary parts of the two bins
            auto gload = [&](float2 (&g)[BPT][CG], int pb) {
                const int pn = min(pb, D - 1);
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    if constexpr (BPT == 1) {
                        g[0][c] = __ldg(gp + ((size_t)pn * CG + c) * 512);
                    } else {
                        // L2 only: a table line is used once per CTA, and the two CTAs of an SM are not in step
                        const float4 v = __ldcg(reinterpret_cast<const float4*>(p.gtab) + ((size_t)pn * CG + c) * 256 + r_mac);
                        g[0][c] = make_float2(v.x, v.y);
                        g[BPT - 1][c] = make_float2(v.z, v.w);
                    }
                }
            };
            float2 g0[BPT][CG], g1[BPT][CG];
            gload(g0, t * P);
            gload(g1, t * P + 1);
            __syncthreads();
            // ------------- last radix-2 stage fused with the multiply-accumulate --------------------------
            const float4* trow = T + r_mac * RS;
            auto mac_step = [&](int q, const float2 (&g)[BPT][CG]) {
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    const float4 v = trow[b * P + q];                    // (E.re, O.re, E.im, O.im)
                    if constexpr (BPT == 1) {
                        const float xr = fmaf(wc.x, v.y, fmaf(-wc.y, v.w, v.x));
                        const float xi = fmaf(wc.x, v.w, fmaf(wc.y, v.y, v.z));
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            acc[0][c][b].x = fmaf(g[0][c].x, xr, acc[0][c][b].x);
                            acc[0][c][b].x = fmaf(-g[0][c].y, xi, acc[0][c][b].x);
                            acc[0][c][b].y = fmaf(g[0][c].x, xi, acc[0][c][b].y);
                            acc[0][c][b].y = fmaf(g[0][c].y, xr, acc[0][c][b].y);
                        }
                    } else {
                        // X[k'] = E + W O, X[k'+256] = E - W O, both bins in one packed register pair
                        const float pr = fmaf(wc.x, v.y, -wc.y * v.w), pi = fmaf(wc.x, v.w, wc.y * v.y);
                        const pk_t xr = pk_make(v.x + pr, v.x - pr), xi = pk_make(v.z + pi, v.z - pi);
                        const pk_t nxi = xi ^ 0x8000000080000000ull;      // sign flips: off the FMA pipe
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            const pk_t gre = pk_make(g[0][c].x, g[0][c].y), gim = pk_make(g[BPT - 1][c].x, g[BPT - 1][c].y);
                            pk_t are = pk_make(acc[0][c][b].x, acc[0][c][b].y), aim = pk_make(acc[BPT - 1][c][b].x, acc[BPT - 1][c][b].y);
                            are = pk_fma(gre, xr, are);
                            are = pk_fma(gim, nxi, are);
                            aim = pk_fma(gre, xi, aim);
                            aim = pk_fma(gim, xr, aim);
                            pk_split(are, acc[0][c][b].x, acc[0][c][b].y);
                            pk_split(aim, acc[BPT - 1][c][b].x, acc[BPT - 1][c][b].y);
                        }
                    }
                }
            };
            if (t * P + P <= D) {
                // full tile: statically rotated register sets, no copies
                float2 g2[BPT][CG];
                static_for<P>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    if constexpr (q % 3 == 0) { gload(g2, t * P + q + 2); mac_step(q, g0); }
                    else if constexpr (q % 3 == 1) { gload(g0, t * P + q + 2); mac_step(q, g1); }
                    else { gload(g1, t * P + q + 2); mac_step(q, g2); }
                });
            } else {
                for (int q = 0; q < P && t * P + q < D; ++q) {
                    float2 g2[BPT][CG];
                    gload(g2, t * P + q + 2);
                    mac_step(q, g0);
#pragma unroll
                    for (int h = 0; h < BPT; ++h)
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            g0[h][c] = g1[h][c];
                            g1[h][c] = g2[h][c];
                        }
                }
            }
        }
        __syncthreads();                      // every warp is done reading the tile

        // ------------- output spectra -> shared (layout of the shared inverse), inverse, store -------------
        float2* ytile = reinterpret_cast<float2*>(T);
#pragma unroll
        for (int h = 0; h < BPT; ++h) {
            const int kbin = kq + 256 * (BPT == 1 ? hb0 : h);
            const int yrow = (kbin & 31) * 16 + (kbin >> 5);      // slot the inverse transform expects
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    if constexpr (BPT == 1) ytile[yrow * YS + b * CG + c] = acc[0][c][b];
                    else ytile[yrow * YS + b * CG + c] = h == 0 ? make_float2(acc[0][c][b].x, acc[BPT - 1][c][b].x)
                                                                : make_float2(acc[0][c][b].y, acc[BPT - 1][c][b].y);
                }
        }
        inverse_and_store<512, CG, BT, NT>(ytile, tw512, p, blk0);
        __syncthreads();
    }
}

template <int CG, int BT, int STG = 0>
static int launch_channelize2_cg(const ChannelizeParams& p, const CUtensorMap& tmap, int64_t tmap_row0, int n_sm,
                                 cudaStream_t st) {
    auto kern = k_channelize2<CG, BT, STG>;
    static bool configured = false;
    if (!configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo2<BT>::smem));
        configured = true;
    }
    const int nsets = (p.nblocks + BT - 1) / BT;
    const int slots = n_sm * (4 / BT);                      // persistent: one (BT=4) or two (BT=2) CTAs per SM
    const int grid = nsets < slots ? nsets : slots;
    kern<<<grid, Geo2<BT>::NT, Geo2<BT>::smem, st>>>(p, tmap, tmap_row0);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
