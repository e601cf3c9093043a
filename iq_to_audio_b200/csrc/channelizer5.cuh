// Channel-bank kernel, fifth generation: the polyphase branches are processed in MIRROR PAIRS.
//
// The reference designs its channel filter with scipy.signal.firwin (src/iq_to_audio/processing.py:599-620):
// real, symmetric taps h[n] = h[L-1-n].  With A = ceil((L-1)/D) and r = A*D - (L-1), branch p and branch
// p' = r - p (p <= r, "class 1", Q = A) or p' = r + D - p (p > r, "class 2", Q = A + 1) hold time-reversed copies of
// the same taps, and their spectra obey
//     G[c,p',k] = kappa_c * W_M^{Qk} * conj(G[c,p,k]),      kappa_c = e^{-j w_c (L-1)},  W_M = e^{-2 pi j / M}.
// Writing  kappa_c^{-1/2} G[c,p,k] = a + j b  and  X~_{p'}[k] = W_M^{Qk} X_{p'}[k]  (the transform of branch p'
// with its block window rotated circularly by Q rows -- a placement choice of the TMA copy, no arithmetic):
//     G_p X_p + G_p' X_p' = kappa_c^{1/2} * ( a (X_p + X~_p') + j b (X_p - X~_p') ).
// One TABLE ENTRY and FOUR real multiply-adds serve TWO branch bins (the unpaired form needs two entries and
// eight), and kappa_c^{1/2} joins the NCO rotation of the epilogue.  Same result as generation 4 (the identity is
// exact), half the table traffic through L1 and 2/3 of the multiply-accumulate instructions
// (plan.py: pair_tiles / emulate_block_math_paired; tools/tma_issue_probe.cu for the copy rules used below).
//
// The tensor copy moves 16-byte units whose first column is a multiple of 4 frames (measured: any other column is
// an illegal instruction), so a tile is two ALIGNED column groups: forward group u = columns 4u..4u+3 and mirror
// group w, chosen so that column 4u+i mirrors column 4w+4-i for i = 1, 2, 3 (this needs r % 4 == 0, which
// (L-1) % 4 == 0 gives).  The two columns 4u and 4w that are left over are the mirrors of the left-over columns of
// the NEIGHBOURING tiles (4u mirrors 4(w+1), the previous tile's): the spectrum of column 4w is carried to the next
// tile in a small shared buffer `orph` (every thread only ever touches its own bins there), so a tile is
//   3 pairs from the tile itself + 1 pair (column 4u, carried column) = 4 table entries for 8 columns x 2 blocks.
// The first tile of a class has nothing carried (a zero partner turns the pair step into the plain complex
// product) and the column left at the end of a class -- its own mirror -- takes one extra step per class.
//
// A tile is 16 transform slots like generation 4's 8 branches x 2 blocks: slot = block*8 + side*4 + col.
// Staging (raw int16 frames, rows of 4 frames = 16 B): four regions [block][side] of kRegionRows rows; window row j
// of a forward strip sits in staging row j + df, rotated row j of a mirror strip in row j + dm, where dm =
// (-Q) mod 8 makes the seam of the rotation (row dm + Q) 128-byte aligned as the tensor copy requires, df = dm + 1,
// and block 1 threads take the row group m2 + 2: the eight 16-byte groups of a warp's load then fall in eight
// different bank groups.  Mirror strip = one box of Q + dm rows (the wrapped rows) + two boxes for the linear part.
#pragma once
#include <cuda.h>

#include "channelizer2.cuh"
#include "pair_maps.cuh"

namespace iq2a {

struct Geo5 {
    static constexpr int NT = 256, RS = 17;
    static constexpr size_t tile_bytes = (size_t)256 * RS * 16;
    static constexpr size_t orph_bytes = 2 * 256 * 16;           // carried spectrum: [block][row] (E.re, O.re, E.im, O.im)
    // (the mbarrier, the two control words and the epilogue's 12 block phasors live in the first 16 entries of the
    // twiddle table, which belong to k1 = 0 and are never read: 2 x (smem + 1 KB reserved) is exactly the 228 KB of an SM)
    static constexpr size_t smem = tile_bytes + 4 * (size_t)kRegionBytes + orph_bytes + 512 * sizeof(float2);
};

// ---- pieces shared by the fused kernel (k_channelize5) and the split kernels (channelizer5s.cuh) ----------------

// int16 -> float without the (quarter-rate) I2F unit: see channelizer2.cuh
struct Unpack5 {
    uint32_t xmask, exp_seed;
    pk_t bias_q;
    int iq_swap;
};
__device__ __forceinline__ Unpack5 make_unpack5(const ChannelizeParams& p) {
    const uint32_t qmask = p.q_neg ? 0x7fffu : 0x8000u;
    Unpack5 u;
    u.xmask = p.iq_swap ? (0x80000000u | qmask) : ((qmask << 16) | 0x8000u);
    u.bias_q = pk_bc(p.q_neg ? -8421375.0f : -8421376.0f);
    u.exp_seed = 0x4B000000u + ((uint32_t)p.iq_swap >> 8);
    u.iq_swap = p.iq_swap;
    return u;
}

// The copy of tile t of the block set starting at block blk0 (one thread): per block 3 forward boxes, the wrap
// box and 2 linear boxes of the mirror strip (header comment: staging).
__device__ __forceinline__ void issue_tile5(unsigned char* stage, uint64_t* bar, const PairMaps& maps, const PairGeo& geo,
                                            const ChannelizeParams& p, int64_t tmap_row0, int blk0, int t) {
    const int cls = t >= geo.tiles1;
    const int j = cls ? t - geo.tiles1 : t;
    const int fwd = 4 * ((cls ? geo.g1 : 0) + j);                                  // forward group u
    const int mir = 4 * ((cls ? geo.g1 + geo.g2 : geo.g1) - 1 - j);                // mirror group w
    const int dm = cls ? geo.dm[1] : geo.dm[0], hw = cls ? geo.hw[1] : geo.hw[0], hl = cls ? geo.hl[1] : geo.hl[0];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#ifdef IQ2A_EXPERIMENT_NO_COPY          // timing experiment only (garbage results): what the copies cost the SM
    mbar_arrive(bar);
    return;
#endif
    mbar_expect_tx(bar, (uint32_t)(2 * (3 * kFwdBoxRows + hw + 2 * hl) * 16));
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int64_t row0 = p.mg_begin + (int64_t)(blk0 + b) * p.ld - p.vd;
        const int rw = (int)(row0 - tmap_row0);                // tensor row of window row 0
        unsigned char* rf = stage + (b * 2) * kRegionBytes;
        unsigned char* rm = rf + kRegionBytes;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            tma_load_2d(rf + k * kFwdBoxRows * 16, &maps.fwd, fwd, rw - (dm + 1) + k * kFwdBoxRows, bar);
        // (one cp.async.bulk.tensor per descriptor in the source: the descriptor address stays uniform)
        if (cls) {
            // staging rows [0, hw) <- window rows [512 - hw, 512): rows dm.. are the wrapped part of the rotation
            tma_load_2d(rm, &maps.wrap[1], mir, rw + 512 - hw, bar);
            // staging rows [hw, hw + 2 hl) <- window rows [0, 2 hl)
            tma_load_2d(rm + hw * 16, &maps.lin[1], mir, rw, bar);
            tma_load_2d(rm + (hw + hl) * 16, &maps.lin[1], mir, rw + hl, bar);
        } else {
            tma_load_2d(rm, &maps.wrap[0], mir, rw + 512 - hw, bar);
            tma_load_2d(rm + hw * 16, &maps.lin[0], mir, rw, bar);
            tma_load_2d(rm + (hw + hl) * 16, &maps.lin[0], mir, rw + hl, bar);
        }
    }
}

// Pass 1 of a tile: two 16-point DIFs (even / odd rows) per thread, packed, twiddle, store to T[k1*16 + m2][slot].
// `after_load` runs once the thread's rows are in registers (the staging buffer is free for this warp),
// `before_store` before the first store (the tile must be free).
template <class F1, class F2>
__device__ __forceinline__ void pass1_5(const uint32_t* st, int m2p1, int slot, float4* T, const float2* tw256,
                                        const Unpack5& u, F1&& after_load, F2&& before_store) {
    constexpr int RS = 17;
    const pk_t bias_i = pk_bc(-8421376.0f);
    pk_t re[16], im[16];
    auto unpack = [&](auto swapped) {
        constexpr uint32_t sel_lo = 0x7610u, sel_hi = 0x7632u;
        constexpr uint32_t sel_i = decltype(swapped)::value ? sel_hi : sel_lo;
        constexpr uint32_t sel_q = decltype(swapped)::value ? sel_lo : sel_hi;
        uint32_t ex = u.exp_seed;
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) {
            const int row = 32 * m1 + 2 * m2p1;
            const uint32_t w0 = st[row * 4] ^ u.xmask, w1 = st[(row + 1) * 4] ^ u.xmask;
            const uint32_t i0 = __byte_perm(w0, ex, sel_i), i1 = __byte_perm(w1, ex, sel_i);
            const uint32_t q0 = __byte_perm(w0, ex, sel_q), q1 = __byte_perm(w1, ex, sel_q);
            ex = i0;
            re[m1] = pk_add(pk_make_u(i0, i1), bias_i);
            im[m1] = pk_add(pk_make_u(q0, q1), u.bias_q);
        }
    };
    if (u.iq_swap) unpack(std::true_type{});
    else unpack(std::false_type{});
    after_load();
    pk_dif<16>(re, im);
    const uint32_t dst = smem_u32(T) + (m2p1 * RS + slot) * 16;
    static_for<16>([&](auto kc) {
        constexpr int k1 = decltype(kc)::value;
        if constexpr (k1 != 0) {
            // W_256^{m2 k1}: an 8-byte load that the row groups of a warp share; ptxas turns the (w, w) pairs into
            // scalar-broadcast operands of the packed ops
            const pk_t xr = re[bitrev<16>(k1)], xi = im[bitrev<16>(k1)];
            const float2 w = tw256[k1 * 16 + m2p1];
            const pk_t wr = pk_make(w.x, w.x), wi = pk_make(w.y, w.y);
            re[bitrev<16>(k1)] = pk_sub(pk_mul(xr, wr), pk_mul(xi, wi));
            im[bitrev<16>(k1)] = pk_fma(xr, wi, pk_mul(xi, wr));
        }
    });
    before_store();
    static_for<16>([&](auto kc) {
        constexpr int k1 = decltype(kc)::value;
        sts64_at<16 * (k1 * 16) * RS>(dst, re[bitrev<16>(k1)]);
        sts64_at<16 * (k1 * 16) * RS + 8>(dst, im[bitrev<16>(k1)]);
    });
}

// Pass 2 of a tile: 16-point DIF over m2, in place (thread: column `slot`, rows k1*16 .. k1*16+15).
__device__ __forceinline__ void pass2_5(float4* T, int k1, int slot) {
    constexpr int RS = 17;
    const uint32_t colp = smem_u32(T) + ((k1 * 16) * RS + slot) * 16;
    pk_t re[16], im[16];
    static_for<16>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const ulonglong2 v = lds128_at<16 * i * RS>(colp);
        re[i] = v.x;
        im[i] = v.y;
    });
    pk_dif<16>(re, im);
    static_for<16>([&](auto kc) {
        constexpr int k2 = decltype(kc)::value;
        sts64_at<16 * k2 * RS>(colp, re[bitrev<16>(k2)]);
        sts64_at<16 * k2 * RS + 8>(colp, im[bitrev<16>(k2)]);
    });
}

// X[k'] = E + W O, X[k'+256] = E - W O of a packed (E, O) spectrum: both bins in one register pair
__device__ __forceinline__ void radix2_5(float2 wc, pk_t e_o_re, pk_t e_o_im, pk_t& xr, pk_t& xi) {
    float e_r, o_r, e_i, o_i;
    pk_split(e_o_re, e_r, o_r);
    pk_split(e_o_im, e_i, o_i);
    const float pr = fmaf(wc.x, o_r, -wc.y * o_i), pi = fmaf(wc.x, o_i, wc.y * o_r);
    xr = pk_make(e_r + pr, e_r - pr);
    xi = pk_make(e_i + pi, e_i - pi);
}

// acc += a (Xf + Xm) + j b (Xf - Xm) for block b: table entry (a_k', a_k'+256, b_k', b_k'+256) per channel
template <int CG, int BT>
__device__ __forceinline__ void pair_mac5(pk_t (&acc)[2][CG][BT], int b, float2 wc, ulonglong2 vf, ulonglong2 vm,
                                          const float4 (&g)[CG]) {
    {
        pk_t sr, si;
        radix2_5(wc, pk_add(vf.x, vm.x), pk_add(vf.y, vm.y), sr, si);
#pragma unroll
        for (int c = 0; c < CG; ++c) {
            const pk_t ga = pk_make(g[c].x, g[c].y);
            acc[0][c][b] = pk_fma(ga, sr, acc[0][c][b]);
            acc[1][c][b] = pk_fma(ga, si, acc[1][c][b]);
        }
    }
    {
        pk_t dr, di;
        radix2_5(wc, pk_sub(vf.x, vm.x), pk_sub(vf.y, vm.y), dr, di);
        const pk_t ndi = di ^ 0x8000000080000000ull;       // sign flips: off the FMA pipe
#pragma unroll
        for (int c = 0; c < CG; ++c) {
            const pk_t gb = pk_make(g[c].z, g[c].w);
            acc[0][c][b] = pk_fma(gb, ndi, acc[0][c][b]);
            acc[1][c][b] = pk_fma(gb, dr, acc[1][c][b]);
        }
    }
}

// acc += e X for block b (plain complex product; entry (e.re_k', e.re_k'+256, e.im_k', e.im_k'+256)): the column
// left over at the end of a class
template <int CG, int BT>
__device__ __forceinline__ void single_mac5(pk_t (&acc)[2][CG][BT], int b, float2 wc, ulonglong2 v, const float4 (&g)[CG]) {
    pk_t xr, xi;
    radix2_5(wc, v.x, v.y, xr, xi);
    const pk_t nxi = xi ^ 0x8000000080000000ull;
#pragma unroll
    for (int c = 0; c < CG; ++c) {
        const pk_t er = pk_make(g[c].x, g[c].y), ei = pk_make(g[c].z, g[c].w);
        acc[0][c][b] = pk_fma(er, xr, acc[0][c][b]);
        acc[0][c][b] = pk_fma(ei, nxi, acc[0][c][b]);
        acc[1][c][b] = pk_fma(er, xi, acc[1][c][b]);
        acc[1][c][b] = pk_fma(ei, xr, acc[1][c][b]);
    }
}

// Epilogue of a block set: inverse 512-point transforms of the CG x 2 output spectra, overlap rows dropped, NCO
// rotation, complex64 store.  The thread that owns bins (k', k'+256) first does the radix-2 step of the inverse
//   y[2n'+e] = sum_k' (Y[k'] + (-1)^e Y[k'+256]) W_512^{-e k'} W_256^{-n' k'},
// which leaves two 256-point inverse transforms per spectrum (even and odd output rows); those ride in the two
// halves of packed f32x2 registers through two 16-point passes exactly like the forward transforms (the inverse
// of a packed DIF is the forward DIF with real and imaginary registers exchanged).  Row rho of the result holds
// output rows 2 rho and 2 rho + 1.  The rotation table entries a thread needs are fixed (rho = thread index), so
// they are requested before the passes and have landed when the store needs them.
template <int CG>
__device__ __forceinline__ void inverse_store5(float4* Y, const float2* tw256, float2* s_base, const ChannelizeParams& p,
                                               int blk0, pk_t (&acc)[2][CG][2], float2 wc, int tid) {
    constexpr int BT = 2, NS = CG * BT, YS4 = NS + 1, NT = Geo5::NT;
    static_assert((size_t)(256 * YS4 + 16) * 16 <= Geo5::tile_bytes && 16 * NS <= NT, "layout");
    auto yaddr = [&](int rho, int sy) { return smem_u32(Y) + (uint32_t)(rho * YS4 + (rho >> 4) + sy) * 16; };
    // (s_base: CG x 2 block phasors, in the caller's shared memory -- the kernel has no static allocation to spare)
    // ---- radix-2 step on the accumulators, spectra -> shared ------------------------------------------------
    {
        const uint32_t dst = yaddr(tid, 0);
#pragma unroll
        for (int b = 0; b < BT; ++b)
#pragma unroll
            for (int c = 0; c < CG; ++c) {
                float r0, r1, i0, i1;
                pk_split(acc[0][c][b], r0, r1);
                pk_split(acc[1][c][b], i0, i1);
                const float dr = r0 - r1, di = i0 - i1;
                // Z1 = (Y[k'] - Y[k'+256]) * conj(W_512^{k'})
                const float z1r = fmaf(dr, wc.x, di * wc.y), z1i = fmaf(di, wc.x, -dr * wc.y);
                const pk_t re = pk_make(r0 + r1, z1r), im = pk_make(i0 + i1, z1i);
                asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(dst + (b * CG + c) * 16), "l"(re), "l"(im) : "memory");
            }
    }
    // ---- rotation operands of this thread's two output rows, requested now ------------------------------------
    const int ld = p.ld, r0 = 2 * tid - p.vd;
    float2 rot_a[CG], rot_b[CG];
#pragma unroll
    for (int c = 0; c < CG; ++c) {
        const float2* __restrict__ rt = p.rot + (size_t)c * ld;
        rot_a[c] = (r0 >= 0 && r0 < ld) ? __ldg(rt + r0) : make_float2(0.f, 0.f);
        rot_b[c] = (r0 + 1 >= 0 && r0 + 1 < ld) ? __ldg(rt + r0 + 1) : make_float2(0.f, 0.f);
    }
    if (tid < NS) {
        const int b = tid / CG, c = tid % CG;
        const int64_t mg_b = p.mg_begin + (int64_t)(blk0 + b) * ld;
        s_base[tid] = phasor_f32(nco_phase(p.phase, c, p.w[c], mg_b * (int64_t)p.decim) + p.phase_bias[c]);
    }
    __syncthreads();
    // ---- inverse pass A: 16-point transform over the high bin index a (k' = 16 a + b'), then conj(W_256^{b' alpha}) ----
    if (tid < 16 * NS) {
        const int bp = tid & 15, sy = tid >> 4;
        const uint32_t base = yaddr(bp * 16, sy);
        pk_t re[16], im[16];
        static_for<16>([&](auto ac) {
            constexpr int a = decltype(ac)::value;
            const ulonglong2 v = lds128_at<16 * a * YS4>(base);
            re[a] = v.x;
            im[a] = v.y;
        });
        pk_dif<16>(im, re);                   // inverse transform: the forward one on (im, re)
        static_for<16>([&](auto kc) {
            constexpr int al = decltype(kc)::value;
            pk_t xr = re[bitrev<16>(al)], xi = im[bitrev<16>(al)];
            if constexpr (al != 0) {
                const float2 w = tw256[al * 16 + bp];
                const pk_t wr = pk_make(w.x, w.x), wi = pk_make(w.y, w.y);
                const pk_t yr = pk_fma(xi, wi, pk_mul(xr, wr));               // x * conj(w)
                const pk_t yi = pk_sub(pk_mul(xi, wr), pk_mul(xr, wi));
                xr = yr;
                xi = yi;
            }
            sts64_at<16 * al * YS4>(base, xr);
            sts64_at<16 * al * YS4 + 8>(base, xi);
        });
    }
    __syncthreads();
    // ---- inverse pass B: 16-point transform over b' for fixed alpha: row beta*16 + alpha <- y'[alpha + 16 beta] ----
    if (tid < 16 * NS) {
        const int al = tid & 15, sy = tid >> 4;
        const uint32_t base = yaddr(al, sy);
        pk_t re[16], im[16];
        static_for<16>([&](auto bc) {
            constexpr int b = decltype(bc)::value;
            const ulonglong2 v = lds128_at<16 * (b * 16 * YS4 + b)>(base);
            re[b] = v.x;
            im[b] = v.y;
        });
        pk_dif<16>(im, re);
        static_for<16>([&](auto kc) {
            constexpr int be = decltype(kc)::value;
            sts64_at<16 * (be * 16 * YS4 + be)>(base, re[bitrev<16>(be)]);
            sts64_at<16 * (be * 16 * YS4 + be) + 8>(base, im[bitrev<16>(be)]);
        });
    }
    __syncthreads();
    // ---- drop the overlap rows, rotate by the NCO, store: thread rho writes rows 2 rho - vd and 2 rho - vd + 1 ----
    const uint32_t src = yaddr(tid, 0);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        const int blk = blk0 + b;
        if (blk >= p.nblocks) continue;
        const int64_t row0 = (int64_t)blk * ld;
        const int64_t left = p.mg_end - p.mg_begin - row0;
        const int rows = left < (int64_t)ld ? (int)left : ld;
#pragma unroll
        for (int c = 0; c < CG; ++c) {
            if (c >= p.nchan) continue;
            ulonglong2 v;
            asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(src + (b * CG + c) * 16) : "memory");
            float ar, br, ai, bi;
            pk_split(v.x, ar, br);
            pk_split(v.y, ai, bi);
            const float2 base = s_base[b * CG + c];
            float2* __restrict__ out = p.out + (size_t)c * p.out_stride + row0;
            if (r0 >= 0 && r0 < rows) out[r0] = cmul(make_float2(ar, ai), cmul(base, rot_a[c]));
            if (r0 + 1 >= 0 && r0 + 1 < rows) out[r0 + 1] = cmul(make_float2(br, bi), cmul(base, rot_b[c]));
        }
    }
}

template <int CG>
__global__ void __launch_bounds__(Geo5::NT, 2)
k_channelize5(const ChannelizeParams p, const __grid_constant__ PairMaps maps, const PairGeo geo, const int64_t tmap_row0) {
    constexpr int RS = Geo5::RS, NT = Geo5::NT, BT = 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* T = reinterpret_cast<float4*>(smem_raw);
    unsigned char* stage = smem_raw + Geo5::tile_bytes;
    unsigned char* orph = stage + 4 * kRegionBytes;
    float2* tw256 = reinterpret_cast<float2*>(orph + Geo5::orph_bytes);      // [k1][m2] = W_256^{m2 k1} = W_512^{2 m2 k1}
    float2* wctab = tw256 + 256;                              // [r] = W_512^{k'(r)}: the last radix-2 stage's twiddle of row r
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw256);       // entries 0..15 (k1 = 0) of the table are never read
    int* s_ctl = reinterpret_cast<int*>(bar + 1);             // [0]: 8 x tiles taken out of the staging buffer (+ warps of the current one), [1]: next block set
    uint64_t* bar_free = reinterpret_cast<uint64_t*>(tw256 + 14);   // "every warp has left the tile": 8 arrivals per tile

    const int tid = threadIdx.x;
    const int slot = tid & 15, rg = tid >> 4;
    const int b_slot = slot >> 3, side = (slot >> 2) & 1, col = slot & 3;
    const int m2p1 = (rg + 2 * b_slot) & 15;           // pass-1 row group of this thread (see header: banks)

    if (tid >= 16) tw256[tid] = p.twid[2 * (((tid & 15) * (tid >> 4)) & 255)];
    wctab[tid] = p.twid[(tid >> 4) + 16 * (tid & 15)];
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar_free, NT / 32);
        s_ctl[0] = 0;
    }
    __syncthreads();

    const int nsets = (p.nblocks + BT - 1) / BT;

    const int r_mac = tid;
    const int kq = (r_mac >> 4) + 16 * (r_mac & 15);          // bin within the 256-point halves
    const float4* __restrict__ gtab = reinterpret_cast<const float4*>(p.gtab) + r_mac;
    const uint32_t orow = smem_u32(orph) + r_mac * 16;        // this thread's carried spectrum (block b: + b * 4096)

    const Unpack5 unpack = make_unpack5(p);

    // staging word pointer of this thread per class (row offset df / dm folded in)
    const int tiles1 = geo.tiles1, ntiles = geo.ntiles;
    const uint32_t* const st_base = reinterpret_cast<const uint32_t*>(stage + (b_slot * 2 + side) * kRegionBytes) + col;
    const uint32_t* const st_c0 = st_base + (side ? geo.dm[0] : geo.dm[0] + 1) * 4;
    const uint32_t* const st_c1 = st_base + (side ? geo.dm[1] : geo.dm[1] + 1) * 4;

    // The copy of a tile is issued by ONE thread: the first tile of a set by thread 0 as soon as the set is known (for
    // every set but a CTA's first: before the epilogue of the set before, which does not touch the staging buffer),
    // the others by lane 0 of the last warp that has taken its rows out of the staging buffer (a counter in shared
    // memory) -- as early as the buffer is free, half a pass before the barrier that every warp waits at.
    if (tid == 0 && (int)blockIdx.x < nsets) issue_tile5(stage, bar, maps, geo, p, tmap_row0, blockIdx.x * BT, 0);

    for (int set = blockIdx.x; set < nsets;) {
        const int blk0 = set * BT;
        // acc[0][c][b] = (re_k', re_k'+256), acc[1][c][b] = (im_k', im_k'+256)
        pk_t acc[2][CG][BT];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[h][c][b] = 0ull;

        for (int t = 0; t < ntiles; ++t) {
            // every warp bumps s_ctl[0] once per tile after the wait below, so the tile's sequence number -- and with
            // it the phase of the barrier -- is s_ctl[0] / 8 for every thread that gets here (no register spent on it)
            // The tile may be overwritten when every warp is done reading it (the multiply-accumulate phase of the
            // tile before, or the epilogue): a warp says so HERE, before its own pass-1 arithmetic, and waits for the
            // others only when it is about to store -- a CTA barrier at the store would make every warp wait for the
            // slowest warp's pass 1 as well.
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_free);
            mbar_wait(bar, ((uint32_t)*reinterpret_cast<volatile int*>(s_ctl) >> 3) & 1u);
            // ------------- pass 1 (rows -> registers -> 16-point DIFs -> tile), pass 2 (in place) ----------------------
            pass1_5(t >= tiles1 ? st_c1 : st_c0, m2p1, slot, T, tw256, unpack,
                    [&] {
                        // this warp's rows are in registers: the last of the 8 warps to get here starts the next copy
                        __syncwarp();
                        if ((tid & 31) == 0 && (atomicAdd(&s_ctl[0], 1) & 7) == 7 && t + 1 < ntiles)
                            issue_tile5(stage, bar, maps, geo, p, tmap_row0, blk0, t + 1);
                    },
                    // (this warp has bumped s_ctl[0] for this tile: the tile's sequence number is (s_ctl[0] - 1) / 8)
                    [&] { mbar_wait(bar_free, ((uint32_t)(*reinterpret_cast<volatile int*>(s_ctl) - 1) >> 3) & 1u); });
            __syncthreads();
            pass2_5(T, rg, slot);
            // table entry q of this tile: per (entry, c, r) one float4 (a_k', a_k'+256, b_k', b_k'+256)
            auto gload = [&](float4 (&g)[CG], int e) {
#pragma unroll
                for (int c = 0; c < CG; ++c) g[c] = __ldcg(gtab + ((size_t)e * CG + c) * 256);
            };
#ifdef IQ2A_G3SETS
            float4 g0[CG], g1[CG], g2[CG];
#else
            float4 g0[CG], g1[CG];
#endif
            gload(g0, t * 4);
            gload(g1, t * 4 + 1);
            // Row r of the tile is read in the multiply-accumulate phase by thread r and was written in pass 2 by
            // the 16 threads with the same r >> 4, one column each: the threads of ONE half-warp.  A warp-level
            // barrier is all this hand-over needs, so the warps of a CTA drift through pass 2 and the
            // multiply-accumulate phase independently.
            __syncwarp();
            // ------------- pair butterfly + last radix-2 stage + multiply-accumulate -----------------------
            // slots of block b: forward columns b*8 + 0..3, mirror columns b*8 + 4..7; column i pairs with
            // mirror column 4 - i (i = 1, 2, 3); forward column 0 pairs with the carried spectrum
            const uint32_t trow = smem_u32(T) + r_mac * RS * 16;
            const float2 wc = wctab[tid];                            // W_512^{k'}: not kept in registers through the passes
            const bool chain_start = t == 0 || t == tiles1;
            const bool chain_end = t + 1 == tiles1 || t + 1 == ntiles;
#ifdef IQ2A_G3SETS
            gload(g2, t * 4 + 2);
            pair_mac5<CG, BT>(acc, 0, wc, lds128_at<16 * 1>(trow), lds128_at<16 * 7>(trow), g0);
            pair_mac5<CG, BT>(acc, 1, wc, lds128_at<16 * 9>(trow), lds128_at<16 * 15>(trow), g0);
            gload(g0, t * 4 + 3);
            pair_mac5<CG, BT>(acc, 0, wc, lds128_at<16 * 2>(trow), lds128_at<16 * 6>(trow), g1);
            pair_mac5<CG, BT>(acc, 1, wc, lds128_at<16 * 10>(trow), lds128_at<16 * 14>(trow), g1);
            if (chain_end) gload(g1, ntiles * 4 + (t + 1 == ntiles ? 1 : 0));
            pair_mac5<CG, BT>(acc, 0, wc, lds128_at<16 * 3>(trow), lds128_at<16 * 5>(trow), g2);
            pair_mac5<CG, BT>(acc, 1, wc, lds128_at<16 * 11>(trow), lds128_at<16 * 13>(trow), g2);
            float4 (&g_orph)[CG] = g0;
            float4 (&g_end)[CG] = g1;
#else
            // two table-entry register sets, each reloaded as soon as its step is done (one step ahead)
            pair_mac5<CG, BT>(acc, 0, wc, lds128_at<16 * 1>(trow), lds128_at<16 * 7>(trow), g0);
            pair_mac5<CG, BT>(acc, 1, wc, lds128_at<16 * 9>(trow), lds128_at<16 * 15>(trow), g0);
            gload(g0, t * 4 + 2);
            pair_mac5<CG, BT>(acc, 0, wc, lds128_at<16 * 2>(trow), lds128_at<16 * 6>(trow), g1);
            pair_mac5<CG, BT>(acc, 1, wc, lds128_at<16 * 10>(trow), lds128_at<16 * 14>(trow), g1);
            gload(g1, t * 4 + 3);
            pair_mac5<CG, BT>(acc, 0, wc, lds128_at<16 * 3>(trow), lds128_at<16 * 5>(trow), g0);
            pair_mac5<CG, BT>(acc, 1, wc, lds128_at<16 * 11>(trow), lds128_at<16 * 13>(trow), g0);
            if (chain_end) gload(g0, ntiles * 4 + (t + 1 == ntiles ? 1 : 0));
            float4 (&g_orph)[CG] = g1;
            float4 (&g_end)[CG] = g0;
#endif
            {
                const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
                const ulonglong2 c0 = chain_start ? zero : lds128_at<0>(orow);
                const ulonglong2 c1 = chain_start ? zero : lds128_at<4096>(orow);
                const ulonglong2 m0 = lds128_at<16 * 4>(trow), m1 = lds128_at<16 * 12>(trow);
                pair_mac5<CG, BT>(acc, 0, wc, lds128_at<0>(trow), c0, g_orph);
                pair_mac5<CG, BT>(acc, 1, wc, lds128_at<16 * 8>(trow), c1, g_orph);
                if (chain_end) {
                    // the column left over at the end of a class is its own mirror: plain complex product
                    single_mac5<CG, BT>(acc, 0, wc, m0, g_end);
                    single_mac5<CG, BT>(acc, 1, wc, m1, g_end);
                } else {
                    // carry the spectrum of mirror column 0 to the next tile (own rows only: no barrier needed)
                    asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(orow), "l"(m0.x), "l"(m0.y) : "memory");
                    asm volatile("st.shared.v2.b64 [%0+4096], {%1,%2};" ::"r"(orow), "l"(m1.x), "l"(m1.y) : "memory");
                }
            }
        }
        __syncthreads();                      // every warp is done reading the tile

        // next block set: a global counter when the launch provides one (CTAs run at different speeds: the L2 slice
        // an SM is close to, the other CTA on the SM), the static stride otherwise
        if (tid == 0) {
            const int next = p.set_counter ? (int)gridDim.x + atomicAdd(p.set_counter, 1) : set + (int)gridDim.x;
            s_ctl[1] = next;
            if (next < nsets) issue_tile5(stage, bar, maps, geo, p, tmap_row0, next * BT, 0);
        }
        inverse_store5<CG>(T, tw256, tw256 + 2, p, blk0, acc, wctab[tid], tid);
        __syncthreads();
        set = s_ctl[1];
    }
}

template <int CG>
static int launch_channelize5_cg(const ChannelizeParams& p, const PairMaps& maps, const PairGeo& geo, int64_t tmap_row0,
                                 int n_sm, cudaStream_t st) {
    auto kern = k_channelize5<CG>;
    static int occ = 0;
    if (!occ) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo5::smem));
        // the shared-memory budget is exact (Geo5::smem): make sure the two CTAs per SM the design counts on really fit
        IQ2A_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Geo5::NT, Geo5::smem));
        if (occ < 2) {
            set_error("k_channelize5<%d>: %d CTA per SM fits (2 expected) -- shared memory or registers grew", CG, occ);
            occ = 0;
            return IQ2A_ERR_STATE;
        }
    }
    const int nsets = (p.nblocks + 1) / 2;
    const int slots = n_sm * 2;                             // persistent: two CTAs per SM
    const int grid = nsets < slots ? nsets : slots;
    kern<<<grid, Geo5::NT, Geo5::smem, st>>>(p, maps, geo, tmap_row0);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
