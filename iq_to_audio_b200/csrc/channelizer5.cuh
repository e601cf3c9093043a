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
    static constexpr size_t smem = tile_bytes + 4 * (size_t)kRegionBytes + orph_bytes + 256 * sizeof(float2) + 16;
};

template <int CG>
__global__ void __launch_bounds__(Geo5::NT, 2)
k_channelize5(const ChannelizeParams p, const __grid_constant__ PairMaps maps, const PairGeo geo, const int64_t tmap_row0) {
    constexpr int RS = Geo5::RS, NT = Geo5::NT, BT = 2;
    constexpr int NS = CG * BT, YS = NS | 1;
    static_assert((size_t)512 * YS * sizeof(float2) <= Geo5::tile_bytes, "layout");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* T = reinterpret_cast<float4*>(smem_raw);
    unsigned char* stage = smem_raw + Geo5::tile_bytes;
    unsigned char* orph = stage + 4 * kRegionBytes;
    float2* tw256 = reinterpret_cast<float2*>(orph + Geo5::orph_bytes);      // W_256^t = W_512^{2t}
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw256 + 256);

    const int tid = threadIdx.x;
    const int slot = tid & 15, rg = tid >> 4;
    const int b_slot = slot >> 3, side = (slot >> 2) & 1, col = slot & 3;
    const int m2p1 = (rg + 2 * b_slot) & 15;           // pass-1 row group of this thread (see header: banks)

    tw256[tid] = p.twid[2 * tid];
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();

    const int nsets = (p.nblocks + BT - 1) / BT;
    uint32_t parity = 0;

    const int r_mac = tid;
    const int kq = (r_mac >> 4) + 16 * (r_mac & 15);          // bin within the 256-point halves
    const float2 wc = p.twid[kq];                             // W_512^{k'}
    const float4* __restrict__ gtab = reinterpret_cast<const float4*>(p.gtab) + r_mac;
    const uint32_t orow = smem_u32(orph) + r_mac * 16;        // this thread's carried spectrum (block b: + b * 4096)

    const uint32_t qmask = p.q_neg ? 0x7fffu : 0x8000u;
    const uint32_t xmask = p.iq_swap ? (0x80000000u | qmask) : ((qmask << 16) | 0x8000u);
    const pk_t bias_i = pk_bc(-8421376.0f), bias_q = pk_bc(p.q_neg ? -8421375.0f : -8421376.0f);
    const uint32_t exp_seed = 0x4B000000u + ((uint32_t)p.iq_swap >> 8);

    // staging word pointer of this thread per class (row offset df / dm folded in)
    const int tiles1 = geo.tiles1, ntiles = geo.ntiles;
    const uint32_t* const st_base = reinterpret_cast<const uint32_t*>(stage + (b_slot * 2 + side) * kRegionBytes) + col;
    const uint32_t* const st_c0 = st_base + (side ? geo.dm[0] : geo.dm[0] + 1) * 4;
    const uint32_t* const st_c1 = st_base + (side ? geo.dm[1] : geo.dm[1] + 1) * 4;

    for (int set = blockIdx.x; set < nsets; set += gridDim.x) {
        const int blk0 = set * BT;
        // acc[0][c][b] = (re_k', re_k'+256), acc[1][c][b] = (im_k', im_k'+256)
        pk_t acc[2][CG][BT];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[h][c][b] = 0ull;

        // The copy of a tile is issued by lane 0 of every warp (12 boxes over 8 warps), so that no warp is held up
        // for long: thread 0 posts the byte count, warp w takes boxes w and w + 8 of the list
        //   block b: 3 forward boxes, the wrap box, 2 linear boxes  (b = 0: boxes 0..5, b = 1: boxes 6..11).
        auto issue = [&](int t) {
            const int cls = t >= tiles1;
            const int j = cls ? t - tiles1 : t;
            const int fwd = 4 * ((cls ? geo.g1 : 0) + j);                                  // forward group u
            const int mir = 4 * ((cls ? geo.g1 + geo.g2 : geo.g1) - 1 - j);                // mirror group w
            const int dm = cls ? geo.dm[1] : geo.dm[0], hw = cls ? geo.hw[1] : geo.hw[0], hl = cls ? geo.hl[1] : geo.hl[0];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tid == 0) mbar_expect_tx(bar, (uint32_t)(BT * (3 * kFwdBoxRows + hw + 2 * hl) * 16));
            for (int bx = tid >> 5; bx < 12; bx += 8) {
                const int b = bx >= 6, k = bx - 6 * b;
                const int64_t row0 = p.mg_begin + (int64_t)(blk0 + b) * p.ld - p.vd;
                const int rw = (int)(row0 - tmap_row0);                // tensor row of window row 0
                unsigned char* rf = stage + (b * 2) * kRegionBytes;
                unsigned char* rm = rf + kRegionBytes;
                // (one cp.async.bulk.tensor per descriptor in the source: the descriptor address stays uniform)
                if (k < 3) tma_load_2d(rf + k * kFwdBoxRows * 16, &maps.fwd, fwd, rw - (dm + 1) + k * kFwdBoxRows, bar);
                else if (k == 3) {
                    // staging rows [0, hw) <- window rows [512 - hw, 512): rows dm.. are the wrapped part of the rotation
                    if (cls) tma_load_2d(rm, &maps.wrap[1], mir, rw + 512 - hw, bar);
                    else tma_load_2d(rm, &maps.wrap[0], mir, rw + 512 - hw, bar);
                } else {
                    // staging rows [hw, hw + 2 hl) <- window rows [0, 2 hl)
                    if (cls) tma_load_2d(rm + (hw + (k - 4) * hl) * 16, &maps.lin[1], mir, rw + (k - 4) * hl, bar);
                    else tma_load_2d(rm + (hw + (k - 4) * hl) * 16, &maps.lin[0], mir, rw + (k - 4) * hl, bar);
                }
            }
        };
        if ((tid & 31) == 0) issue(0);

        // X[k'] = E + W O, X[k'+256] = E - W O of a packed (E, O) spectrum: both bins in one register pair
        auto radix2 = [&](pk_t e_o_re, pk_t e_o_im, pk_t& xr, pk_t& xi) {
            float e_r, o_r, e_i, o_i;
            pk_split(e_o_re, e_r, o_r);
            pk_split(e_o_im, e_i, o_i);
            const float pr = fmaf(wc.x, o_r, -wc.y * o_i), pi = fmaf(wc.x, o_i, wc.y * o_r);
            xr = pk_make(e_r + pr, e_r - pr);
            xi = pk_make(e_i + pi, e_i - pi);
        };
        // acc += a (Xf + Xm) + j b (Xf - Xm) for block b, table entry (a_k', a_k'+256, b_k', b_k'+256) per channel
        auto pair_mac = [&](int b, ulonglong2 vf, ulonglong2 vm, const float4 (&g)[CG]) {
            {
                pk_t sr, si;
                radix2(pk_add(vf.x, vm.x), pk_add(vf.y, vm.y), sr, si);
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    const pk_t ga = pk_make(g[c].x, g[c].y);
                    acc[0][c][b] = pk_fma(ga, sr, acc[0][c][b]);
                    acc[1][c][b] = pk_fma(ga, si, acc[1][c][b]);
                }
            }
            {
                pk_t dr, di;
                radix2(pk_sub(vf.x, vm.x), pk_sub(vf.y, vm.y), dr, di);
                const pk_t ndi = di ^ 0x8000000080000000ull;       // sign flips: off the FMA pipe
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    const pk_t gb = pk_make(g[c].z, g[c].w);
                    acc[0][c][b] = pk_fma(gb, ndi, acc[0][c][b]);
                    acc[1][c][b] = pk_fma(gb, dr, acc[1][c][b]);
                }
            }
        };

        for (int t = 0; t < ntiles; ++t) {
            mbar_wait(bar, parity);
            parity ^= 1;
            // ------------- pass 1: two 16-point DIFs (even / odd rows) per thread, packed ---------------
            {
                const uint32_t* st = t >= tiles1 ? st_c1 : st_c0;
                pk_t re[16], im[16];
                auto unpack = [&](auto swapped) {
                    constexpr uint32_t sel_lo = 0x7610u, sel_hi = 0x7632u;
                    constexpr uint32_t sel_i = decltype(swapped)::value ? sel_hi : sel_lo;
                    constexpr uint32_t sel_q = decltype(swapped)::value ? sel_lo : sel_hi;
                    uint32_t ex = exp_seed;
#pragma unroll
                    for (int m1 = 0; m1 < 16; ++m1) {
                        const int row = 32 * m1 + 2 * m2p1;
                        const uint32_t w0 = st[row * 4] ^ xmask, w1 = st[(row + 1) * 4] ^ xmask;
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
                const uint32_t dst = smem_u32(T) + (m2p1 * RS + slot) * 16;
                static_for<16>([&](auto kc) {
                    constexpr int k1 = decltype(kc)::value;
                    if constexpr (k1 != 0) {
                        // W_256^{m2 k1}: an 8-byte load that the two row groups of a warp share; ptxas turns the
                        // (w, w) pairs into scalar-broadcast operands of the packed ops
                        const pk_t xr = re[bitrev<16>(k1)], xi = im[bitrev<16>(k1)];
                        const float2 w = tw256[(m2p1 * k1) & 255];
                        const pk_t wr = pk_make(w.x, w.x), wi = pk_make(w.y, w.y);
                        re[bitrev<16>(k1)] = pk_sub(pk_mul(xr, wr), pk_mul(xi, wi));
                        im[bitrev<16>(k1)] = pk_fma(xr, wi, pk_mul(xi, wr));
                    }
                });
                __syncthreads();       // the tile is free: every warp has left the previous multiply-accumulate,
                                       // and every thread has taken its rows out of the staging buffer
                if ((tid & 31) == 0 && t + 1 < ntiles) issue(t + 1);
                static_for<16>([&](auto kc) {
                    constexpr int k1 = decltype(kc)::value;
                    sts64_at<16 * (k1 * 16) * RS>(dst, re[bitrev<16>(k1)]);
                    sts64_at<16 * (k1 * 16) * RS + 8>(dst, im[bitrev<16>(k1)]);
                });
            }
            __syncthreads();
            // ------------- pass 2: 16-point DIF over m2, in place ----------------------------------------
            {
                const int k1 = rg;
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
            // table entry q of this tile: per (entry, c, r) one float4 (a_k', a_k'+256, b_k', b_k'+256)
            auto gload = [&](float4 (&g)[CG], int e) {
#pragma unroll
                for (int c = 0; c < CG; ++c) g[c] = __ldcg(gtab + ((size_t)e * CG + c) * 256);
            };
            float4 g0[CG], g1[CG], g2[CG];
            gload(g0, t * 4);
            gload(g1, t * 4 + 1);
            __syncthreads();
            // ------------- pair butterfly + last radix-2 stage + multiply-accumulate -----------------------
            // slots of block b: forward columns b*8 + 0..3, mirror columns b*8 + 4..7; column i pairs with
            // mirror column 4 - i (i = 1, 2, 3); forward column 0 pairs with the carried spectrum
            const uint32_t trow = smem_u32(T) + r_mac * RS * 16;
            gload(g2, t * 4 + 2);
            pair_mac(0, lds128_at<16 * 1>(trow), lds128_at<16 * 7>(trow), g0);
            pair_mac(1, lds128_at<16 * 9>(trow), lds128_at<16 * 15>(trow), g0);
            gload(g0, t * 4 + 3);
            pair_mac(0, lds128_at<16 * 2>(trow), lds128_at<16 * 6>(trow), g1);
            pair_mac(1, lds128_at<16 * 10>(trow), lds128_at<16 * 14>(trow), g1);
            const bool chain_start = t == 0 || t == tiles1;
            const bool chain_end = t + 1 == tiles1 || t + 1 == ntiles;
            if (chain_end) gload(g1, ntiles * 4 + (t + 1 == ntiles ? 1 : 0));
            pair_mac(0, lds128_at<16 * 3>(trow), lds128_at<16 * 5>(trow), g2);
            pair_mac(1, lds128_at<16 * 11>(trow), lds128_at<16 * 13>(trow), g2);
            {
                const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
                const ulonglong2 c0 = chain_start ? zero : lds128_at<0>(orow);
                const ulonglong2 c1 = chain_start ? zero : lds128_at<4096>(orow);
                const ulonglong2 m0 = lds128_at<16 * 4>(trow), m1 = lds128_at<16 * 12>(trow);
                pair_mac(0, lds128_at<0>(trow), c0, g0);
                pair_mac(1, lds128_at<16 * 8>(trow), c1, g0);
                if (chain_end) {
                    // the column left over at the end of a class is its own mirror: plain complex product with
                    // the entry (e.re, e.im) packed like a pair entry
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        pk_t xr, xi;
                        radix2(b ? m1.x : m0.x, b ? m1.y : m0.y, xr, xi);
                        const pk_t nxi = xi ^ 0x8000000080000000ull;
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            const pk_t er = pk_make(g1[c].x, g1[c].y), ei = pk_make(g1[c].z, g1[c].w);
                            acc[0][c][b] = pk_fma(er, xr, acc[0][c][b]);
                            acc[0][c][b] = pk_fma(ei, nxi, acc[0][c][b]);
                            acc[1][c][b] = pk_fma(er, xi, acc[1][c][b]);
                            acc[1][c][b] = pk_fma(ei, xr, acc[1][c][b]);
                        }
                    }
                } else {
                    // carry the spectrum of mirror column 0 to the next tile (own rows only: no barrier needed)
                    asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(orow), "l"(m0.x), "l"(m0.y) : "memory");
                    asm volatile("st.shared.v2.b64 [%0+4096], {%1,%2};" ::"r"(orow), "l"(m1.x), "l"(m1.y) : "memory");
                }
            }
        }
        __syncthreads();                      // every warp is done reading the tile

        // ------------- output spectra -> shared (layout of the shared inverse), inverse, store -------------
        float2* ytile = reinterpret_cast<float2*>(T);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kbin = kq + 256 * h;
            const int yrow = (kbin & 31) * 16 + (kbin >> 5);      // slot the inverse transform expects
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    float re0, re1, im0, im1;
                    pk_split(acc[0][c][b], re0, re1);
                    pk_split(acc[1][c][b], im0, im1);
                    ytile[yrow * YS + b * CG + c] = h == 0 ? make_float2(re0, im0) : make_float2(re1, im1);
                }
        }
        inverse_and_store<512, CG, BT, NT>(ytile, p.twid, p, blk0);
        __syncthreads();
    }
}

template <int CG>
static int launch_channelize5_cg(const ChannelizeParams& p, const PairMaps& maps, const PairGeo& geo, int64_t tmap_row0,
                                 int n_sm, cudaStream_t st) {
    auto kern = k_channelize5<CG>;
    static bool configured = false;
    if (!configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo5::smem));
        configured = true;
    }
    const int nsets = (p.nblocks + 1) / 2;
    const int slots = n_sm * 2;                             // persistent: two CTAs per SM
    const int grid = nsets < slots ? nsets : slots;
    kern<<<grid, Geo5::NT, Geo5::smem, st>>>(p, maps, geo, tmap_row0);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

}  // namespace iq2a
