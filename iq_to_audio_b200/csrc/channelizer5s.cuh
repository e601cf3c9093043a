// Many-channel form of the mirror-pair channel bank: the forward transforms are computed ONCE per block set and
// shared by every channel group.
//
// k_channelize5 takes at most 6 channels per launch, so a bank of C channels repeats the forward branch transforms
// ceil(C / 6) times -- at 256 channels (BASELINE configs[4]) they were 3/4 of all work.  Here the two halves of the
// fused kernel run as two kernels over a "wave" of block sets small enough for its spectra to stay in L2:
//   k_forward5   (set, chunk of tiles):  TMA staging, pass 1, pass 2 as in k_channelize5, then the tile's spectra
//                go to scratch[set][tile][slot][row] (float4 (E.re, O.re, E.im, O.im), 64 KB per tile)
//   k_mac5<CG>   (set, channel group):   the multiply-accumulate phase of k_channelize5 with the spectra read from
//                scratch (coalesced: thread r reads row r), then the same epilogue (inverse_store5)
// The spectrum a tile hands to the next one (channelizer5.cuh: `orph`) is simply read from the previous tile's
// scratch.  Table entries of a group stay in L2 across the sets of a wave because the group index is the slow grid
// dimension.  Same arithmetic in the same order as the fused kernel: identical channel samples.
#pragma once
#include "channelizer5.cuh"

namespace iq2a {

struct Geo5F {
    static constexpr size_t smem = Geo5::tile_bytes + 4 * (size_t)kRegionBytes + 256 * sizeof(float2) + 32;
};

__global__ void __launch_bounds__(Geo5::NT, 2)
k_forward5(const ChannelizeParams p, const __grid_constant__ PairMaps maps, const PairGeo geo, const int64_t tmap_row0,
           const SplitParams sp) {
    constexpr int RS = Geo5::RS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* T = reinterpret_cast<float4*>(smem_raw);
    unsigned char* stage = smem_raw + Geo5::tile_bytes;
    float2* tw256 = reinterpret_cast<float2*>(stage + 4 * kRegionBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw256 + 256);
    int* s_ctl = reinterpret_cast<int*>(bar + 1);

    const int tid = threadIdx.x;
    const int slot = tid & 15, rg = tid >> 4;
    const int b_slot = slot >> 3, side = (slot >> 2) & 1, col = slot & 3;
    const int m2p1 = (rg + 2 * b_slot) & 15;
    tw256[tid] = p.twid[2 * (((tid & 15) * (tid >> 4)) & 255)];
    if (tid == 0) {
        mbar_init(bar, 1);
        s_ctl[0] = 0;
    }
    __syncthreads();
    const Unpack5 unpack = make_unpack5(p);
    const int tiles1 = geo.tiles1, ntiles = geo.ntiles;
    const uint32_t* const st_base = reinterpret_cast<const uint32_t*>(stage + (b_slot * 2 + side) * kRegionBytes) + col;
    const uint32_t* const st_c0 = st_base + (side ? geo.dm[0] : geo.dm[0] + 1) * 4;
    const uint32_t* const st_c1 = st_base + (side ? geo.dm[1] : geo.dm[1] + 1) * 4;

    const int chunks = (ntiles + sp.tiles_per_cta - 1) / sp.tiles_per_cta;
    const int set_l = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const int blk0 = (sp.set0 + set_l) * 2;
    const int t0 = chunk * sp.tiles_per_cta, t1 = min(ntiles, t0 + sp.tiles_per_cta);
    float4* __restrict__ dst = sp.scratch + ((size_t)set_l * ntiles) * 16 * 256 + tid;

    if (tid == 0) issue_tile5(stage, bar, maps, geo, p, tmap_row0, blk0, t0);
    for (int t = t0; t < t1; ++t) {
        mbar_wait(bar, ((uint32_t)*reinterpret_cast<volatile int*>(s_ctl) >> 3) & 1u);
        pass1_5(t >= tiles1 ? st_c1 : st_c0, m2p1, slot, T, tw256, unpack,
                [&] {
                    __syncwarp();
                    if ((tid & 31) == 0 && (atomicAdd(&s_ctl[0], 1) & 7) == 7 && t + 1 < t1)
                        issue_tile5(stage, bar, maps, geo, p, tmap_row0, blk0, t + 1);
                },
                [&] { __syncthreads(); });       // the tile is free: every warp has copied the previous tile out
        __syncthreads();
        pass2_5(T, rg, slot);
        __syncwarp();                             // row r was written by the threads of r's own half-warp
        // spectra of the tile -> scratch[set][t][slot][row]: thread r copies row r of every slot (coalesced stores)
        const uint32_t trow = smem_u32(T) + tid * RS * 16;
        float4* __restrict__ d = dst + (size_t)t * 16 * 256;
        static_for<16>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const ulonglong2 v = lds128_at<16 * s>(trow);
            float4 f;
            pk_split(v.x, f.x, f.y);
            pk_split(v.y, f.z, f.w);
            __stcg(d + s * 256, f);
        });
    }
}

template <int CG>
__global__ void __launch_bounds__(Geo5::NT, 2)
k_mac5(const ChannelizeParams base, const PairGeo geo, const SplitParams sp) {
    constexpr int BT = 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* T = reinterpret_cast<float4*>(smem_raw);                                  // epilogue workspace
    float2* tw256 = reinterpret_cast<float2*>(smem_raw + Geo5::tile_bytes);
    const int tid = threadIdx.x;
    if (tid >= 16) tw256[tid] = base.twid[2 * (((tid & 15) * (tid >> 4)) & 255)];     // entries 0..15: the epilogue's phasors

    const SplitGroup grp = sp.groups[blockIdx.y];
    const int set_l = blockIdx.x;
    const int blk0 = (sp.set0 + set_l) * 2;
    const int tiles1 = geo.tiles1, ntiles = geo.ntiles;
    const int cnt = grp.count;                                                       // <= CG
    const int kq = (tid >> 4) + 16 * (tid & 15);
    const float2 wc = base.twid[kq];
    const float4* __restrict__ gtab = sp.gtab5 + grp.g5_off + tid;                   // [entry][cnt][256]
    const float4* __restrict__ X = sp.scratch + ((size_t)set_l * ntiles) * 16 * 256 + tid;   // + (t*16 + slot)*256

    pk_t acc[2][CG][BT];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < CG; ++c)
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[h][c][b] = 0ull;

    auto gload = [&](float4 (&g)[CG], int e) {
#pragma unroll
        for (int c = 0; c < CG; ++c)
            g[c] = c < cnt ? __ldcg(gtab + ((size_t)e * cnt + c) * 256) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto xload = [&](int t, int s) {
        const float4 f = __ldcg(X + ((size_t)t * 16 + s) * 256);
        return make_ulonglong2(pk_make(f.x, f.y), pk_make(f.z, f.w));
    };
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);

    // one step = one table entry applied to both blocks; operands of step n+1 are requested before step n computes
    float4 g_a[CG], g_b[CG];
    ulonglong2 xa[4], xb[4];                              // (vf, vm) of block 0, (vf, vm) of block 1
    auto fetch = [&](float4 (&g)[CG], ulonglong2 (&x)[4], int t, int q) {
        gload(g, t * 4 + q);
        const bool chain_start = t == 0 || t == tiles1;
        if (q < 3) {
            x[0] = xload(t, q + 1);
            x[1] = xload(t, 7 - q);
            x[2] = xload(t, 9 + q);
            x[3] = xload(t, 15 - q);
        } else {
            x[0] = xload(t, 0);
            x[1] = chain_start ? zero : xload(t - 1, 4);
            x[2] = xload(t, 8);
            x[3] = chain_start ? zero : xload(t - 1, 12);
        }
    };
    auto apply = [&](const float4 (&g)[CG], const ulonglong2 (&x)[4]) {
        pair_mac5<CG, BT>(acc, 0, wc, x[0], x[1], g);
        pair_mac5<CG, BT>(acc, 1, wc, x[2], x[3], g);
    };
    fetch(g_a, xa, 0, 0);
    for (int t = 0; t < ntiles; ++t) {
        fetch(g_b, xb, t, 1);
        apply(g_a, xa);
        fetch(g_a, xa, t, 2);
        apply(g_b, xb);
        fetch(g_b, xb, t, 3);
        apply(g_a, xa);
        const bool chain_end = t + 1 == tiles1 || t + 1 == ntiles;
        if (chain_end) {
            // the column left over at the end of a class is its own mirror: plain complex product
            gload(g_a, ntiles * 4 + (t + 1 == ntiles ? 1 : 0));
            xa[0] = xload(t, 4);
            xa[1] = xload(t, 12);
            apply(g_b, xb);
            single_mac5<CG, BT>(acc, 0, wc, xa[0], g_a);
            single_mac5<CG, BT>(acc, 1, wc, xa[1], g_a);
            if (t + 1 < ntiles) fetch(g_a, xa, t + 1, 0);
        } else {
            if (t + 1 < ntiles) fetch(g_a, xa, t + 1, 0);
            apply(g_b, xb);
        }
    }

    // the epilogue of the fused kernel on this group's slice of the bank
    ChannelizeParams p = base;
    p.nchan = cnt;
    p.rot = sp.rot + (size_t)grp.first * base.ld;
    p.out = sp.out + (size_t)grp.first * base.out_stride;
    p.phase.tab = sp.phase_tab + (size_t)grp.first * base.phase.nseg;
#pragma unroll
    for (int c = 0; c < CG; ++c) {
        p.w[c] = c < cnt ? sp.w[grp.first + c] : 0.0;
        p.phase_bias[c] = c < cnt ? sp.phase_bias[grp.first + c] : 0.0;
    }
    __syncthreads();                                       // tw256 is in place
    inverse_store5<CG>(T, tw256, tw256 + 2, p, blk0, acc, wc, tid);
}

struct Geo5M {
    static constexpr size_t smem = Geo5::tile_bytes + 256 * sizeof(float2);
};

}  // namespace iq2a
