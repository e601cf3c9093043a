// Split (many-channel) form of the mirror-pair channel bank: instantiations and the wave loop.
#include <algorithm>

#include "channelizer5s.cuh"
#include "stage.cuh"

namespace iq2a {

template <int CG>
static int launch_mac5(const ChannelizeParams& p, const PairGeo& geo, const SplitParams& sp, int ngroups, cudaStream_t st) {
    auto kern = k_mac5<CG>;
    static bool configured = false;
    if (!configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo5M::smem));
        configured = true;
    }
    kern<<<dim3(sp.nsets, ngroups), Geo5::NT, Geo5M::smem, st>>>(p, geo, sp);
    IQ2A_CUDA_TRY(cudaGetLastError());
    return IQ2A_OK;
}

// One pass over rows [p.mg_begin, p.mg_end) for every channel group of `sp.groups` (each <= cg_max channels):
// waves of `wave_sets` block sets, per wave one k_forward5 launch and one k_mac5 launch.
int launch_channelize5_split(const ChannelizeParams& p, const PairMaps& maps, const PairGeo& geo, int64_t tmap_row0,
                             SplitParams sp, int ngroups, int cg_max, int wave_sets, int cta_slots, cudaStream_t st, int64_t* launches) {
    static bool configured = false;
    if (!configured) {
        IQ2A_CUDA_TRY(cudaFuncSetAttribute(k_forward5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo5F::smem));
        configured = true;
    }
    const int nsets = (p.nblocks + 1) / 2;
    // tiles per k_forward5 CTA: as few as possible (more CTAs) while a full wave still fits the CTA slots in ONE round
    // -- 18 sets x 20 chunks of 4 tiles would be 360 CTAs for 296 slots, i.e. a second round that is 3/4 empty
    const int waves_sets = std::min(wave_sets, nsets);
    sp.tiles_per_cta = geo.ntiles;
    for (int tpc = 2; tpc <= geo.ntiles; ++tpc)
        if (waves_sets * ((geo.ntiles + tpc - 1) / tpc) <= cta_slots) { sp.tiles_per_cta = tpc; break; }
    const int chunks = (geo.ntiles + sp.tiles_per_cta - 1) / sp.tiles_per_cta;
    for (int s0 = 0; s0 < nsets; s0 += wave_sets) {
        sp.set0 = s0;
        sp.nsets = std::min(wave_sets, nsets - s0);
        k_forward5<<<sp.nsets * chunks, Geo5::NT, Geo5F::smem, st>>>(p, maps, geo, tmap_row0, sp);
        IQ2A_CUDA_TRY(cudaGetLastError());
        int rc;
        switch (cg_max) {
            case 1: rc = launch_mac5<1>(p, geo, sp, ngroups, st); break;
            case 2: rc = launch_mac5<2>(p, geo, sp, ngroups, st); break;
            case 3: rc = launch_mac5<3>(p, geo, sp, ngroups, st); break;
            case 4: rc = launch_mac5<4>(p, geo, sp, ngroups, st); break;
            case 5: rc = launch_mac5<5>(p, geo, sp, ngroups, st); break;
            default: set_error("split channel bank: unsupported group size %d", cg_max); return IQ2A_ERR_INVALID;
        }
        if (rc) return rc;
        if (launches) *launches += 2;
    }
    return IQ2A_OK;
}

}  // namespace iq2a
