// One CG instantiation of the mirror-pair channel-bank kernel per translation unit (build.py passes -DIQ2A_CG=..).
#include "channelizer5.cuh"

#ifndef IQ2A_CG
#error "compile with -DIQ2A_CG=<1..6>"
#endif
#define IQ2A_CAT2(a, b) a##b
#define IQ2A_CAT(a, b) IQ2A_CAT2(a, b)

namespace iq2a {
int IQ2A_CAT(launch_channelize5_, IQ2A_CG)(const ChannelizeParams& p, const PairMaps& maps, const PairGeo& geo,
                                           int64_t tmap_row0, int n_sm, cudaStream_t st) {
    return launch_channelize5_cg<IQ2A_CG>(p, maps, geo, tmap_row0, n_sm, st);
}
}  // namespace iq2a
