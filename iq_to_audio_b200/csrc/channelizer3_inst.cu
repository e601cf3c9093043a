// One CG instantiation of the third-generation channel-bank kernel per translation unit.
#include "channelizer3.cuh"

#ifndef IQ2A_CG
#error "compile with -DIQ2A_CG=<1..6>"
#endif
#define IQ2A_CAT2(a, b) a##b
#define IQ2A_CAT(a, b) IQ2A_CAT2(a, b)

namespace iq2a {
int IQ2A_CAT(launch_channelize3_, IQ2A_CG)(const ChannelizeParams& p, const CUtensorMap& tmap, int64_t tmap_row0,
                                           int n_sm, cudaStream_t st) {
    return launch_channelize3_cg<IQ2A_CG>(p, tmap, tmap_row0, n_sm, st);
}
}  // namespace iq2a
