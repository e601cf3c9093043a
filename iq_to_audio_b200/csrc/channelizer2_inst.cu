// One CG instantiation of the second-generation channel-bank kernel per translation unit
// (build.py passes -DIQ2A_CG=..).
#include "channelizer2.cuh"

#ifndef IQ2A_CG
#error "compile with -DIQ2A_CG=<1..6>"
#endif
#define IQ2A_CAT2(a, b) a##b
#define IQ2A_CAT(a, b) IQ2A_CAT2(a, b)

namespace iq2a {
int IQ2A_CAT(launch_channelize2b_, IQ2A_CG)(const ChannelizeParams& p, const CUtensorMap& tmap, int64_t tmap_row0,
                                            int n_sm, cudaStream_t st) {
    return launch_channelize2_cg<IQ2A_CG, 2>(p, tmap, tmap_row0, n_sm, st);
}
// same kernel with cp.async staging (no tensor map: any D, any 4-byte aligned buffer, partial rows)
int IQ2A_CAT(launch_channelize2c_, IQ2A_CG)(const ChannelizeParams& p, int n_sm, cudaStream_t st) {
    CUtensorMap unused{};
    return launch_channelize2_cg<IQ2A_CG, 2, 1>(p, unused, 0, n_sm, st);
}
}  // namespace iq2a
