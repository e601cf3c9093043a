// One (M, CG) instantiation of the channel-bank kernel per translation unit, so the
// twelve variants compile in parallel (build.py passes -DIQ2A_M=.. -DIQ2A_CG=..).
#include "channelizer.cuh"

#ifndef IQ2A_M
#error "compile with -DIQ2A_M=<512|1024> -DIQ2A_CG=<1..6>"
#endif
#define IQ2A_CAT2(a, b, c) a##b##_##c
#define IQ2A_CAT(a, b, c) IQ2A_CAT2(a, b, c)

namespace iq2a {
int IQ2A_CAT(launch_channelize_, IQ2A_M, IQ2A_CG)(const ChannelizeParams& p, int codec, int n_sm, cudaStream_t st) {
    return launch_fmt<IQ2A_M, IQ2A_CG>(p, codec, n_sm, st);
}
}  // namespace iq2a
