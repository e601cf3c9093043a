// Dispatcher over the per-(M, CG) instantiations of the channel-bank kernel
// (channelizer.cuh holds the kernel, channelizer_inst.cu the instantiations).
#include "common.cuh"
#include "stage.cuh"
#include "../../include/iq2a_b200.h"

namespace iq2a {

#define IQ2A_DECL(M, CG) int launch_channelize_##M##_##CG(const ChannelizeParams&, int, int, cudaStream_t);
#define IQ2A_FOR_CG(M) IQ2A_DECL(M, 1) IQ2A_DECL(M, 2) IQ2A_DECL(M, 3) IQ2A_DECL(M, 4) IQ2A_DECL(M, 5) IQ2A_DECL(M, 6)
IQ2A_FOR_CG(512)
IQ2A_FOR_CG(1024)

// cg: channels per launch compiled in (1..6); p.gtab must be laid out [D][cg][M].
int launch_channelize(const ChannelizeParams& p, int m_fft, int cg, int codec, int n_sm, cudaStream_t st) {
    if (p.nblocks <= 0) return IQ2A_OK;
#define IQ2A_CASE(M, CG) if (m_fft == M && cg == CG) return launch_channelize_##M##_##CG(p, codec, n_sm, st);
#define IQ2A_CASES(M) IQ2A_CASE(M, 1) IQ2A_CASE(M, 2) IQ2A_CASE(M, 3) IQ2A_CASE(M, 4) IQ2A_CASE(M, 5) IQ2A_CASE(M, 6)
    IQ2A_CASES(512)
    IQ2A_CASES(1024)
    set_error("unsupported transform size %d / channel group %d", m_fft, cg);
    return IQ2A_ERR_INVALID;
}

int channelize_max_group(int m_fft) { return m_fft == 1024 ? 6 : 6; }

}  // namespace iq2a
