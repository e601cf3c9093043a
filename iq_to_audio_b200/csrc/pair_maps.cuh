// Staging geometry and tensor maps of the mirror-pair channel bank (channelizer5.cuh), shared with the host-side
// launcher (channelizer.cu).
#pragma once
#include <cuda.h>

namespace iq2a {

constexpr int kRegionRows = 528;                       // 3 forward boxes; wrap + 2 linear boxes never need more
constexpr int kRegionBytes = kRegionRows * 16;
constexpr int kFwdBoxRows = 176;                       // 3 boxes cover rows -df .. 527-df of the window

struct PairMaps {
    CUtensorMap fwd;        // box 4 x kFwdBoxRows
    CUtensorMap wrap[2];    // box 4 x (Q + dm), per class
    CUtensorMap lin[2];     // box 4 x hl, per class
};

}  // namespace iq2a
