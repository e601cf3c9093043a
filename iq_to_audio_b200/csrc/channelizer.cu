// Dispatcher over the per-(M, CG) instantiations of the channel-bank kernel
// (channelizer.cuh holds the kernel, channelizer_inst.cu the instantiations).
#include <cuda.h>

#include "common.cuh"
#include "stage.cuh"
#include "pair_maps.cuh"
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

#define IQ2A_DECL2B(CG) int launch_channelize2b_##CG(const ChannelizeParams&, const CUtensorMap&, int64_t, int, cudaStream_t);
IQ2A_DECL2B(1) IQ2A_DECL2B(2) IQ2A_DECL2B(3) IQ2A_DECL2B(4) IQ2A_DECL2B(5) IQ2A_DECL2B(6)
#define IQ2A_DECL2C(CG) int launch_channelize2c_##CG(const ChannelizeParams&, int, cudaStream_t);
IQ2A_DECL2C(1) IQ2A_DECL2C(2) IQ2A_DECL2C(3) IQ2A_DECL2C(4) IQ2A_DECL2C(5) IQ2A_DECL2C(6)
#define IQ2A_DECL5(CG) int launch_channelize5_##CG(const ChannelizeParams&, const PairMaps&, const PairGeo&, int64_t, int, cudaStream_t);
IQ2A_DECL5(1) IQ2A_DECL5(2) IQ2A_DECL5(3) IQ2A_DECL5(4) IQ2A_DECL5(5) IQ2A_DECL5(6)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

bool channelize2_available() { return encode_tiled_fn() != nullptr; }

// [rows][D] int16-frame matrix starting at `base` (16-byte aligned, D % 4 == 0), box = box_cols x box_rows
static int encode_rows_map(CUtensorMap* tmap, const void* base, int D, int64_t rows, int box_cols, int box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return IQ2A_ERR_STATE; }
    const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)D * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return IQ2A_ERR_CUDA; }
    return IQ2A_OK;
}

// Generation-4 kernel over rows [p.mg_begin, p.mg_end): `base` points at the int16 frame whose global
// index is tmap_row0 * D (16-byte aligned), `rows` complete rows of D frames are readable from there.
int launch_channelize2(const ChannelizeParams& p, int cg, const void* base, int64_t tmap_row0, int64_t rows,
                       int n_sm, cudaStream_t st) {
    if (p.nblocks <= 0) return IQ2A_OK;
    CUtensorMap tmap;
    if (int rc = encode_rows_map(&tmap, base, p.decim, rows, 8, 172)) return rc;   // tiles of 8 branches x 3 boxes of 172 rows
    switch (cg) {
#define IQ2A_CASE2B(CG) case CG: return launch_channelize2b_##CG(p, tmap, tmap_row0, n_sm, st);
        IQ2A_CASE2B(1) IQ2A_CASE2B(2) IQ2A_CASE2B(3) IQ2A_CASE2B(4) IQ2A_CASE2B(5) IQ2A_CASE2B(6)
    }
    set_error("unsupported channel group %d", cg);
    return IQ2A_ERR_INVALID;
}

// Mirror-pair geometry for symmetric taps of length ntaps (channelizer5.cuh).  False when the kernel cannot take
// it: the tensor copy moves aligned groups of 4 columns, so D and r = A*D - (ntaps-1) must be multiples of 4, and
// the wrap box (Q + dm rows) must fit one tensor box of <= 256 rows.
bool pair_geometry(int ntaps, int D, PairGeo* geo) {
    if (ntaps < 2 || D < 4 || D % 4 != 0 || (ntaps - 1) % 4 != 0) return false;
    const int a = (ntaps - 1 + D - 1) / D;
    const int r = a * D - (ntaps - 1);
    PairGeo g{};
    g.a = a;
    g.g1 = r / 4;
    g.g2 = D / 4 - g.g1;
    g.tiles1 = (g.g1 + 1) / 2;
    g.ntiles = g.tiles1 + (g.g2 + 1) / 2;
    for (int k = 0; k < 2; ++k) {
        const int q = a + k;
        g.dm[k] = (8 - q % 8) % 8;
        g.hw[k] = q + g.dm[k];
        g.hl[k] = 8 * ((512 - q + 15) / 16);
        if (g.hw[k] > 256 || g.hw[k] < 8 || g.hw[k] + 2 * g.hl[k] > kRegionRows || q >= 496) return false;
    }
    *geo = g;
    return true;
}

int launch_channelize5(const ChannelizeParams& p, int cg, const PairGeo& geo, const void* base, int64_t tmap_row0,
                       int64_t rows, int n_sm, cudaStream_t st) {
    if (p.nblocks <= 0) return IQ2A_OK;
    PairMaps maps;
    int rc;
    if ((rc = encode_rows_map(&maps.fwd, base, p.decim, rows, 4, kFwdBoxRows))) return rc;
    for (int k = 0; k < 2; ++k) {
        if ((rc = encode_rows_map(&maps.wrap[k], base, p.decim, rows, 4, geo.hw[k]))) return rc;
        if ((rc = encode_rows_map(&maps.lin[k], base, p.decim, rows, 4, geo.hl[k]))) return rc;
    }
    switch (cg) {
#define IQ2A_CASE5(CG) case CG: return launch_channelize5_##CG(p, maps, geo, tmap_row0, n_sm, st);
        IQ2A_CASE5(1) IQ2A_CASE5(2) IQ2A_CASE5(3) IQ2A_CASE5(4) IQ2A_CASE5(5) IQ2A_CASE5(6)
    }
    set_error("unsupported channel group %d", cg);
    return IQ2A_ERR_INVALID;
}

int launch_channelize5_split(const ChannelizeParams& p, const PairMaps& maps, const PairGeo& geo, int64_t tmap_row0,
                             SplitParams sp, int ngroups, int cg_max, int wave_sets, int cta_slots, cudaStream_t st, int64_t* launches);

size_t channelize5_scratch_bytes_per_set(const PairGeo& geo) { return (size_t)geo.ntiles * 16 * 256 * sizeof(float4); }

int launch_channelize5_many(const ChannelizeParams& p, const PairGeo& geo, const void* base, int64_t tmap_row0, int64_t rows,
                            const SplitParams& sp, int ngroups, int cg_max, int wave_sets, int n_sm, cudaStream_t st, int64_t* launches) {
    if (p.nblocks <= 0) return IQ2A_OK;
    PairMaps maps;
    int rc;
    if ((rc = encode_rows_map(&maps.fwd, base, p.decim, rows, 4, kFwdBoxRows))) return rc;
    for (int k = 0; k < 2; ++k) {
        if ((rc = encode_rows_map(&maps.wrap[k], base, p.decim, rows, 4, geo.hw[k]))) return rc;
        if ((rc = encode_rows_map(&maps.lin[k], base, p.decim, rows, 4, geo.hl[k]))) return rc;
    }
    return launch_channelize5_split(p, maps, geo, tmap_row0, sp, ngroups, cg_max, wave_sets, 2 * n_sm, st, launches);
}

// Generation-4 kernel with cp.async staging: no tensor map, the kernel bounds-checks against p.raw_n0 / p.raw_len.
int launch_channelize2_cp(const ChannelizeParams& p, int cg, int n_sm, cudaStream_t st) {
    if (p.nblocks <= 0) return IQ2A_OK;
    switch (cg) {
#define IQ2A_CASE2C(CG) case CG: return launch_channelize2c_##CG(p, n_sm, st);
        IQ2A_CASE2C(1) IQ2A_CASE2C(2) IQ2A_CASE2C(3) IQ2A_CASE2C(4) IQ2A_CASE2C(5) IQ2A_CASE2C(6)
    }
    set_error("unsupported channel group %d", cg);
    return IQ2A_ERR_INVALID;
}

int channelize_max_group(int m_fft) { return m_fft == 1024 ? 6 : 6; }

}  // namespace iq2a
