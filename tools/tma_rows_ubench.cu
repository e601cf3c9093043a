// TMA row-rate micro-benchmark for B200: how fast does one SM's TMA unit fill shared memory from a
// [rows][D] int32 matrix when the box is only W frames (16 / 32 / 64 bytes) wide?  The channel-bank
// kernel stages PCM tiles this way (one column strip of the decimated-by-D matrix per tile), so the
// per-row cost of the copy decides how narrow a strip a tile may use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rows_ubench tma_rows_ubench.cu && ./tma_rows_ubench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct Cfg {
    int W;            // frames per box row
    int box_rows;     // rows per box
    int boxes;        // boxes per fill
    int depth;        // fills in flight (1 or 2)
    int D;            // frames per matrix row
    int fills;        // fills per CTA
    int col_tiles;    // column strips before moving to the next row range
    int wrap_sets;    // > 0: row ranges repeat after this many sets (L2-resident run)
};

__global__ void __launch_bounds__(256) k_fill(const __grid_constant__ CUtensorMap tmap, Cfg c, int smem_pad, uint32_t* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);           // two barriers
    unsigned char* buf = smem + 128;
    const int fill_bytes = c.W * 4 * c.box_rows * c.boxes;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
    __syncthreads();
    uint32_t acc = 0;
    if (threadIdx.x == 0) {
        uint32_t par[2] = {0, 0};
        auto issue = [&](int f) {
            const int s = f % c.depth;
            int set = f / c.col_tiles; const int t = f % c.col_tiles;
            // a CTA walks its own row range like a block set of the channel bank; wrap_sets > 0: 16 row ranges
            // shared by all CTAs, so that every request hits L2 and only the copy unit itself is measured
            const int row0 = (c.wrap_sets > 0 ? (int)(blockIdx.x % 16) : (int)(blockIdx.x + set * gridDim.x)) *
                             (c.box_rows * c.boxes - 64);
            mbar_expect_tx(bar + s, fill_bytes);
            for (int b = 0; b < c.boxes; ++b)
                tma_load_2d(buf + (size_t)s * fill_bytes + (size_t)b * c.box_rows * c.W * 4, &tmap, t * c.W,
                            row0 + b * c.box_rows, bar + s);
        };
        for (int f = 0; f < c.depth && f < c.fills; ++f) issue(f);
        for (int f = 0; f < c.fills; ++f) {
            const int s = f % c.depth;
            mbar_wait(bar + s, par[s]);
            par[s] ^= 1;
            acc += *reinterpret_cast<volatile uint32_t*>(buf + (size_t)s * fill_bytes);
            if (f + c.depth < c.fills) issue(f + c.depth);
        }
    }
    if (threadIdx.x == 0) sink[blockIdx.x] = acc + smem_pad;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
    const int D = 104;
    const int64_t rows_all = 6 * 1000 * 1000;                    // 2.5 GB: larger than L2
    uint32_t* d; cudaMalloc(&d, (size_t)rows_all * D * 4); cudaMemset(d, 1, (size_t)rows_all * D * 4);
    uint32_t* sink; cudaMalloc(&sink, 4096 * 4);
    printf("%s, %d SMs, %.3f GHz\n", prop.name, nsm, clk_khz / 1e6);
    printf("%-44s %9s %9s %9s %9s\n", "variant", "us/fill", "GB/s", "clk/row", "B/clk/SM");
    struct V { const char* name; int W, box_rows, boxes, depth, ctas_per_sm; };
    const V vs[] = {
        {"32 B rows, 1032 rows/fill, depth 1, 2 CTA/SM", 8, 172, 6, 1, 2},
        {"32 B rows, 1032 rows/fill, depth 2, 2 CTA/SM", 8, 172, 6, 2, 2},
        {"16 B rows, 2112 rows/fill, depth 1, 2 CTA/SM", 4, 176, 12, 1, 2},
        {"16 B rows, 2112 rows/fill, depth 2, 2 CTA/SM", 4, 176, 12, 2, 2},
        {"64 B rows,  516 rows/fill, depth 1, 2 CTA/SM", 16, 172, 3, 1, 2},
        {"64 B rows,  516 rows/fill, depth 2, 2 CTA/SM", 16, 172, 3, 2, 2},
        {"32 B rows, 2064 rows/fill, depth 1, 1 CTA/SM", 8, 172, 12, 1, 1},
        {"32 B rows, 2064 rows/fill, depth 2, 1 CTA/SM", 8, 172, 12, 2, 1},
        {"32 B rows, 1032 rows/fill, depth 1, 1 CTA/SM", 8, 172, 6, 1, 1},
        {"32 B rows,  256-row boxes, depth 2, 2 CTA/SM", 8, 256, 4, 2, 2},
        {"128 B rows, 256 rows/fill, depth 2, 2 CTA/SM", 32, 128, 2, 2, 2},
    };
    for (int pass = 0; pass < 2; ++pass) {
    const int64_t rows = pass == 0 ? rows_all : 200 * 1000;      // 2.5 GB (DRAM) / 83 MB (L2-resident)
    printf("---- matrix of %lld rows (%.0f MB)\n", (long long)rows, rows * D * 4 / 1e6);
    for (const V& v : vs) {
        CUtensorMap tmap;
        const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)D * 4};
        const cuuint32_t box[2] = {(cuuint32_t)v.W, (cuuint32_t)v.box_rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", v.name, (int)r); continue; }
        Cfg c{v.W, v.box_rows, v.boxes, v.depth, D, 0, D / v.W, 0};
        const int fill_bytes = v.W * 4 * v.box_rows * v.boxes;
        const int grid = nsm * v.ctas_per_sm;
        const int rows_per_set = v.box_rows * v.boxes - 64;
        int sets = (int)((rows - 4096) / rows_per_set / grid);
        if (sets < 1) sets = 1;
        c.fills = sets * c.col_tiles;
        if (c.fills > 13 * 24) c.fills = 13 * 24;
        if (pass == 1) c.fills = 13 * 24, c.wrap_sets = sets;
        // shared memory sized so that exactly ctas_per_sm CTAs fit
        size_t smem = 128 + (size_t)fill_bytes * v.depth;
        const size_t want = v.ctas_per_sm == 2 ? 100 * 1024 : 200 * 1024;
        if (smem < want) smem = want;
        cudaFuncSetAttribute(k_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_fill<<<grid, 256, smem>>>(tmap, c, 0, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_fill<<<grid, 256, smem>>>(tmap, c, 0, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)grid * c.fills * fill_bytes;
        const double nrows = (double)grid * c.fills * v.box_rows * v.boxes;
        const double clk = ms * 1e-3 * clk_khz * 1e3;
        printf("%-44s %9.3f %9.1f %9.3f %9.2f\n", v.name, ms * 1e3 / c.fills, bytes / (ms * 1e-3) / 1e9,
               clk / (nrows / nsm), bytes / nsm / clk);
    }
    }
    return 0;
}
