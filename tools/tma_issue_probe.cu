// Probe: may the boxes of one mbarrier phase be issued by lane 0 of SEVERAL warps while thread 0 alone posts
// expect_tx (so that a complete_tx can reach the barrier before the expect_tx does)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue_probe tma_issue_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
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
struct Maps { CUtensorMap a, b[2]; };

// mode 0: thread 0 issues everything; 1: lane 0 of each warp issues its boxes, thread 0 posts expect_tx first in
// program order (no ordering between warps); 2: like 1 but a run-time choice between two descriptors by if/else
__global__ void __launch_bounds__(256) k_probe(const __grid_constant__ Maps maps, int mode, int iters, int sel, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    unsigned char* buf = smem + 128;
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t parity = 0, acc = 0;
    for (int it = 0; it < iters; ++it) {
        const int row0 = (blockIdx.x * 37 + it * 11) * 8;
        if (mode == 0) {
            if (tid == 0) {
                mbar_expect_tx(bar, 12 * 64 * 16);
                for (int bx = 0; bx < 12; ++bx) tma_load_2d(buf + bx * 64 * 16, &maps.a, 4 * (bx % 5), row0 + bx * 64, bar);
            }
        } else if ((tid & 31) == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tid == 0) mbar_expect_tx(bar, 12 * 64 * 16);
            for (int bx = tid >> 5; bx < 12; bx += 8) {
                // mode 2: second descriptor, aligned columns; 3: columns 4k+1 (4 bytes past a 16-byte boundary);
                // 4: a negative, aligned column (-4); 5: columns 4k+2 (8 bytes past)
                const int c0 = mode == 3 ? 4 * (bx % 5) + 1 : mode == 4 ? 4 * (bx % 5) - 4 : mode == 5 ? 4 * (bx % 5) + 2 : 4 * (bx % 5);
                if (mode == 1 || bx < 6) tma_load_2d(buf + bx * 64 * 16, &maps.a, c0, row0 + bx * 64, bar);
                else if (sel) tma_load_2d(buf + bx * 64 * 16, &maps.b[1], c0, row0 + bx * 64, bar);
                else tma_load_2d(buf + bx * 64 * 16, &maps.b[0], c0, row0 + bx * 64, bar);
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        acc += reinterpret_cast<const uint32_t*>(buf)[tid * 12];
        __syncthreads();
    }
    out[blockIdx.x * 256 + tid] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#include <cstdlib>
int main(int argc, char** argv) {
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
    const int D = 104; const int64_t rows = 400000;
    uint32_t* d; cudaMalloc(&d, (size_t)rows * D * 4); cudaMemset(d, 1, (size_t)rows * D * 4);
    uint32_t* out; cudaMalloc(&out, 296 * 256 * 4);
    Maps maps;
    CUtensorMap* all[3] = {&maps.a, &maps.b[0], &maps.b[1]};
    for (CUtensorMap* m : all) {
        const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)D * 4};
        const cuuint32_t box[2] = {4, 64};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    for (int mode = 0; mode < 6; ++mode) {
        if (only >= 0 && mode != only) continue;
        k_probe<<<296, 256, 100 * 1024>>>(maps, mode, 200, 1, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
