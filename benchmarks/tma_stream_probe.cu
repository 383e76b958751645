// Experiment (not part of the library): how fast can 128 CTAs stream a [K, N] row-major int8 weight
// matrix through TMA when each CTA owns a 128-byte wide column slab (the GEMM's access pattern: every
// box row is a separate 128-byte segment, 16 KB apart) versus the same bytes pre-tiled so that each
// 128 x 128 tile is one contiguous 16 KB run?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I <pkg>/csrc \
//        benchmarks/tma_stream_probe.cu -o build/tma_stream_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "pa_common.cuh"
using namespace pa;

constexpr int ST = 8, TILE = 16384;

__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// mode 0: tensor-map boxes [128 rows][128 B] from the row-major matrix; mode 1: contiguous 16 KB bulk copies
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* tiled,
                                                       int K, int N, int mode, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bar0 = base + ST * TILE;
    const int nkb = K / 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * ST; ++s) mbar_init(bar0 + s * 8, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int n0 = blockIdx.x * 128;
    if (threadIdx.x == 0) {  // producer
        for (int i = 0; i < nkb; ++i) {
            const int s = i % ST;
            mbar_wait(bar0 + (ST + s) * 8, ((i / ST) & 1) ^ 1);
            mbar_arrive_expect_tx(bar0 + s * 8, TILE);
            if (mode == 0) tma2d(base + s * TILE, &tm, n0, i * 128, bar0 + s * 8);
            else bulk_g2s_nohint(base + s * TILE, tiled + ((size_t)i * (N / 128) + blockIdx.x) * TILE, TILE, bar0 + s * 8);
        }
    } else if (threadIdx.x == 32) {  // consumer: just frees the slot
        unsigned long long acc = 0;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % ST;
            mbar_wait(bar0 + s * 8, (i / ST) & 1);
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + s * TILE));
            acc += v;
            mbar_arrive(bar0 + (ST + s) * 8);
        }
        if (acc == 0x1234567) *sink = acc;
    }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 4096, N = argc > 2 ? atoi(argv[2]) : 16384;
    const int promo = argc > 3 ? atoi(argv[3]) : 2;  // 0 none, 1 64B, 2 128B, 3 256B
    uint8_t* W[4];
    for (int i = 0; i < 4; ++i) { cudaMalloc(&W[i], (size_t)K * N); cudaMemset(W[i], i + 1, (size_t)K * N); }
    unsigned long long* sink; cudaMalloc(&sink, 8);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    CUtensorMap tm[4];
    for (int i = 0; i < 4; ++i) {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)K}; cuuint64_t strides[1] = {(cuuint64_t)N};
        cuuint32_t box[2] = {128, 128}; cuuint32_t es[2] = {1, 1};
        enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, W[i], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    const size_t smem = ST * TILE + 2 * ST * 8 + 1024;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int it = 0; it < 8; ++it) stream_kernel<<<N / 128, 64, smem>>>(tm[it & 3], W[it & 3], K, N, mode, sink);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        const int iters = 40;
        for (int it = 0; it < iters; ++it) stream_kernel<<<N / 128, 64, smem>>>(tm[it & 3], W[it & 3], K, N, mode, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("K=%d N=%d promo=%d mode=%s: %.2f us per pass, %.0f GB/s (err %s)\n", K, N, promo,
               mode == 0 ? "row-major slabs (tensor map)" : "pre-tiled 16 KB runs (bulk)", ms / iters * 1e3,
               (double)K * N / (ms / iters * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
