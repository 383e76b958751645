// p2p_combine.cu -- inter-GPU split-KV combine over NVLink peer memory (north-star long-context
// mode).  The reference has no multi-device code; the exchange step is new.
//
// Each rank holds a contiguous page range of ONE sequence and produces un-normalised partials
// (m, l, O) per (row, head) with pa_paged_decode_f16_partial.  The message is tiny (D+2 floats per
// row-head, 16.6 KB per rank at the Llama-7B shape), so the exchange is latency-bound: instead of
// an NCCL all-gather followed by a combine kernel, ONE kernel
//   1. stores this rank's partial row straight into slot `rank` of every peer's exchange buffer
//      (plain st.global to the peer's NVLink-mapped address),
//   2. publishes it with a release.sys flag per (src rank, row),
//   3. acquires the flags of all source ranks for its row, and
//   4. LSE-combines the `world` partials locally.
// Rows are independent, so there is no grid-wide barrier: CTA `row` on every rank only waits for
// CTA `row` of the other ranks.  Buffers are double-buffered by epoch parity (a rank can be at
// most one step ahead of its slowest peer because step e+1's wait needs every peer's e+1 data,
// which a peer only sends after its step-e kernel finished).
#include <cstring>

#include "pa_common.cuh"

namespace pa {

// Exchange buffer layout (per rank, identical on all ranks):
//   data  : [2 parity][world src][rows][D + 2] float   (O[0..D), m, l)
//   flags : [2 parity][world src][rows] uint32 (epoch of the data in the slot)
__host__ __device__ inline size_t xbuf_data_floats(int world, int rows, int D) {
    return (size_t)2 * world * rows * (D + 2);
}
__host__ __device__ inline size_t xbuf_bytes(int world, int rows, int D) {
    return xbuf_data_floats(world, rows, D) * sizeof(float) + (size_t)2 * world * rows * sizeof(uint32_t);
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_volatile_f32(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void splitkv_exchange_combine_kernel(const float* __restrict__ pm, const float* __restrict__ pl,
                                                const float* __restrict__ po, uint8_t* const* __restrict__ peers,
                                                int rank, int world, int rows, int D, uint32_t* __restrict__ epochs,
                                                long long timeout_cycles, float* __restrict__ out,
                                                float* __restrict__ lse_out, int* __restrict__ status) {
    extern __shared__ float sm[];  // [world] m, [world] l
    const int row = blockIdx.x;
    // Per-row epoch counter in device memory (so the launch is CUDA-graph replayable): CTA `row` is
    // the only reader/writer of epochs[row] on this rank; all ranks advance in lock step.
    const uint32_t epoch = epochs[row] + 1u;
    const int par = epoch & 1u;
    const size_t row_stride = D + 2;
    const size_t slot = ((size_t)(par * world + rank) * rows + row) * row_stride;
    const size_t flag_idx = (size_t)(par * world + rank) * rows + row;
    const size_t data_bytes = xbuf_data_floats(world, rows, D) * sizeof(float);

    // 1. scatter my partial row to every rank (self included)
    const float m_mine = pm[row], l_mine = pl[row];
    for (int p = 0; p < world; ++p) {
        float* dst = reinterpret_cast<float*>(peers[p]) + slot;
        for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = po[(size_t)row * D + d];
        if (threadIdx.x == 0) {
            dst[D] = m_mine;
            dst[D + 1] = l_mine;
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish
    if (threadIdx.x < world) {
        uint32_t* f = reinterpret_cast<uint32_t*>(peers[threadIdx.x] + data_bytes) + flag_idx;
        st_release_sys(f, epoch);
    }
    // 3. wait for every source rank's row
    uint8_t* mine = peers[rank];
    bool ok = true;
    if (threadIdx.x < world) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(mine + data_bytes) +
                            (size_t)(par * world + threadIdx.x) * rows + row;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != epoch) {
            if (clock64() - t0 > timeout_cycles) {
                ok = false;
                break;
            }
        }
        if (!ok && status) atomicExch(status, 1);
        const float* src = reinterpret_cast<const float*>(mine) + ((size_t)(par * world + threadIdx.x) * rows + row) * row_stride;
        sm[threadIdx.x] = ok ? ld_volatile_f32(src + D) : -INFINITY;
        sm[world + threadIdx.x] = ok ? ld_volatile_f32(src + D + 1) : 0.f;
    }
    __syncthreads();
    // 4. combine (same math as lse_combine_kernel / oracle orc_lse_combine)
    float M = -INFINITY;
    for (int s = 0; s < world; ++s) M = fmaxf(M, sm[s]);
    float L = 0.f;
    for (int s = 0; s < world; ++s) L += (sm[s] == -INFINITY) ? 0.f : sm[world + s] * __expf(sm[s] - M);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float O = 0.f;
        for (int s = 0; s < world; ++s) {
            const float w = (sm[s] == -INFINITY) ? 0.f : __expf(sm[s] - M);
            const float* src = reinterpret_cast<const float*>(mine) + ((size_t)(par * world + s) * rows + row) * row_stride;
            O = fmaf(ld_volatile_f32(src + d), w, O);
        }
        out[(size_t)row * D + d] = O / (L + 1e-6f);
    }
    if (threadIdx.x == 0 && lse_out) lse_out[row] = (L > 0.f) ? M + logf(L) : -INFINITY;
    if (threadIdx.x == 0) epochs[row] = epoch;
}

}  // namespace pa

using namespace pa;

PA_API size_t pa_splitkv_exchange_bytes(int world, int rows, int head_dim) {
    if (world <= 0 || rows <= 0 || head_dim <= 0) return 0;
    return xbuf_bytes(world, rows, head_dim);
}

// cudaMalloc'd (IPC-exportable), zero-filled buffer + its 64-byte IPC handle.
PA_API int pa_p2p_alloc(size_t bytes, void** d_ptr, unsigned char* handle64) {
    PA_CHECK_ARG(bytes > 0 && d_ptr && handle64);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return PA_OK;
}

PA_API int pa_p2p_open(const unsigned char* handle64, void** d_peer_ptr) {
    PA_CHECK_ARG(handle64 && d_peer_ptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? PA_OK : (int)e;
}

PA_API int pa_p2p_close(void* d_peer_ptr) {
    PA_CHECK_ARG(d_peer_ptr);
    cudaError_t e = cudaIpcCloseMemHandle(d_peer_ptr);
    return e == cudaSuccess ? PA_OK : (int)e;
}

PA_API int pa_p2p_free(void* d_ptr) {
    PA_CHECK_ARG(d_ptr);
    cudaError_t e = cudaFree(d_ptr);
    return e == cudaSuccess ? PA_OK : (int)e;
}

PA_API int pa_splitkv_exchange_combine(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                                       void* const* d_peer_bufs, int rank, int world, int rows, int head_dim,
                                       uint32_t* d_epochs, float* d_out, float* d_lse_out, int* d_status,
                                       pa_stream_t stream) {
    PA_CHECK_ARG(d_part_m && d_part_l && d_part_o && d_peer_bufs && d_out && d_epochs);
    PA_CHECK_ARG(world > 0 && world <= 64 && rank >= 0 && rank < world && rows >= 0 && head_dim > 0);
    if (rows == 0) return PA_OK;
    int threads = head_dim < 64 ? 64 : (head_dim > 256 ? 256 : head_dim);
    if (threads < world) threads = 64;
    const long long timeout_cycles = 4000000000ll;  // ~2 s: a missing peer must not hang the GPU
    splitkv_exchange_combine_kernel<<<rows, threads, 2 * world * sizeof(float), as_stream(stream)>>>(
        d_part_m, d_part_l, d_part_o, reinterpret_cast<uint8_t* const*>(d_peer_bufs), rank, world, rows, head_dim,
        d_epochs, timeout_cycles, d_out, d_lse_out, d_status);
    PA_RETURN_LAUNCH_STATUS();
}
