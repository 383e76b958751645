// p2p_combine.cu -- inter-GPU split-KV combine (north-star long-context mode).  The reference has
// no multi-device code; the exchange step is new.
//
// Each rank holds a contiguous page range of ONE sequence and produces un-normalised partials (m, l, O) per
// (row, head).  The message is tiny (D+2 floats per row-head, 16.6 KB per rank at the Llama-7B shape), so the
// exchange is latency-bound.  Three forms, all with the same result:
//   * pa_paged_decode_{f16,i8}_splitkv (paged_decode.cu): decode, row merge, send and receive+combine in ONE
//     launch -- the warp that finishes a row's last chunk stores it into every peer's buffer as flag-in-data
//     packets (xchg.cuh), rows are received at the end of the same kernel;
//   * pa_splitkv_exchange_combine (here): the same packet exchange as a stand-alone kernel after
//     pa_paged_decode_*_partial -- phase 1 sends all rows (never blocks), phase 2 receives and combines;
//   * pa_nccl_allgather_combine (here): the north star's baseline -- ncclAllGather of (m, l, O) over NVLink /
//     NVSwitch followed by pa_lse_combine's kernel.  NCCL is loaded at run time (dlopen "libnccl.so.2": inside
//     a PyTorch process that is the library torch already loaded), so libpa_b200.so has no link-time dependency.
#include <dlfcn.h>

#include <cstring>

#include "pa_common.cuh"
#include "xchg.cuh"

namespace pa {

// grid-stride over rows, one warp per row: all sends first, then all receives (a send never waits, so no
// ordering between ranks or CTAs is assumed and any number of rows fits any grid).
// phases: 1 = send, 2 = receive + combine, 3 = both (the normal call)
template <int D>
__global__ void __launch_bounds__(256) splitkv_exchange_combine_kernel(
    const float* __restrict__ pm, const float* __restrict__ pl, const float* __restrict__ po,
    uint8_t* const* __restrict__ peers, int rank, int world, int rows, uint32_t* __restrict__ epochs,
    float* __restrict__ out, float* __restrict__ lse_out, int* __restrict__ status, int phases) {
    constexpr int VEC = D / 32;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    if (phases & 1) {
        for (int row = gw; row < rows; row += nw) {
            float O[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) O[e] = po[(size_t)row * D + lane * VEC + e];
            xchg::send_row<D>(peers, epochs, rank, world, rows, row, O, pm[row], pl[row], lane);
        }
    }
    if (phases & 2) {
        for (int row = gw; row < rows; row += nw)
            xchg::recv_row<D>(peers, epochs, rank, world, rows, row, out, lse_out, status, lane);
    }
}

// ---- NCCL through dlopen -------------------------------------------------------------------------------
struct NcclId128 {  // ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
    char b[128];
};
struct NcclApi {
    // ncclResult_t is an int enum (0 = success); ncclFloat = 7.
    int (*GetUniqueId)(void*);
    int (*CommInitRank)(void**, int, NcclId128, int);
    int (*CommDestroy)(void*);
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    bool ok = false;
};
static NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
            api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
            api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GroupStart &&
                     api.GroupEnd;
        }
    }
    return api;
}

}  // namespace pa

using namespace pa;

PA_API size_t pa_splitkv_exchange_bytes(int world, int rows, int head_dim) {
    if (world <= 0 || rows <= 0 || head_dim <= 0) return 0;
    return xchg::buffer_bytes(world, rows, head_dim);
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

static int exchange_launch(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                           void* const* d_peer_bufs, int rank, int world, int rows, int head_dim,
                           uint32_t* d_epochs, float* d_out, float* d_lse_out, int* d_status,
                           pa_stream_t stream, int phases) {
    PA_CHECK_ARG(d_peer_bufs && d_epochs);
    PA_CHECK_ARG(!(phases & 1) || (d_part_m && d_part_l && d_part_o));
    PA_CHECK_ARG(!(phases & 2) || d_out);
    PA_CHECK_ARG(world > 0 && world <= 32 && rank >= 0 && rank < world && rows >= 0);
    if (head_dim != 64 && head_dim != 128) return PA_ERR_UNSUPPORTED;
    if (rows == 0) return PA_OK;
    const DeviceInfo& di = device_info();
    int blocks = (rows + 7) / 8;
    if (di.ok && blocks > di.sm_count * 8) blocks = di.sm_count * 8;  // every CTA resident; rows are grid-strided
    uint8_t* const* peers = reinterpret_cast<uint8_t* const*>(d_peer_bufs);
    if (head_dim == 128)
        splitkv_exchange_combine_kernel<128><<<blocks, 256, 0, as_stream(stream)>>>(
            d_part_m, d_part_l, d_part_o, peers, rank, world, rows, d_epochs, d_out, d_lse_out, d_status, phases);
    else
        splitkv_exchange_combine_kernel<64><<<blocks, 256, 0, as_stream(stream)>>>(
            d_part_m, d_part_l, d_part_o, peers, rank, world, rows, d_epochs, d_out, d_lse_out, d_status, phases);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_splitkv_exchange_combine(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                                       void* const* d_peer_bufs, int rank, int world, int rows, int head_dim,
                                       uint32_t* d_epochs, float* d_out, float* d_lse_out, int* d_status,
                                       pa_stream_t stream) {
    return exchange_launch(d_part_m, d_part_l, d_part_o, d_peer_bufs, rank, world, rows, head_dim, d_epochs, d_out,
                           d_lse_out, d_status, stream, 3);
}

// The two halves on their own.  SEND never waits for anybody; RECV polls until every rank's packets of the current
// step are in this rank's buffer.  (Several ranks emulated on ONE device must use them in this order -- all sends,
// then the receives -- so that no kernel ever waits for another kernel on the same GPU.)
PA_API int pa_splitkv_exchange_send(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                                    void* const* d_peer_bufs, int rank, int world, int rows, int head_dim,
                                    const uint32_t* d_epochs, pa_stream_t stream) {
    return exchange_launch(d_part_m, d_part_l, d_part_o, d_peer_bufs, rank, world, rows, head_dim,
                           const_cast<uint32_t*>(d_epochs), nullptr, nullptr, nullptr, stream, 1);
}
PA_API int pa_splitkv_exchange_recv(void* const* d_peer_bufs, int rank, int world, int rows, int head_dim,
                                    uint32_t* d_epochs, float* d_out, float* d_lse_out, int* d_status,
                                    pa_stream_t stream) {
    return exchange_launch(nullptr, nullptr, nullptr, d_peer_bufs, rank, world, rows, head_dim, d_epochs, d_out, d_lse_out,
                           d_status, stream, 2);
}

// ---- NCCL form (SURVEY 8b "pa_nccl_* init / allgather-combine", 8e) ---------------------------------------
PA_API int pa_nccl_unique_id(unsigned char* id128) {
    PA_CHECK_ARG(id128);
    NcclApi& n = nccl();
    if (!n.ok) return PA_ERR_UNSUPPORTED;
    return n.GetUniqueId(id128) == 0 ? PA_OK : PA_ERR_NCCL;
}

PA_API int pa_nccl_init(const unsigned char* id128, int rank, int world, void** comm) {
    PA_CHECK_ARG(id128 && comm && world > 0 && rank >= 0 && rank < world);
    NcclApi& n = nccl();
    if (!n.ok) return PA_ERR_UNSUPPORTED;
    NcclId128 id;
    memcpy(id.b, id128, 128);
    return n.CommInitRank(comm, world, id, rank) == 0 ? PA_OK : PA_ERR_NCCL;
}

PA_API int pa_nccl_destroy(void* comm) {
    PA_CHECK_ARG(comm);
    NcclApi& n = nccl();
    if (!n.ok) return PA_ERR_UNSUPPORTED;
    return n.CommDestroy(comm) == 0 ? PA_OK : PA_ERR_NCCL;
}

PA_API size_t pa_nccl_gather_bytes(int world, int rows, int head_dim) {
    if (world <= 0 || rows <= 0 || head_dim <= 0) return 0;
    return (size_t)world * rows * (head_dim + 2) * sizeof(float);
}

// Three all-gathers in one NCCL group straight into the [n_parts][rows] / [n_parts][rows][D] layout of
// pa_lse_combine (no pack / unpack passes), then the combine kernel, all on `stream`.
PA_API int pa_nccl_allgather_combine(void* comm, int world, const float* d_part_m, const float* d_part_l,
                                     const float* d_part_o, int rows, int head_dim, void* d_gather_ws,
                                     size_t gather_bytes, float* d_out, float* d_lse_out, pa_stream_t stream) {
    PA_CHECK_ARG(comm && d_part_m && d_part_l && d_part_o && d_gather_ws && d_out && world > 0 && rows >= 0 && head_dim > 0);
    PA_CHECK_ARG(gather_bytes >= pa_nccl_gather_bytes(world, rows, head_dim));
    if (rows == 0) return PA_OK;
    NcclApi& n = nccl();
    if (!n.ok) return PA_ERR_UNSUPPORTED;
    float* gm = static_cast<float*>(d_gather_ws);
    float* gl = gm + (size_t)world * rows;
    float* go = gl + (size_t)world * rows;
    cudaStream_t st = as_stream(stream);
    constexpr int kNcclFloat = 7;
    int rc = n.GroupStart();
    rc |= n.AllGather(d_part_m, gm, (size_t)rows, kNcclFloat, comm, st);
    rc |= n.AllGather(d_part_l, gl, (size_t)rows, kNcclFloat, comm, st);
    rc |= n.AllGather(d_part_o, go, (size_t)rows * head_dim, kNcclFloat, comm, st);
    rc |= n.GroupEnd();
    if (rc != 0) return PA_ERR_NCCL;
    return pa_lse_combine(gm, gl, go, world, rows, head_dim, d_out, d_lse_out, stream);
}
