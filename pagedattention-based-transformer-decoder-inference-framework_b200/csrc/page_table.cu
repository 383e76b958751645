// page_table.cu -- device-resident page table, page gather and KV append.
//
// Replaces kv_cache/page_table.{hpp,cpp} and the addressing of
// kv_cache/kv_tile_cache.hpp:21-34.  All of it is HBM-bound integer/byte
// work: 128-bit coalesced moves, grids sized in multiples of the SM count.
#include <mutex>

#include "pa_common.cuh"

namespace pa {

const DeviceInfo& device_info() {
    // One entry per device ordinal; decode kernels size their grids from it.
    static DeviceInfo infos[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        static DeviceInfo bad;
        return bad;
    }
    std::lock_guard<std::mutex> lk(mu);
    DeviceInfo& d = infos[dev];
    if (!d.ok) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
            d.sm_count = p.multiProcessorCount;
            d.cc_major = p.major;
            d.cc_minor = p.minor;
            d.max_smem_optin = (int)p.sharedMemPerBlockOptin;
            d.ok = true;
        }
    }
    return d;
}

// page_table.hpp:39-49 -- index + bounds-checked lookup.
__device__ __forceinline__ int pt_lookup(const int32_t* __restrict__ table, int64_t total_entries,
                                         int beam, int head, int tile, int num_heads,
                                         int num_tiles) {
    // The reference computes the flat index in 32-bit int (page_table.hpp:41); out-of-range
    // components therefore alias into other rows.  We keep the flat-index bounds check
    // (idx < 0 || idx >= total) and do the arithmetic in 64 bit.
    int64_t idx = (int64_t)beam * ((int64_t)num_heads * num_tiles) + (int64_t)head * num_tiles + tile;
    if (idx < 0 || idx >= total_entries) return -1;
    return table[idx];
}

__global__ void page_table_fill_kernel(int32_t* __restrict__ table, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // 128-bit stores over the aligned body.
    int64_t n4 = n / 4;
    int4* t4 = reinterpret_cast<int4*>(table);
    int4 vv = make_int4(v, v, v, v);
    for (int64_t j = i; j < n4; j += stride) t4[j] = vv;
    for (int64_t j = n4 * 4 + i; j < n; j += stride) table[j] = v;
}

__global__ void page_table_update_kernel(int32_t* __restrict__ table, int64_t n_entries,
                                         const int32_t* __restrict__ idx,
                                         const int32_t* __restrict__ page, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t k = idx[i];
    if (k < 0 || k >= n_entries) return;
    table[k] = page[i];
}

__global__ void page_table_lookup_kernel(const int32_t* __restrict__ table, int num_beams,
                                         int num_heads, int num_tiles,
                                         const int32_t* __restrict__ beam,
                                         const int32_t* __restrict__ head,
                                         const int32_t* __restrict__ tile,
                                         int32_t* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = pt_lookup(table, (int64_t)num_beams * num_heads * num_tiles, beam[i], head[i], tile[i],
                       num_heads, num_tiles);
}

// One warp per page: 32 lanes x 16 B per iteration.  page_bytes % 16 == 0.
__global__ void kv_gather_kernel(const uint8_t* __restrict__ pool, uint8_t* __restrict__ dense,
                                 const int32_t* __restrict__ table, int num_beams, int num_heads,
                                 int num_tiles, int total_pages, int64_t page_bytes,
                                 const int32_t* __restrict__ beam_ids, int64_t n_pages_out,
                                 uint32_t fill_word) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total_entries = (int64_t)num_beams * num_heads * num_tiles;
    const int64_t vecs = page_bytes / 16;
    for (int64_t p = warp0; p < n_pages_out; p += nwarps) {
        int tile = (int)(p % num_tiles);
        int64_t rh = p / num_tiles;
        int head = (int)(rh % num_heads);
        int r = (int)(rh / num_heads);
        int beam = beam_ids ? beam_ids[r] : r;
        int page = pt_lookup(table, total_entries, beam, head, tile, num_heads, num_tiles);
        uint4* dst = reinterpret_cast<uint4*>(dense + p * page_bytes);
        if (page < 0 || page >= total_pages) {  // kv_tile_cache.hpp:23 -> nullptr
            uint4 f = make_uint4(fill_word, fill_word, fill_word, fill_word);
            for (int64_t v = lane; v < vecs; v += 32) dst[v] = f;
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(pool + (int64_t)page * page_bytes);
            for (int64_t v = lane; v < vecs; v += 32) dst[v] = ldg_stream_128(src + v);
        }
    }
}

// Append.  One warp per (row r, head h).  MODE 0: raw copy of elem_bytes rows
// (fp16 in, fp16 pool); MODE 1: f32 -> f16 round-to-nearest-even; MODE 2: f32 -> int8
// with the per-row minmax scale (int8_quant.cpp:59-64 then :15-28); MODE 3: raw copy of f32 rows into an
// fp32 pool (KVTileCache<float>, kv_tile_cache.cpp:127).
template <int MODE>
__global__ void kv_append_kernel(void* __restrict__ k_pool, void* __restrict__ v_pool,
                                 float* __restrict__ k_scales, float* __restrict__ v_scales,
                                 const int32_t* __restrict__ table, int num_beams, int num_heads,
                                 int num_tiles, int total_pages, int tile_size, int head_dim,
                                 const void* __restrict__ new_k, const void* __restrict__ new_v,
                                 const int32_t* __restrict__ beam_ids,
                                 const int32_t* __restrict__ positions, int R) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (int64_t)R * num_heads) return;
    const int h = (int)(w % num_heads);
    const int r = (int)(w / num_heads);
    const int beam = beam_ids ? beam_ids[r] : r;
    const int pos = positions[r];
    // A position past the table's capacity writes NOTHING.  The reference's lookup only bounds the FLAT index
    // (page_table.hpp:44-49), so tile_id >= num_tiles aliases into the next head's / beam's entries and
    // get_write_ptr (kv_tile_cache.hpp:29-34) would hand out somebody else's page -- silent corruption of another
    // sequence's cache (found by tests/test_bounds_guards.py).  Reads keep the reference's flat-index rule.
    if (pos < 0 || pos / tile_size >= num_tiles || (unsigned)beam >= (unsigned)num_beams) return;
    const int page = pt_lookup(table, (int64_t)num_beams * num_heads * num_tiles, beam, h,
                               pos / tile_size, num_heads, num_tiles);
    if (page < 0 || page >= total_pages) return;  // get_write_ptr -> nullptr
    const int row = pos % tile_size;
    const int64_t dst_el = ((int64_t)page * tile_size + row) * head_dim;
    const int64_t src_el = ((int64_t)r * num_heads + h) * head_dim;

    if (MODE == 3) {
        const float* sk = static_cast<const float*>(new_k) + src_el;
        const float* sv = static_cast<const float*>(new_v) + src_el;
        float* dk = static_cast<float*>(k_pool) + dst_el;
        float* dv = static_cast<float*>(v_pool) + dst_el;
        for (int d = lane * 4; d < head_dim; d += 128) {
            *reinterpret_cast<uint4*>(dk + d) = *reinterpret_cast<const uint4*>(sk + d);
            *reinterpret_cast<uint4*>(dv + d) = *reinterpret_cast<const uint4*>(sv + d);
        }
    } else if (MODE == 0) {
        const __half* sk = static_cast<const __half*>(new_k) + src_el;
        const __half* sv = static_cast<const __half*>(new_v) + src_el;
        __half* dk = static_cast<__half*>(k_pool) + dst_el;
        __half* dv = static_cast<__half*>(v_pool) + dst_el;
        for (int d = lane * 8; d < head_dim; d += 256) {
            *reinterpret_cast<uint4*>(dk + d) = *reinterpret_cast<const uint4*>(sk + d);
            *reinterpret_cast<uint4*>(dv + d) = *reinterpret_cast<const uint4*>(sv + d);
        }
    } else if (MODE == 1) {
        const float* sk = static_cast<const float*>(new_k) + src_el;
        const float* sv = static_cast<const float*>(new_v) + src_el;
        __half* dk = static_cast<__half*>(k_pool) + dst_el;
        __half* dv = static_cast<__half*>(v_pool) + dst_el;
        for (int d = lane * 4; d < head_dim; d += 128) {
            float4 a = *reinterpret_cast<const float4*>(sk + d);
            float4 b = *reinterpret_cast<const float4*>(sv + d);
            __half2 a0 = __floats2half2_rn(a.x, a.y), a1 = __floats2half2_rn(a.z, a.w);
            __half2 b0 = __floats2half2_rn(b.x, b.y), b1 = __floats2half2_rn(b.z, b.w);
            uint2 ua, ub;
            ua.x = *reinterpret_cast<uint32_t*>(&a0); ua.y = *reinterpret_cast<uint32_t*>(&a1);
            ub.x = *reinterpret_cast<uint32_t*>(&b0); ub.y = *reinterpret_cast<uint32_t*>(&b1);
            *reinterpret_cast<uint2*>(dk + d) = ua;
            *reinterpret_cast<uint2*>(dv + d) = ub;
        }
    } else {
        const float* src[2] = {static_cast<const float*>(new_k) + src_el,
                               static_cast<const float*>(new_v) + src_el};
        int8_t* dst[2] = {static_cast<int8_t*>(k_pool) + dst_el, static_cast<int8_t*>(v_pool) + dst_el};
        float* sc[2] = {k_scales, v_scales};
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
            // compute_minmax_scale: abs_max = max(|min|, |max|) == max |x|.
            float am = 0.f;
            for (int d = lane; d < head_dim; d += 32) am = fmaxf(am, fabsf(src[kv][d]));
            am = warp_max(am);
            const float scale = __fdiv_rn(127.f, __fadd_rn(am, 1e-6f));
            if (lane == 0) sc[kv][(int64_t)page * tile_size + row] = scale;
            for (int d = lane * 4; d < head_dim; d += 128) {
                float4 x = *reinterpret_cast<const float4*>(src[kv] + d);
                float xs[4] = {x.x, x.y, x.z, x.w};
                uint32_t packed = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float rr = roundf(__fmul_rn(xs[e], scale));  // half away from zero
                    rr = fminf(127.f, fmaxf(-128.f, rr));
                    int q = (int)rr;
                    packed |= (uint32_t)(q & 0xff) << (8 * e);
                }
                *reinterpret_cast<uint32_t*>(dst[kv] + d) = packed;
            }
        }
    }
}

// Copy-on-write page copies: pool[dst[i]] := pool[src[i]] for K, V (and their scale rows).
// One warp per (copy, pool): 32 lanes x 16 B per iteration.
__global__ void kv_copy_pages_kernel(uint8_t* __restrict__ k_pool, uint8_t* __restrict__ v_pool,
                                     float* __restrict__ k_scales, float* __restrict__ v_scales,
                                     const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int n,
                                     int total_pages, int64_t page_bytes, int tile_size) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (int64_t)n * 2) return;
    const int i = (int)(w >> 1);
    const int s = src[i], d = dst[i];
    if (s < 0 || s >= total_pages || d < 0 || d >= total_pages || s == d) return;
    uint8_t* pool = (w & 1) ? v_pool : k_pool;
    const uint4* sp = reinterpret_cast<const uint4*>(pool + (int64_t)s * page_bytes);
    uint4* dp = reinterpret_cast<uint4*>(pool + (int64_t)d * page_bytes);
    for (int64_t v = lane; v < page_bytes / 16; v += 32) dp[v] = sp[v];
    float* sc = (w & 1) ? v_scales : k_scales;
    if (sc)
        for (int t = lane; t < tile_size; t += 32) sc[(int64_t)d * tile_size + t] = sc[(int64_t)s * tile_size + t];
}

static int grid_for(int64_t threads_needed, int block, int max_blocks) {
    int64_t b = (threads_needed + block - 1) / block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

}  // namespace pa

using namespace pa;

PA_API int pa_version(void) { return 100; }

PA_API const char* pa_error_string(int status) {
    switch (status) {
        case PA_OK: return "ok";
        case PA_ERR_INVALID_ARG: return "pa_b200: invalid argument";
        case PA_ERR_UNSUPPORTED: return "pa_b200: unsupported shape (head_dim in {64,128}, tile_size % 16 == 0)";
        case PA_ERR_WORKSPACE: return "pa_b200: workspace too small";
        case PA_ERR_NO_DEVICE: return "pa_b200: no sm_100 CUDA device is current";
        case PA_ERR_NCCL: return "pa_b200: an NCCL call failed";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "pa_b200: unknown error";
}

PA_API int pa_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    return PA_OK;
}

PA_API int pa_page_table_clear(int32_t* d_table, int64_t n_entries, pa_stream_t stream) {
    PA_CHECK_ARG(d_table && n_entries >= 0);
    if (n_entries == 0) return PA_OK;
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    int grid = grid_for((n_entries + 3) / 4, 256, d.sm_count * 8);
    page_table_fill_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_table, n_entries, -1);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_page_table_update(int32_t* d_table, int64_t n_entries, const int32_t* d_idx,
                                const int32_t* d_page, int n, pa_stream_t stream) {
    PA_CHECK_ARG(d_table && n_entries > 0 && n >= 0);
    if (n == 0) return PA_OK;
    PA_CHECK_ARG(d_idx && d_page);
    page_table_update_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(d_table, n_entries,
                                                                          d_idx, d_page, n);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_page_table_lookup(const int32_t* d_table, int num_beams, int num_heads,
                                int num_tiles, const int32_t* d_beam, const int32_t* d_head,
                                const int32_t* d_tile, int32_t* d_out, int n, pa_stream_t stream) {
    PA_CHECK_ARG(d_table && num_beams > 0 && num_heads > 0 && num_tiles > 0 && n >= 0);
    if (n == 0) return PA_OK;
    PA_CHECK_ARG(d_beam && d_head && d_tile && d_out);
    page_table_lookup_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(
        d_table, num_beams, num_heads, num_tiles, d_beam, d_head, d_tile, d_out, n);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_kv_gather(const void* d_pool, void* d_dense, const int32_t* d_table, int num_beams,
                        int num_heads, int num_tiles, int total_pages, int tile_size,
                        int head_dim, int elem_bytes, const int32_t* d_beam_ids, int R,
                        int fill_byte, pa_stream_t stream) {
    PA_CHECK_ARG(d_pool && d_dense && d_table);
    PA_CHECK_ARG(num_beams > 0 && num_heads > 0 && num_tiles > 0 && total_pages > 0 && R >= 0);
    PA_CHECK_ARG(tile_size > 0 && head_dim > 0 && (elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4));
    const int64_t page_bytes = (int64_t)tile_size * head_dim * elem_bytes;
    if (page_bytes % 16 != 0) return PA_ERR_UNSUPPORTED;
    PA_CHECK_ARG(((uintptr_t)d_pool % 16) == 0 && ((uintptr_t)d_dense % 16) == 0);
    if (R == 0) return PA_OK;
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    const int64_t n_pages_out = (int64_t)R * num_heads * num_tiles;
    int grid = grid_for(n_pages_out * 32, 256, d.sm_count * 8);
    uint32_t fb = (uint32_t)(fill_byte & 0xff);
    uint32_t fw = fb | (fb << 8) | (fb << 16) | (fb << 24);
    kv_gather_kernel<<<grid, 256, 0, as_stream(stream)>>>(
        static_cast<const uint8_t*>(d_pool), static_cast<uint8_t*>(d_dense), d_table, num_beams,
        num_heads, num_tiles, total_pages, page_bytes, d_beam_ids, n_pages_out, fw);
    PA_RETURN_LAUNCH_STATUS();
}

template <int MODE>
static int launch_append(void* k_pool, void* v_pool, float* k_scales, float* v_scales,
                         const int32_t* table, int num_beams, int num_heads, int num_tiles,
                         int total_pages, int tile_size, int head_dim, const void* new_k,
                         const void* new_v, const int32_t* beam_ids, const int32_t* positions,
                         int R, pa_stream_t stream) {
    PA_CHECK_ARG(k_pool && v_pool && table && new_k && new_v && positions);
    PA_CHECK_ARG(num_beams > 0 && num_heads > 0 && num_tiles > 0 && total_pages > 0 && R >= 0);
    PA_CHECK_ARG(tile_size > 0 && head_dim > 0);
    if (MODE == 2) PA_CHECK_ARG(k_scales && v_scales);
    if (head_dim % 8 != 0) return PA_ERR_UNSUPPORTED;
    if (R == 0) return PA_OK;
    const int64_t warps = (int64_t)R * num_heads;
    const int64_t blocks = (warps * 32 + 255) / 256;
    kv_append_kernel<MODE><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        k_pool, v_pool, k_scales, v_scales, table, num_beams, num_heads, num_tiles, total_pages,
        tile_size, head_dim, new_k, new_v, beam_ids, positions, R);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_kv_append_f16(void* d_k_pool, void* d_v_pool, const int32_t* d_table, int num_beams,
                            int num_heads, int num_tiles, int total_pages, int tile_size,
                            int head_dim, const void* d_new_k, const void* d_new_v,
                            const int32_t* d_beam_ids, const int32_t* d_positions, int R,
                            pa_stream_t stream) {
    return launch_append<0>(d_k_pool, d_v_pool, nullptr, nullptr, d_table, num_beams, num_heads,
                            num_tiles, total_pages, tile_size, head_dim, d_new_k, d_new_v,
                            d_beam_ids, d_positions, R, stream);
}

PA_API int pa_kv_append_f32(float* d_k_pool, float* d_v_pool, const int32_t* d_table, int num_beams,
                            int num_heads, int num_tiles, int total_pages, int tile_size, int head_dim,
                            const float* d_new_k, const float* d_new_v, const int32_t* d_beam_ids,
                            const int32_t* d_positions, int R, pa_stream_t stream) {
    return launch_append<3>(d_k_pool, d_v_pool, nullptr, nullptr, d_table, num_beams, num_heads,
                            num_tiles, total_pages, tile_size, head_dim, d_new_k, d_new_v,
                            d_beam_ids, d_positions, R, stream);
}

PA_API int pa_kv_append_f32_f16(void* d_k_pool, void* d_v_pool, const int32_t* d_table,
                                int num_beams, int num_heads, int num_tiles, int total_pages,
                                int tile_size, int head_dim, const float* d_new_k,
                                const float* d_new_v, const int32_t* d_beam_ids,
                                const int32_t* d_positions, int R, pa_stream_t stream) {
    return launch_append<1>(d_k_pool, d_v_pool, nullptr, nullptr, d_table, num_beams, num_heads,
                            num_tiles, total_pages, tile_size, head_dim, d_new_k, d_new_v,
                            d_beam_ids, d_positions, R, stream);
}

PA_API int pa_kv_append_f32_i8(int8_t* d_k_pool, int8_t* d_v_pool, float* d_k_scales,
                               float* d_v_scales, const int32_t* d_table, int num_beams,
                               int num_heads, int num_tiles, int total_pages, int tile_size,
                               int head_dim, const float* d_new_k, const float* d_new_v,
                               const int32_t* d_beam_ids, const int32_t* d_positions, int R,
                               pa_stream_t stream) {
    return launch_append<2>(d_k_pool, d_v_pool, d_k_scales, d_v_scales, d_table, num_beams,
                            num_heads, num_tiles, total_pages, tile_size, head_dim, d_new_k,
                            d_new_v, d_beam_ids, d_positions, R, stream);
}

PA_API int pa_kv_copy_pages(void* d_k_pool, void* d_v_pool, float* d_k_scales, float* d_v_scales,
                            const int32_t* d_src_pages, const int32_t* d_dst_pages, int n, int total_pages,
                            int tile_size, int head_dim, int elem_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_k_pool && d_v_pool && d_src_pages && d_dst_pages && n >= 0 && total_pages > 0);
    PA_CHECK_ARG(tile_size > 0 && head_dim > 0 && (elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4));
    const int64_t page_bytes = (int64_t)tile_size * head_dim * elem_bytes;
    if (page_bytes % 16 != 0) return PA_ERR_UNSUPPORTED;
    if (n == 0) return PA_OK;
    const int64_t threads = (int64_t)n * 2 * 32;
    kv_copy_pages_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<uint8_t*>(d_k_pool), static_cast<uint8_t*>(d_v_pool), d_k_scales, d_v_scales, d_src_pages,
        d_dst_pages, n, total_pages, page_bytes, tile_size);
    PA_RETURN_LAUNCH_STATUS();
}
