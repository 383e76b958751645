// prefill.cu -- multi-query (prefill / chunked-prefill) attention over the paged KV cache:
// SURVEY 8(f) row 1.  The reference's headers promise it -- q/out laid out [B, H, T, D] and an
// `is_prefill` flag (attention/attention_config.hpp:8-9,17), causal mask helper
// attention/attention_kernel_utils.cuh:70-79 (key_pos > query_pos is masked) -- but ship no code.
//
// Query t of row b attends the cached keys [0, ctx_start[b] + t] of table row beam(b).  Every
// (b, t) pair is an independent decode row that shares its pages with the other queries of b, so
// the op is expressed on the decode kernels: a pack kernel transposes q to [B*Tq, H, D] and writes
// the per-row table-row / context-length arrays (beam indirection + causal limit); for fp16 pages with
// head_dim 128 the tensor-core group kernel then treats 4 consecutive query positions as one group
// (every K/V unit staged once per 4 queries, mma.sync QK^T and PV), otherwise the split-KV decode kernel
// runs over B*Tq rows (the shared pages are served by the 126 MB L2); an unpack kernel transposes the
// result back.  K/V of the Tq new tokens must already be in the pages
// (pa_kv_append_* with one row per (b, t)).
#include "pa_common.cuh"

namespace pa {

__global__ void prefill_pack_kernel(const float* __restrict__ q, float* __restrict__ q_rows,
                                    int32_t* __restrict__ beam_rows, int32_t* __restrict__ ctx_rows,
                                    const int32_t* __restrict__ beam_ids, const int32_t* __restrict__ ctx_start,
                                    int B, int H, int Tq, int D) {
    // one CTA per (b, t): gathers the H head vectors of the query
    const int r = blockIdx.x;
    const int b = r / Tq, t = r - b * Tq;
    if (threadIdx.x == 0) {
        beam_rows[r] = beam_ids ? beam_ids[b] : b;
        ctx_rows[r] = (ctx_start ? ctx_start[b] : 0) + t + 1;
    }
    const int n4 = H * D / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const int h = (i * 4) / D, d = (i * 4) - h * D;
        const float4 v = *reinterpret_cast<const float4*>(q + (((int64_t)b * H + h) * Tq + t) * D + d);
        *reinterpret_cast<float4*>(q_rows + ((int64_t)r * H + h) * D + d) = v;
    }
}

__global__ void prefill_unpack_kernel(const float* __restrict__ out_rows, float* __restrict__ out, int B, int H,
                                      int Tq, int D) {
    const int r = blockIdx.x;
    const int b = r / Tq, t = r - b * Tq;
    const int n4 = H * D / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const int h = (i * 4) / D, d = (i * 4) - h * D;
        const float4 v = *reinterpret_cast<const float4*>(out_rows + ((int64_t)r * H + h) * D + d);
        *reinterpret_cast<float4*>(out + (((int64_t)b * H + h) * Tq + t) * D + d) = v;
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace pa

using namespace pa;

PA_API size_t pa_prefill_workspace_bytes(int B, int Tq, int num_heads, int head_dim, int num_tiles, int tile_size) {
    if (B < 0 || Tq <= 0 || num_heads <= 0 || head_dim <= 0 || num_tiles <= 0 || tile_size <= 0) return 0;
    const int64_t R = (int64_t)B * Tq;
    if (R > 0x7fffffff) return 0;
    const size_t rows_bytes = align256((size_t)R * num_heads * head_dim * sizeof(float));
    return 2 * rows_bytes + 2 * align256((size_t)R * sizeof(int32_t)) +
           pa_decode_workspace_bytes((int)R, num_heads, head_dim, num_tiles, tile_size) + 256;
}

static int prefill_entry(int kv, const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                         const float* d_k_scales, const float* d_v_scales, const int32_t* d_table, int num_beams,
                         int num_heads, int num_tiles, int total_pages, const int32_t* d_beam_ids,
                         const int32_t* d_ctx_start, int B, int Tq, int head_dim, int tile_size, float temperature,
                         void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_q && d_out && d_k_pool && d_v_pool && d_table && d_workspace);
    PA_CHECK_ARG(B >= 0 && Tq > 0 && num_heads > 0 && head_dim > 0 && head_dim % 4 == 0);
    if (B == 0) return PA_OK;
    const int64_t R64 = (int64_t)B * Tq;
    PA_CHECK_ARG(R64 <= 0x7fffffff);
    const int R = (int)R64;
    if (workspace_bytes < pa_prefill_workspace_bytes(B, Tq, num_heads, head_dim, num_tiles, tile_size))
        return PA_ERR_WORKSPACE;
    uint8_t* w = static_cast<uint8_t*>(d_workspace);
    const size_t rows_bytes = align256((size_t)R * num_heads * head_dim * sizeof(float));
    const size_t ids_bytes = align256((size_t)R * sizeof(int32_t));
    float* q_rows = reinterpret_cast<float*>(w);
    float* out_rows = reinterpret_cast<float*>(w + rows_bytes);
    int32_t* beam_rows = reinterpret_cast<int32_t*>(w + 2 * rows_bytes);
    int32_t* ctx_rows = reinterpret_cast<int32_t*>(w + 2 * rows_bytes + ids_bytes);
    void* dws = w + 2 * rows_bytes + 2 * ids_bytes;
    const size_t dws_bytes = workspace_bytes - (2 * rows_bytes + 2 * ids_bytes);
    cudaStream_t st = as_stream(stream);
    prefill_pack_kernel<<<R, 256, 0, st>>>(d_q, q_rows, beam_rows, ctx_rows, d_beam_ids, d_ctx_start, B, num_heads, Tq,
                                           head_dim);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int T = num_tiles * tile_size;  // upper bound; the per-row ctx array is the causal limit
    int stt;
    // fp16, head_dim 128: consecutive query positions of a row share every page, so W of them form a
    // "beam group" of the tensor-core group kernel (each K/V unit is staged once for W queries; the
    // per-row context length is the causal limit).  Otherwise one decode row per query.
    const int W = (Tq % 4 == 0) ? 4 : ((Tq % 2 == 0) ? 2 : 1);
    if (kv == 0 && head_dim == 128 && tile_size % 16 == 0 && ((uintptr_t)d_k_pool % 128 == 0) &&
        ((uintptr_t)d_v_pool % 128 == 0))
        stt = pa_paged_decode_f16_group(q_rows, out_rows, d_k_pool, d_v_pool, d_table, num_beams, num_heads,
                                        num_tiles, total_pages, beam_rows, ctx_rows, R, T, head_dim, tile_size,
                                        temperature, nullptr, W, nullptr, dws, dws_bytes, stream);
    else if (kv == 0)
        stt = pa_paged_decode_f16(q_rows, out_rows, d_k_pool, d_v_pool, d_table, num_beams, num_heads, num_tiles,
                                  total_pages, beam_rows, ctx_rows, R, T, head_dim, tile_size, temperature, nullptr,
                                  nullptr, dws, dws_bytes, stream);
    else
        stt = pa_paged_decode_i8(q_rows, out_rows, static_cast<const int8_t*>(d_k_pool),
                                 static_cast<const int8_t*>(d_v_pool), d_k_scales, d_v_scales, d_table, num_beams,
                                 num_heads, num_tiles, total_pages, beam_rows, ctx_rows, R, T, head_dim, tile_size,
                                 temperature, nullptr, nullptr, dws, dws_bytes, stream);
    if (stt != PA_OK) return stt;
    prefill_unpack_kernel<<<R, 256, 0, st>>>(out_rows, d_out, B, num_heads, Tq, head_dim);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_paged_prefill_f16(const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                                const int32_t* d_table, int num_beams, int num_heads, int num_tiles, int total_pages,
                                const int32_t* d_beam_ids, const int32_t* d_ctx_start, int B, int Tq, int head_dim,
                                int tile_size, float temperature, void* d_workspace, size_t workspace_bytes,
                                pa_stream_t stream) {
    return prefill_entry(0, d_q, d_out, d_k_pool, d_v_pool, nullptr, nullptr, d_table, num_beams, num_heads, num_tiles,
                         total_pages, d_beam_ids, d_ctx_start, B, Tq, head_dim, tile_size, temperature, d_workspace,
                         workspace_bytes, stream);
}

PA_API int pa_paged_prefill_i8(const float* d_q, float* d_out, const int8_t* d_k_pool, const int8_t* d_v_pool,
                               const float* d_k_scales, const float* d_v_scales, const int32_t* d_table, int num_beams,
                               int num_heads, int num_tiles, int total_pages, const int32_t* d_beam_ids,
                               const int32_t* d_ctx_start, int B, int Tq, int head_dim, int tile_size,
                               float temperature, void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_k_scales && d_v_scales);
    return prefill_entry(1, d_q, d_out, d_k_pool, d_v_pool, d_k_scales, d_v_scales, d_table, num_beams, num_heads,
                         num_tiles, total_pages, d_beam_ids, d_ctx_start, B, Tq, head_dim, tile_size, temperature,
                         d_workspace, workspace_bytes, stream);
}
