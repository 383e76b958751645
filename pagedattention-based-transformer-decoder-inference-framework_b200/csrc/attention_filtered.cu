// attention_filtered.cu -- decode attention WITH the reference's in-attention top-k / top-p filter and its optional
// side outputs (pre-softmax logits, post-filter attention weights).
//
// This is the CPU kernel's algorithm stage by stage (attention_cpu/cpu_attention_kernel.cpp:61-126):
//   K pass   scores[t] = dot(q, K[t]) / temperature, -1e9 where the tile is unmapped                 (:61-87)
//   softmax  over ALL T scores: exp(x - max) / (sum + 1e-6)                       (softmax_lut.cpp:203-231)
//   filter   apply_topk_topp_filter: rank-based zeroing, NO renormalisation       (softmax_lut.cpp:233-256)
//   V pass   out[d] = sum_t p[t] * V[t][d], unmapped tiles skipped                                 (:103-117)
// and CPUAttentionOutput::logits / ::attention_weights (cpu_attention_kernel.hpp:34-39).  The filter needs every
// probability of a row before any of them can be used, so this cannot be the one-pass online-softmax kernel of
// paged_decode.cu; it is the explicit three-stage form, for the callers that ask for filtering or the side outputs
// (defaults top_k = 0, top_p = 1 keep the hot path).  Softmax and filter are the kernels of sampling.cu
// (pa_softmax_temperature, pa_topk_topp_filter: radix select with the reference's tie order).
// The temperature is applied ONCE, as in the hot path.
#include "pa_common.cuh"

extern "C" int pa_softmax_temperature(const float*, int, int, float, float*, pa_stream_t);
extern "C" int pa_topk_topp_filter(float*, int, int, int, float, int, float, pa_stream_t);

namespace pa {

struct FilteredArgs {
    const float* q;
    float* out;
    const uint8_t* k_pool;
    const uint8_t* v_pool;
    const float* k_scales;
    const float* v_scales;
    const int32_t* table;
    const int32_t* beam_ids;
    const int32_t* ctx_lens;
    const float* rope;
    float* scores;  // [rows][T]
    const float* probs;  // [rows][T]
    int num_beams, H, num_tiles, total_pages, B, T, D, tile_size;
    float inv_temperature;
};

template <int KV>
__device__ __forceinline__ float load_elem(const uint8_t* pool, const float* scales, int64_t tok, int D, int d) {
    if (KV == 0) return __half2float(reinterpret_cast<const __half*>(pool)[tok * D + d]);
    if (KV == 2) return reinterpret_cast<const float*>(pool)[tok * D + d];
    return __fdiv_rn((float)reinterpret_cast<const int8_t*>(pool)[tok * D + d], scales[tok]);  // int8_quant.cpp:41
}

__device__ __forceinline__ int filtered_page(const FilteredArgs& a, int beam, int h, int tile) {
    const int64_t idx = ((int64_t)beam * a.H + h) * a.num_tiles + tile;
    if (beam < 0 || beam >= a.num_beams || tile < 0 || tile >= a.num_tiles) return -1;
    const int page = a.table[idx];
    return (page < 0 || page >= a.total_pages) ? -1 : page;
}

// K pass: one warp per (row, token): lanes split the head dimension.  scores beyond the row's context = -inf
// (they take no part in the softmax); unmapped tiles inside it = -1e9 (the reference's initial value).
template <int KV>
__global__ void __launch_bounds__(256) filtered_scores_kernel(const FilteredArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t rows = (int64_t)a.B * a.H;
    if (w >= rows * a.T) return;
    const int64_t row = w / a.T;
    const int t = (int)(w - row * a.T);
    const int b = (int)(row / a.H), h = (int)(row % a.H);
    int ctx = a.ctx_lens ? a.ctx_lens[b] : a.T;
    ctx = ctx < 0 ? 0 : (ctx > a.T ? a.T : ctx);
    float s = -INFINITY;
    if (t < ctx) {
        const int beam = a.beam_ids ? a.beam_ids[b] : b;
        const int page = filtered_page(a, beam, h, t / a.tile_size);
        s = -1e9f;
        if (page >= 0) {
            const int64_t tok = (int64_t)page * a.tile_size + t % a.tile_size;
            float acc = 0.f;
            for (int d = lane * 2; d < a.D; d += 64) {
                float q0 = a.q[row * a.D + d], q1 = a.q[row * a.D + d + 1];
                if (a.rope) {  // cpu_attention_kernel.cpp:13-19
                    const float cs = a.rope[d], sn = a.rope[d + 1];
                    const float r0 = q0 * cs - q1 * sn, r1 = q0 * sn + q1 * cs;
                    q0 = r0;
                    q1 = r1;
                }
                acc = fmaf(q0, load_elem<KV>(a.k_pool, a.k_scales, tok, a.D, d), acc);
                acc = fmaf(q1, load_elem<KV>(a.k_pool, a.k_scales, tok, a.D, d + 1), acc);
            }
            s = warp_sum(acc) * a.inv_temperature;
        }
    }
    if (lane == 0) a.scores[row * a.T + t] = s;
}

// V pass: one CTA per row, thread d owns output dimension d (coalesced V rows, broadcast probabilities).
template <int KV>
__global__ void __launch_bounds__(128) filtered_pv_kernel(const FilteredArgs a) {
    const int64_t row = blockIdx.x;
    const int b = (int)(row / a.H), h = (int)(row % a.H);
    int ctx = a.ctx_lens ? a.ctx_lens[b] : a.T;
    ctx = ctx < 0 ? 0 : (ctx > a.T ? a.T : ctx);
    const int beam = a.beam_ids ? a.beam_ids[b] : b;
    const float* p = a.probs + row * a.T;
    for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
        float acc = 0.f;
        for (int t0 = 0; t0 < ctx; t0 += a.tile_size) {
            const int page = filtered_page(a, beam, h, t0 / a.tile_size);
            if (page < 0) continue;  // :108 skip
            const int n = min(a.tile_size, ctx - t0);
            for (int i = 0; i < n; ++i) {
                const float w = p[t0 + i];
                if (w != 0.f) acc = fmaf(w, load_elem<KV>(a.v_pool, a.v_scales, (int64_t)page * a.tile_size + i, a.D, d), acc);
            }
        }
        a.out[row * a.D + d] = acc;
    }
}

template <int KV>
static int launch_filtered(FilteredArgs& a, int top_k, float top_p, float* weights, cudaStream_t st, pa_stream_t stream) {
    const int64_t rows = (int64_t)a.B * a.H;
    const int64_t warps = rows * a.T;
    filtered_scores_kernel<KV><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    int rc = pa_softmax_temperature(a.scores, (int)rows, a.T, 1.0f, weights, stream);
    if (rc != PA_OK) return rc;
    if (top_k > 0 || top_p < 1.0f) {
        rc = pa_topk_topp_filter(weights, (int)rows, a.T, top_k, top_p, -1, 0.f, stream);
        if (rc != PA_OK) return rc;
    }
    a.probs = weights;
    filtered_pv_kernel<KV><<<(unsigned)rows, 128, 0, st>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? PA_OK : (int)e;
}

}  // namespace pa

using namespace pa;

PA_API size_t pa_attention_filtered_workspace_bytes(int B, int num_heads, int T) {
    if (B <= 0 || num_heads <= 0 || T <= 0) return 0;
    return (size_t)2 * B * num_heads * T * sizeof(float);
}

PA_API int pa_paged_attention_filtered(const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                                       const float* d_k_scales, const float* d_v_scales, int kv_kind,
                                       const int32_t* d_table, int num_beams, int num_heads, int num_tiles, int total_pages,
                                       const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B, int T, int head_dim,
                                       int tile_size, float temperature, const float* d_rope, int top_k, float top_p,
                                       float* d_logits_out, float* d_weights_out, void* d_workspace, size_t workspace_bytes,
                                       pa_stream_t stream) {
    PA_CHECK_ARG(d_q && d_out && d_k_pool && d_v_pool && d_table);
    PA_CHECK_ARG(num_beams > 0 && num_heads > 0 && num_tiles > 0 && total_pages > 0 && B >= 0 && T >= 0);
    PA_CHECK_ARG(temperature != 0.f && tile_size > 0 && head_dim > 0 && head_dim % 2 == 0);
    PA_CHECK_ARG(kv_kind >= 0 && kv_kind <= 2 && (kv_kind != 1 || (d_k_scales && d_v_scales)));
    PA_CHECK_ARG(T <= num_tiles * tile_size);
    if (B == 0 || T == 0) return PA_OK;
    const size_t n = (size_t)B * num_heads * T;
    float* ws = static_cast<float*>(d_workspace);
    size_t used = 0;
    float* logits = d_logits_out;
    if (!logits) {
        logits = ws;
        used += n;
    }
    float* weights = d_weights_out;
    if (!weights) {
        weights = ws + used;
        used += n;
    }
    if (used && (!d_workspace || workspace_bytes < used * sizeof(float))) return PA_ERR_WORKSPACE;
    FilteredArgs a{};
    a.q = d_q; a.out = d_out;
    a.k_pool = static_cast<const uint8_t*>(d_k_pool);
    a.v_pool = static_cast<const uint8_t*>(d_v_pool);
    a.k_scales = d_k_scales; a.v_scales = d_v_scales;
    a.table = d_table; a.beam_ids = d_beam_ids; a.ctx_lens = d_ctx_lens; a.rope = d_rope;
    a.scores = logits;
    a.num_beams = num_beams; a.H = num_heads; a.num_tiles = num_tiles; a.total_pages = total_pages;
    a.B = B; a.T = T; a.D = head_dim; a.tile_size = tile_size;
    a.inv_temperature = 1.0f / temperature;
    cudaStream_t st = as_stream(stream);
    if (kv_kind == 0) return launch_filtered<0>(a, top_k, top_p, weights, st, stream);
    if (kv_kind == 1) return launch_filtered<1>(a, top_k, top_p, weights, st, stream);
    return launch_filtered<2>(a, top_k, top_p, weights, st, stream);
}
