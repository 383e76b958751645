/*
 * pa_b200.h -- C-ABI of libpa_b200.so: the B200 (sm_100a) implementation of the
 * reference's decode hot path (paged attention over the tiled KV cache + INT8
 * quantise / dequantise / int8 matmul).  Plain pointers and sizes only; every
 * pointer named d_* (and every tensor argument unless stated otherwise) is a
 * DEVICE pointer; `stream` is a cudaStream_t passed as void*.
 *
 * Every entry point returns an int status: 0 = PA_OK, <0 = PA_ERR_* (argument /
 * support errors, nothing launched), >0 = a cudaError_t from the launch.
 * pa_error_string() renders any of them.  There is no CPU fallback anywhere.
 *
 * Each declaration cites the reference interface it replaces (paths relative to
 * the reference repository root).
 */
#ifndef PA_B200_H_
#define PA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_OK 0
#define PA_ERR_INVALID_ARG (-1)  /* null pointer, non-positive size, misaligned buffer */
#define PA_ERR_UNSUPPORTED (-2)  /* head_dim / tile_size / activation outside the built set */
#define PA_ERR_WORKSPACE (-3)    /* workspace too small; see pa_decode_workspace_bytes */
#define PA_ERR_NO_DEVICE (-4)    /* no sm_100 device is current */
#define PA_ERR_NCCL (-5)         /* an NCCL call failed (pa_nccl_*) */

#define PA_ACT_NONE 0
#define PA_ACT_RELU 1
#define PA_ACT_GELU 2 /* gelu_erf, dnnl_matmul_int8.cpp:48-49 */

typedef void* pa_stream_t; /* cudaStream_t */

/* ---------------------------------------------------------------- runtime */
int pa_version(void);
const char* pa_error_string(int status);
/* sm count / compute capability of the current device. */
int pa_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------- page table */
/* kv_cache/page_table.cpp:14-26,41-47  PageTable::init / clear:
 * every entry := -1 (unmapped). */
int pa_page_table_clear(int32_t* d_table, int64_t n_entries, pa_stream_t stream);

/* kv_cache/page_table.cpp:49-53 PageTable::assign and :64-67 remove, batched:
 * d_table[d_idx[i]] = d_page[i] for i < n (page -1 == remove).  Replaces the
 * reference's one blocking 4-byte cudaMemcpy per entry.  idx is the flat index
 * beam*(H*Tiles) + head*Tiles + tile (page_table.hpp:39-42); out-of-range
 * indices are ignored (the reference asserts). */
int pa_page_table_update(int32_t* d_table, int64_t n_entries, const int32_t* d_idx,
                         const int32_t* d_page, int n, pa_stream_t stream);

/* kv_cache/page_table.hpp:44-49 PageTable::lookup (device), batched: d_out[i] =
 * table entry of (d_beam[i], d_head[i], d_tile[i]) or -1 when the flat index
 * is out of range. */
int pa_page_table_lookup(const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                         const int32_t* d_beam, const int32_t* d_head, const int32_t* d_tile,
                         int32_t* d_out, int n, pa_stream_t stream);

/* -------------------------------------------------------- page pool: gather */
/* kv_cache/kv_tile_cache.hpp:21-26 KVTileCache<T>::get, materialised: copies
 * the pages of rows r < R (table row = d_beam_ids ? d_beam_ids[r] : r) of one
 * pool into dense [R, H, num_tiles*tile_size, D]; unmapped / out-of-range pages
 * are filled with `fill_byte`.  elem_bytes = sizeof(T) (1, 2 or 4).
 * Bit-exact byte mover. */
int pa_kv_gather(const void* d_pool, void* d_dense, const int32_t* d_table, int num_beams,
                 int num_heads, int num_tiles, int total_pages, int tile_size, int head_dim,
                 int elem_bytes, const int32_t* d_beam_ids, int R, int fill_byte,
                 pa_stream_t stream);

/* -------------------------------------------------------- page pool: append */
/* kv_cache/kv_tile_cache.hpp:29-34 KVTileCache<T>::get_write_ptr + the row
 * write the caller performs: row (d_positions[r] % tile_size) of page
 * (beam, head, d_positions[r] / tile_size) := new_k/new_v[r, head, :].
 * Unmapped pages are skipped.  new_k/new_v are [R, H, D].
 *   _f16      : new rows already fp16, byte copy.
 *   _f32_f16  : new rows fp32, rounded to nearest-even fp16 on write.
 *   _f32_i8   : new rows fp32; per (row, head) scale = compute_minmax_scale
 *               (int8_quant.cpp:59-64), q = batch_quantize (int8_quant.cpp:15-28);
 *               scale stored at d_*_scales[page*tile_size + row_in_page]. */
int pa_kv_append_f16(void* d_k_pool, void* d_v_pool, const int32_t* d_table, int num_beams,
                     int num_heads, int num_tiles, int total_pages, int tile_size, int head_dim,
                     const void* d_new_k, const void* d_new_v, const int32_t* d_beam_ids,
                     const int32_t* d_positions, int R, pa_stream_t stream);
/* fp32 rows into an fp32 pool (KVTileCache<float>): raw copy. */
int pa_kv_append_f32(float* d_k_pool, float* d_v_pool, const int32_t* d_table, int num_beams,
                     int num_heads, int num_tiles, int total_pages, int tile_size, int head_dim,
                     const float* d_new_k, const float* d_new_v, const int32_t* d_beam_ids,
                     const int32_t* d_positions, int R, pa_stream_t stream);
int pa_kv_append_f32_f16(void* d_k_pool, void* d_v_pool, const int32_t* d_table, int num_beams,
                         int num_heads, int num_tiles, int total_pages, int tile_size,
                         int head_dim, const float* d_new_k, const float* d_new_v,
                         const int32_t* d_beam_ids, const int32_t* d_positions, int R,
                         pa_stream_t stream);
int pa_kv_append_f32_i8(int8_t* d_k_pool, int8_t* d_v_pool, float* d_k_scales, float* d_v_scales,
                        const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                        int total_pages, int tile_size, int head_dim, const float* d_new_k,
                        const float* d_new_v, const int32_t* d_beam_ids,
                        const int32_t* d_positions, int R, pa_stream_t stream);

/* Copy-on-write support for beam search over shared-prefix pages (north star; no
 * reference implementation): pool[d_dst_pages[i]] := pool[d_src_pages[i]] for the K and V
 * pools (and the int8 scale rows when given), i < n.  Bit-exact byte mover. */
int pa_kv_copy_pages(void* d_k_pool, void* d_v_pool, float* d_k_scales, float* d_v_scales,
                     const int32_t* d_src_pages, const int32_t* d_dst_pages, int n, int total_pages,
                     int tile_size, int head_dim, int elem_bytes, pa_stream_t stream);

/* ------------------------------------------------ paged decode attention */
/* Bytes of scratch the decode entry points need for B rows x H heads of head_dim D over a
 * page table of num_tiles tiles of tile_size tokens. */
size_t pa_decode_workspace_bytes(int B, int num_heads, int head_dim, int num_tiles, int tile_size);

/*
 * attention/paged_flash_attention_kernel_fused.cu:5-90 (launched by
 * attention/attention_tile_launcher.hpp:35-90, dispatched by
 * attention/attention_cuda.cu:41-95), with the math of
 * attention_cpu/cpu_attention_kernel.cpp:36-129 (global softmax; SURVEY App. A
 * D1-D9): for every (b, h)
 *     beam   = d_beam_ids ? d_beam_ids[b] : b
 *     s[t]   = dot(q[b,h,:], K[beam,h,t,:]) / temperature      t < ctx(b)
 *     p      = exp(s - max s) / (sum exp(s - max s) + 1e-6)
 *     out[b,h,:] = sum_t p[t] * V[beam,h,t,:]
 * K/V rows come from pages: page = table[beam][h][t / tile_size]; unmapped or
 * out-of-range pages contribute nothing.  ctx(b) = d_ctx_lens ? d_ctx_lens[b] : T.
 * q, out: [B, H, D] f32 (out may alias q).  Pools: [total_pages][tile_size][D].
 * d_rope: optional [D] interleaved cos/sin, pairwise rotation of q
 * (cpu_attention_kernel.cpp:13-19); NULL = off.  d_lse_out: optional [B, H]
 * log-sum-exp of the scaled scores (replaces the racy rerank_scores side
 * output, ...fused.cu:53-55).  head_dim in {64, 128}; tile_size % 16 == 0.
 *
 * _f16: fp16 K/V pages, split-KV grid kernel, 128-bit page-indirect global
 *       loads, warp-shuffle online softmax, LSE combine.
 * _f16_overlap: paged_flash_attention_kernel_fused_overlap.cu:5-91 -- same
 *       results through the persistent kernel that stages pages into shared
 *       memory with the TMA bulk-copy engine (multi-stage mbarrier ring) so
 *       loads and math overlap.
 * _i8 / _i8_overlap: int8 K/V pages with per-(page,row) f32 scales, dequantised
 *       in the kernel as q/scale (int8_quant.cpp:46-57).
 * _f32 / _f32_overlap: fp32 K/V pages -- KVTileCache<float>, the instantiation
 *       AttentionCUDA::forward takes (attention_config.hpp:15, kv_tile_cache.cpp:127);
 *       the same kernels on 4-byte elements.
 */
int pa_paged_decode_f16(const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                        const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                        int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_lens,
                        int B, int T, int head_dim, int tile_size, float temperature,
                        const float* d_rope, float* d_lse_out, void* d_workspace,
                        size_t workspace_bytes, pa_stream_t stream);
int pa_paged_decode_f16_overlap(const float* d_q, float* d_out, const void* d_k_pool,
                                const void* d_v_pool, const int32_t* d_table, int num_beams,
                                int num_heads, int num_tiles, int total_pages,
                                const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B,
                                int T, int head_dim, int tile_size, float temperature,
                                const float* d_rope, float* d_lse_out, void* d_workspace,
                                size_t workspace_bytes, pa_stream_t stream);
int pa_paged_decode_i8(const float* d_q, float* d_out, const int8_t* d_k_pool,
                       const int8_t* d_v_pool, const float* d_k_scales, const float* d_v_scales,
                       const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                       int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_lens,
                       int B, int T, int head_dim, int tile_size, float temperature,
                       const float* d_rope, float* d_lse_out, void* d_workspace,
                       size_t workspace_bytes, pa_stream_t stream);
int pa_paged_decode_i8_overlap(const float* d_q, float* d_out, const int8_t* d_k_pool,
                               const int8_t* d_v_pool, const float* d_k_scales,
                               const float* d_v_scales, const int32_t* d_table, int num_beams,
                               int num_heads, int num_tiles, int total_pages,
                               const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B, int T,
                               int head_dim, int tile_size, float temperature,
                               const float* d_rope, float* d_lse_out, void* d_workspace,
                               size_t workspace_bytes, pa_stream_t stream);

int pa_paged_decode_f32(const float* d_q, float* d_out, const float* d_k_pool, const float* d_v_pool,
                        const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                        int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_lens,
                        int B, int T, int head_dim, int tile_size, float temperature,
                        const float* d_rope, float* d_lse_out, void* d_workspace,
                        size_t workspace_bytes, pa_stream_t stream);
int pa_paged_decode_f32_overlap(const float* d_q, float* d_out, const float* d_k_pool,
                                const float* d_v_pool, const int32_t* d_table, int num_beams,
                                int num_heads, int num_tiles, int total_pages,
                                const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B,
                                int T, int head_dim, int tile_size, float temperature,
                                const float* d_rope, float* d_lse_out, void* d_workspace,
                                size_t workspace_bytes, pa_stream_t stream);

/* Decode attention WITH the reference's in-attention top-k / top-p filter and its optional side outputs: the CPU
 * kernel's algorithm stage by stage (attention_cpu/cpu_attention_kernel.cpp:61-126) -- K pass, softmax over all T
 * scores (softmax_lut.cpp:203-231), apply_topk_topp_filter (rank-based zeroing, no renormalisation,
 * softmax_lut.cpp:233-256), V pass.  The filter needs a row's every probability before any is used, so this is
 * the explicit three-stage form, not the one-pass kernels above; callers use it only when top_k > 0, top_p < 1
 * or a side output is wanted (cpu_attention_kernel.hpp:20-21 defaults keep the hot path).
 *   kv_kind: 0 fp16 pages, 1 int8 pages + scales, 2 fp32 pages.
 *   d_logits_out  (optional) [B,H,T]: pre-softmax scores, -1e9 for unmapped tiles (CPUAttentionOutput::logits),
 *                 -inf past a row's context;  d_weights_out (optional) [B,H,T]: post-filter probabilities
 *                 (CPUAttentionOutput::attention_weights).  Whichever is NULL lives in d_workspace
 *                 (>= pa_attention_filtered_workspace_bytes(B, num_heads, T)).  Temperature applied once (D3). */
size_t pa_attention_filtered_workspace_bytes(int B, int num_heads, int T);
int pa_paged_attention_filtered(const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                                const float* d_k_scales, const float* d_v_scales, int kv_kind,
                                const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                                int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B,
                                int T, int head_dim, int tile_size, float temperature, const float* d_rope,
                                int top_k, float top_p, float* d_logits_out, float* d_weights_out,
                                void* d_workspace, size_t workspace_bytes, pa_stream_t stream);

/* Beam-aware shared-prefix decode (north star; the reference's only hook is the
 * beam_ids[b] -> page-table-row indirection, ...fused.cu:22).  Rows b = g*beam_width ..
 * g*beam_width + beam_width-1 form beam group g.  Same results as pa_paged_decode_f16, but
 * a K/V page whose id is the same for several beams of a group (shared prefix /
 * not-yet-copied copy-on-write page) is read from HBM ONCE per group and applied to all
 * those beams' queries on the tensor cores (mma.sync m16n8k16, fp16 operands split hi/lo so
 * scores and outputs keep fp32 accuracy).  Pages that differ are staged per distinct id.
 * Requires head_dim == 128, beam_width <= 4, B % beam_width == 0; otherwise
 * PA_ERR_UNSUPPORTED and the caller uses pa_paged_decode_f16[_overlap].  d_ctx_lens (optional)
 * gives every row its own context length (tokens at or past it are masked for that row only):
 * with consecutive query positions as the rows of a group this is causal prefill, 4 queries
 * per K/V byte. */
int pa_paged_decode_f16_group(const float* d_q, float* d_out, const void* d_k_pool,
                              const void* d_v_pool, const int32_t* d_table, int num_beams,
                              int num_heads, int num_tiles, int total_pages,
                              const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B, int T,
                              int head_dim, int tile_size, float temperature, const float* d_rope,
                              int beam_width, float* d_lse_out, void* d_workspace,
                              size_t workspace_bytes, pa_stream_t stream);

/* Beam groups over INT8 pages: the per-row streaming kernel through the beam_ids indirection (the tensor-core group
 * kernel is fp16-only).  Same results as pa_paged_decode_i8_overlap; shared-prefix pages are served from L2 after
 * their first read (evict_first is off when d_beam_ids is given), so HBM sees roughly the unique bytes. */
int pa_paged_decode_i8_group(const float* d_q, float* d_out, const int8_t* d_k_pool, const int8_t* d_v_pool,
                             const float* d_k_scales, const float* d_v_scales, const int32_t* d_table,
                             int num_beams, int num_heads, int num_tiles, int total_pages,
                             const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B, int T, int head_dim,
                             int tile_size, float temperature, const float* d_rope, int beam_width,
                             float* d_lse_out, void* d_workspace, size_t workspace_bytes, pa_stream_t stream);

/* Split-KV across GPUs (north-star long-context mode): same computation over
 * this rank's pages only, emitting UN-normalised partials for an LSE combine:
 *   d_part_m[b,h] = max_t s[t] (natural-log units; -inf if no key)
 *   d_part_l[b,h] = sum_t exp(s[t]-m),  d_part_o[b,h,:] = sum_t exp(s[t]-m) V[t,:]
 * (no positional term enters the math, so a rank needs no token offset: its page-table rows simply hold
 * its own tile range).  One launch: the streaming kernel merges each row when its last chunk finishes.
 * _i8: int8 pages with per-(page,row) scales. */
int pa_paged_decode_f16_partial(const float* d_q, float* d_part_m, float* d_part_l,
                                float* d_part_o, const void* d_k_pool, const void* d_v_pool,
                                const int32_t* d_table, int num_beams, int num_heads,
                                int num_tiles, int total_pages, const int32_t* d_beam_ids,
                                const int32_t* d_ctx_lens, int B, int T, int head_dim,
                                int tile_size, float temperature, const float* d_rope,
                                void* d_workspace, size_t workspace_bytes, pa_stream_t stream);
int pa_paged_decode_i8_partial(const float* d_q, float* d_part_m, float* d_part_l, float* d_part_o,
                               const int8_t* d_k_pool, const int8_t* d_v_pool, const float* d_k_scales,
                               const float* d_v_scales, const int32_t* d_table, int num_beams,
                               int num_heads, int num_tiles, int total_pages, const int32_t* d_beam_ids,
                               const int32_t* d_ctx_lens, int B, int T, int head_dim, int tile_size,
                               float temperature, const float* d_rope, void* d_workspace,
                               size_t workspace_bytes, pa_stream_t stream);

/* LSE combine of n_parts partials (layout [n_parts][rows] for m/l and
 * [n_parts][rows][D] for o, as gathered from n_parts ranks):
 *   M = max_i m_i; w_i = exp(m_i - M); out = sum w_i o_i / (sum w_i l_i + 1e-6). */
int pa_lse_combine(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                   int n_parts, int rows, int head_dim, float* d_out, float* d_lse_out,
                   pa_stream_t stream);

/* --------------------------------------------------------------- int8_quant */
/* attention_cpu/int8_quant.cpp:5-13  q = clamp(round_half_away(x*scale),-128,127) */
int pa_quantize_i8(const float* d_x, int64_t n, float scale, int8_t* d_q, pa_stream_t stream);
/* attention_cpu/int8_quant.cpp:15-28 per-row scale, x [rows, dim] */
int pa_batch_quantize_i8(const float* d_x, const float* d_scales, int rows, int dim, int8_t* d_q,
                         pa_stream_t stream);
/* attention_cpu/int8_quant.cpp:30-36 -> *d_out (one float, device) */
int pa_absmax(const float* d_x, int64_t n, float* d_out, pa_stream_t stream);
/* attention_cpu/int8_quant.cpp:59-64 127/(absmax+1e-6) -> *d_out */
int pa_minmax_scale(const float* d_x, int64_t n, float* d_out, pa_stream_t stream);
/* the same, one scale per row of x [rows, dim] -> d_scales[rows] */
int pa_batch_minmax_scale(const float* d_x, int rows, int dim, float* d_scales, pa_stream_t stream);
/* attention_cpu/int8_quant.cpp:38-44  x = (float)q / scale (IEEE division) */
int pa_dequantize_i8(const int8_t* d_q, int64_t n, float scale, float* d_x, pa_stream_t stream);
/* attention_cpu/int8_quant.cpp:46-57 */
int pa_batch_dequantize_i8(const int8_t* d_q, const float* d_scales, int rows, int dim, float* d_x,
                           pa_stream_t stream);

/* ------------------------------------------------------------ int8 matmul */
/* attention_cpu/dnnl_matmul_int8.cpp:7-76:
 *   acc[b,m,n] = sum_k A[b,m,k] * B[b,k,n]              (s8 x s8 -> s32, exact)
 *   C_s8       = sat_s8(rne(act(alpha*acc + bias[n]))),  alpha = scaleA*scaleB/scaleC
 * A [BATCH,M,K], B [BATCH,K,N], C [BATCH,M,N] row-major (format_tag::abc).
 * d_C_s8 and/or d_C_s32 may be NULL (at least one must be given); d_C_s32
 * receives the raw int32 accumulators (the bit-exact parity target).
 * tcgen05 kind::i8 tensor-core kernel; K % 16 == 0 and N % 16 == 0 required.
 * d_workspace: caller-owned scratch for split-K partial tiles, at least
 * pa_gemm_i8_workspace_bytes(BATCH, M, N, K) bytes (0 = this shape never splits), 16-byte aligned.  The
 * library owns no growable scratch (a pointer captured in a CUDA graph must stay valid; streams must not share
 * it): NULL or too small a workspace makes the GEMM run unsplit -- same result, fewer CTAs. */
size_t pa_gemm_i8_workspace_bytes(int BATCH, int M, int N, int K);
int pa_gemm_i8(const int8_t* d_A, const int8_t* d_B, int8_t* d_C_s8, int32_t* d_C_s32, int BATCH,
               int M, int N, int K, float scaleA, float scaleB, float scaleC, const float* d_bias,
               int act, void* d_workspace, size_t workspace_bytes, pa_stream_t stream);

/* Dynamic-quantisation form of the same GEMM, used by the INT8 decoder MLP ("int8_quant ->
 * dnnl_matmul_int8 -> dequant", attention_cpu/README.md:80-86):
 *   C_f32[b,m,n] = act(alpha_row*acc + bias[n]),  alpha_row = b_dequant / d_a_qscales[b*M+m]
 * d_a_qscales [BATCH*M] are the DEVICE-resident per-row quantisation scales of A in the
 * int8_quant.cpp convention (A_s8 = batch_quantize(x, scales, K), int8_quant.cpp:15-28,
 * so x ~ A_s8 / scale); b_dequant is the dequantisation multiplier of the weights
 * (w ~ B_s8 * b_dequant).  No host synchronisation: the scales produced by
 * pa_batch_minmax_scale feed this call directly. */
int pa_gemm_i8_dequant(const int8_t* d_A, const int8_t* d_B, float* d_C_f32, int BATCH, int M, int N,
                       int K, const float* d_a_qscales, float b_dequant, const float* d_bias, int act,
                       void* d_workspace, size_t workspace_bytes,
                       pa_stream_t stream);

/* The same GEMM with the DYNAMIC ROW QUANTISATION of its output fused into the epilogue -- the
 * "fc1 -> int8_quant -> fc2" hand-off of the INT8 decoder MLP (attention_cpu/README.md:80-86) in one kernel:
 *   v[b,m,n]       = act(alpha_row*acc + bias[n])                  (as pa_gemm_i8_dequant; act none / relu)
 *   d_c_qscale[bm] = 127 / (max_n |v| + 1e-6)                      (compute_minmax_scale, int8_quant.cpp:59-64)
 *   C_s8[b,m,n]    = clamp(round_half_away(v * d_c_qscale[bm]))    (batch_quantize, int8_quant.cpp:15-28)
 * bit-identical to pa_gemm_i8_dequant followed by pa_row_quantize_dynamic_i8, without the f32 round trip (the
 * accumulators wait in TMEM while the row maxima cross the grid).  Needs the whole grid resident at once:
 * PA_ERR_UNSUPPORTED for M <= 128, for more CTA pairs than SMs / 2, for shapes that would split K, and for gelu --
 * callers then use the two separate calls.  d_workspace: >= pa_gemm_i8_dynquant_workspace_bytes(BATCH, M, N) bytes,
 * 16-byte aligned, DEDICATED to this function and ZERO-FILLED ONCE by the caller (its first 16 bytes are the
 * self-resetting counters of the grid barrier; one stream at a time per workspace). */
size_t pa_gemm_i8_dynquant_workspace_bytes(int BATCH, int M, int N);
int pa_gemm_i8_dynquant(const int8_t* d_A, const int8_t* d_B, int8_t* d_C_s8, float* d_c_qscale, int BATCH,
                        int M, int N, int K, const float* d_a_qscale, float b_dequant, const float* d_bias,
                        int act, void* d_workspace, size_t workspace_bytes, pa_stream_t stream);

/* ------------------------------------------------- decoder glue (row a14) */
/* decoder/token_embedding.hpp:19-26: d_out[r,:] = E[d_ids[r],:] (ids outside [0,vocab) give
 * zeros; the reference reads out of bounds).  _i8: int8 table dequantised as q/qscale
 * (int8_quant.cpp:38-44). */
int pa_embedding_f32(const float* d_E, const int32_t* d_ids, int rows, int hidden, int vocab,
                     float* d_out, pa_stream_t stream);
int pa_embedding_i8(const int8_t* d_E, float qscale, const int32_t* d_ids, int rows, int hidden,
                    int vocab, float* d_out, pa_stream_t stream);
/* decoder/layer_norm.hpp:20-37: biased variance, inv_std = 1/sqrt(var+eps), y = (x-mean)*inv_std*gamma+beta.
 * d_out may alias d_x. */
int pa_layer_norm_f32(const float* d_x, const float* d_gamma, const float* d_beta, int rows,
                      int hidden, float eps, float* d_out, pa_stream_t stream);
/* decoder/mlp.hpp:23-41, one layer of the float MLP: out[r,n] = act(bias[n] + sum_k x[r,k]*W[k*N+n]),
 * act in {PA_ACT_NONE, PA_ACT_RELU}.  d_out must not alias d_x.  d_workspace: caller-owned scratch for the
 * K-slice partials, >= pa_linear_workspace_bytes(rows, K, N) bytes (NULL / too small: unsliced, same result
 * up to the summation order).
 * Arithmetic: layers with >= 16 rows, N % 4 == 0, K % 4 == 0, 16-byte aligned operands and >= 5e7 MACs run on
 * the tensor cores (tcgen05 kind::tf32 with every operand split into hi + lo: x.W ~= x_hi.W_lo + x_lo.W_hi +
 * x_hi.W_hi, fp32 accumulation; error <= 1e-5 * sum_k |x||W| stated, <= 1.1e-6 measured -- the fp32 SIMT
 * kernels: 1-3e-7); everything else, and everything with PA_LINEAR_TC=0 in the environment, in fp32 FMA
 * arithmetic proper (only the summation order differs from the reference's scalar loop). */
size_t pa_linear_workspace_bytes(int rows, int K, int N);
int pa_linear_f32(const float* d_x, const float* d_W, const float* d_bias, int rows, int K, int N,
                  int act, float* d_out, void* d_workspace, size_t workspace_bytes, pa_stream_t stream);
/* The same layer on weights repacked ONCE (at load) into the tensor-core kernel's own order:
 * [ceil(N/128) feature tiles][ceil(K/32) K blocks][32 k][128 n] f32, zero padded, so that every block the
 * kernel streams is one contiguous 16 KB run (whole DRAM pages instead of 512-byte rows 4*N bytes apart).
 * pa_linear_pack_bytes: size of the packed copy; pa_linear_pack_f32: W [K, N] -> d_W_packed.
 * pa_linear_f32_packed: any rows >= 1, any N; K % 4 == 0 and 16-byte aligned d_x (else PA_ERR_UNSUPPORTED);
 * workspace as pa_linear_workspace_bytes(rows, K, N); same arithmetic and bits as pa_linear_f32 on the
 * tensor-core kernel (PA_LINEAR_TC=1). */
size_t pa_linear_pack_bytes(int K, int N);
int pa_linear_pack_f32(const float* d_W, float* d_W_packed, int K, int N, pa_stream_t stream);
int pa_linear_f32_packed(const float* d_x, const float* d_W_packed, const float* d_bias, int rows, int K,
                         int N, int act, float* d_out, void* d_workspace, size_t workspace_bytes,
                         pa_stream_t stream);
/* logits[r,v] = dot(x[r,:], E[v,:]) against the (tied) embedding table E [vocab, hidden]
 * (the reference reads logits out of the hidden state, cuda_decoder.cu:58). */
int pa_logits_f32(const float* d_x, const float* d_E, int rows, int hidden, int vocab,
                  float* d_logits, pa_stream_t stream);
int pa_logits_i8(const float* d_x, const int8_t* d_E, float qscale, int rows, int hidden, int vocab,
                 float* d_logits, pa_stream_t stream);
/* decoder/cuda_decoder.cu:7-14 sample_from_logits (divide=1: argmax logits/T) and
 * decoder/int8_decoder.cpp:97-104 sample_from_int8_logits (divide=0: argmax logits*T);
 * first maximum wins (std::max_element). */
int pa_argmax_f32(const float* d_logits, int rows, int vocab, float temperature, int divide,
                  int32_t* d_out_ids, pa_stream_t stream);
/* pa_logits_* followed by pa_argmax_f32 in ONE pass over the embedding table: the sampler is folded
 * into the logits kernel as an atomicMax on an order-preserving 64-bit key (first maximum wins).
 * d_best: [rows] u64 scratch, zero-initialised once by the caller, reset by every call.
 * elem_bytes 4: d_E f32; elem_bytes 1: d_E int8 dequantised as q / qscale. */
int pa_logits_argmax(const float* d_x, const void* d_E, int elem_bytes, float qscale, int rows, int hidden,
                     int vocab, float temperature, int divide, float* d_logits,
                     unsigned long long* d_best, int32_t* d_out_ids, pa_stream_t stream);
/* LayerNorm (decoder/layer_norm.hpp:20-37) fused with the dynamic row quantisation that follows it in the INT8
 * decoder (compute_minmax_scale + batch_quantize per row, int8_quant.cpp:59-64, 15-28): d_scales[row], d_q [rows,
 * hidden]; d_out_f32 (optional) also receives the normalised f32 rows.  Bit-identical to pa_layer_norm_f32 followed
 * by pa_row_quantize_dynamic_i8. */
int pa_layer_norm_quantize_i8(const float* d_x, const float* d_gamma, const float* d_beta, int rows, int hidden,
                              float eps, float* d_out_f32, float* d_scales, int8_t* d_q, pa_stream_t stream);
/* compute_minmax_scale + batch_quantize of every row of x [rows, dim] in one kernel
 * (attention_cpu/int8_quant.cpp:59-64, 15-28); bit-identical to pa_batch_minmax_scale followed by
 * pa_batch_quantize_i8. */
int pa_row_quantize_dynamic_i8(const float* d_x, int rows, int dim, float* d_scales, int8_t* d_q,
                               pa_stream_t stream);
/* attention/attention_kernel_utils.cuh:20-35 apply_rotary_embedding for whole batches: d_q and/or d_k
 * [rows, num_heads, head_dim] f32 are rotated pairwise in place with the interleaved table
 * d_rope[token * head_dim + d] = cos, [.. + d + 1] = sin at token = d_positions[row] (apply_on_k = d_k
 * given).  Rows whose position is outside [0, T) are left untouched.  Bit-exact (same operation order). */
int pa_apply_rope_f32(float* d_q, float* d_k, const float* d_rope, const int32_t* d_positions, int rows,
                      int num_heads, int head_dim, int T, pa_stream_t stream);
/* positions[r] += 1 (and ctx_lens[r] += 1 when given): advances the decode step on the device. */
int pa_advance_positions(int32_t* d_positions, int32_t* d_ctx_lens, int rows, pa_stream_t stream);

/* ------------------------------------------------ sampling */
/* attention_cpu/softmax_lut.cpp:203-231 softmax_lut_vec over vocabulary logits [rows, vocab]:
 * p = exp((x - max)/temperature) / (sum + 1e-6). */
int pa_softmax_temperature(const float* d_logits, int rows, int vocab, float temperature, float* d_probs,
                           pa_stream_t stream);
/* attention_cpu/softmax_lut.cpp:60-100 fused_softmax_lut_inplace / softmax_batch_parallel over int32
 * logits [rows, n] with the exp table of build_exp_lut (:11-18, d_lut [resolution], max_x = 10):
 * BIT-EXACT (table index by truncation, sum in the reference's sequential order). */
int pa_softmax_lut_i32(const int32_t* d_logits, int rows, int n, float scale, const float* d_lut,
                       int resolution, float* d_probs, pa_stream_t stream);
/* attention_cpu/softmax_lut.cpp:233-256 apply_topk_topp_filter, in place: rank entries by
 * (prob, index) descending (ties: larger index first), zero those with rank >= top_k (top_k > 0)
 * or whose higher-ranked probability mass is >= top_p (top_p < 1), no renormalisation; then the
 * EOS rule (probs[eos] > eos_thresh zeroes everything else).  Replaces the serial thread-0 top-k of
 * attention/top_k_top_p_filter.cuh:55-111 by a radix select. */
int pa_topk_topp_filter(float* d_probs, int rows, int vocab, int top_k, float top_p, int eos_token_id,
                        float eos_thresh, pa_stream_t stream);
/* Inverse-CDF sample in index order: first index whose inclusive prefix sum of d_probs exceeds
 * d_uniform[row] * total (uniform numbers in [0,1) supplied by the caller; the reference calls the
 * host rand() from device code, top_k_top_p_filter.cuh:109). */
int pa_sample_from_probs(const float* d_probs, int rows, int vocab, const float* d_uniform,
                         int32_t* d_out_ids, pa_stream_t stream);

/* ------------------------------------------------ prefill */
/* The multi-query form the reference's headers promise (q/out [B, H, T, D] and `is_prefill`,
 * attention/attention_config.hpp:8-9,17; causal rule attention/attention_kernel_utils.cuh:70-79)
 * but do not implement: query t of row b attends the cached keys [0, ctx_start[b] + t] of
 * table row (d_beam_ids ? d_beam_ids[b] : b).  d_q, d_out: [B, num_heads, Tq, head_dim] f32.
 * d_ctx_start: [B] tokens cached BEFORE this chunk (NULL = 0: plain prefill); the Tq new
 * tokens' K/V must already be appended.  Same softmax semantics as pa_paged_decode_*. */
size_t pa_prefill_workspace_bytes(int B, int Tq, int num_heads, int head_dim, int num_tiles, int tile_size);
int pa_paged_prefill_f16(const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                         const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                         int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_start, int B,
                         int Tq, int head_dim, int tile_size, float temperature, void* d_workspace,
                         size_t workspace_bytes, pa_stream_t stream);
int pa_paged_prefill_i8(const float* d_q, float* d_out, const int8_t* d_k_pool, const int8_t* d_v_pool,
                        const float* d_k_scales, const float* d_v_scales, const int32_t* d_table,
                        int num_beams, int num_heads, int num_tiles, int total_pages,
                        const int32_t* d_beam_ids, const int32_t* d_ctx_start, int B, int Tq,
                        int head_dim, int tile_size, float temperature, void* d_workspace,
                        size_t workspace_bytes, pa_stream_t stream);

/* Token-major prefill: d_q, d_out are [B, Tq, num_heads, head_dim] f32 -- the layout the decoders' QKV
 * projection leaves the activations in (CUDADecoder.py:75-80 reshapes [B, n, hidden] into heads), so the
 * caller needs no permute + copy either side of the attention.  The strides are folded into the tcgen05
 * kernel's Q loads and O stores; everything else as pa_paged_prefill_f16/_i8.  No workspace.  Returns
 * PA_ERR_UNSUPPORTED where that kernel does not apply (head_dim other than 64 / 128, pages that are not
 * 16 << k tokens, pools not 128-byte aligned, d_q == d_out): permute and call the [B, H, Tq, D] entry then. */
int pa_paged_prefill_f16_tokmajor(const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                                  const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                                  int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_start,
                                  int B, int Tq, int head_dim, int tile_size, float temperature,
                                  pa_stream_t stream);
int pa_paged_prefill_i8_tokmajor(const float* d_q, float* d_out, const int8_t* d_k_pool, const int8_t* d_v_pool,
                                 const float* d_k_scales, const float* d_v_scales, const int32_t* d_table,
                                 int num_beams, int num_heads, int num_tiles, int total_pages,
                                 const int32_t* d_beam_ids, const int32_t* d_ctx_start, int B, int Tq,
                                 int head_dim, int tile_size, float temperature, pa_stream_t stream);

/* --------------------------------- inter-GPU split-KV exchange over peer memory */
/* Exchange buffers: one per rank, pa_splitkv_exchange_bytes(world, rows, head_dim) bytes
 * (uint4 [2 parity][world][rows][head_dim/2 + 1] flag-in-data packets, csrc/xchg.cuh).
 * pa_p2p_alloc: cudaMalloc'd, zero-filled exchange buffer + its 64-byte CUDA IPC handle;
 * pa_p2p_open: map a peer process's buffer (NVLink P2P); pa_p2p_close / pa_p2p_free. */
size_t pa_splitkv_exchange_bytes(int world, int rows, int head_dim);
int pa_p2p_alloc(size_t bytes, void** d_ptr, unsigned char* handle64);
int pa_p2p_open(const unsigned char* handle64, void** d_peer_ptr);
int pa_p2p_close(void* d_peer_ptr);
int pa_p2p_free(void* d_ptr);
/* Decode + row merge + exchange + combine in ONE launch: the streaming kernel over this rank's
 * pages; the warp that finishes the last chunk of a row merges it and stores it into slot `rank`
 * of every peer's exchange buffer (d_peer_bufs: device array [world] of mapped buffer pointers,
 * this rank's own buffer at index `rank`) as 16-byte packets that carry their own epoch flag --
 * sending never blocks; at the end of the same kernel every row's `world` partials are received
 * from the local buffer and LSE-combined into d_out [B, H, D] (identical on every rank).
 * d_epochs [2*B*H] u32: per-row step counters (first half) and self-resetting row-completion counters (second
 * half) in device memory, zero-initialised ONCE by the caller (the launch is CUDA-graph replayable); *d_status is set to 1, the row written as NaN and its epoch NOT advanced if a
 * peer does not arrive within ~2 s.  Exchange buffers must be sized for rows = B*num_heads. */
int pa_paged_decode_f16_splitkv(const float* d_q, float* d_out, const void* d_k_pool,
                                const void* d_v_pool, const int32_t* d_table, int num_beams,
                                int num_heads, int num_tiles, int total_pages,
                                const int32_t* d_beam_ids, const int32_t* d_ctx_lens, int B, int T,
                                int head_dim, int tile_size, float temperature, const float* d_rope,
                                float* d_lse_out, void* d_workspace, size_t workspace_bytes,
                                void* const* d_peer_bufs, int rank, int world, uint32_t* d_epochs,
                                int* d_status, pa_stream_t stream);
int pa_paged_decode_i8_splitkv(const float* d_q, float* d_out, const int8_t* d_k_pool,
                               const int8_t* d_v_pool, const float* d_k_scales, const float* d_v_scales,
                               const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                               int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_lens,
                               int B, int T, int head_dim, int tile_size, float temperature,
                               const float* d_rope, float* d_lse_out, void* d_workspace,
                               size_t workspace_bytes, void* const* d_peer_bufs, int rank, int world,
                               uint32_t* d_epochs, int* d_status, pa_stream_t stream);
/* The same packet exchange as a stand-alone kernel after pa_paged_decode_*_partial: phase 1 sends
 * every row (never blocks), phase 2 receives and combines.  head_dim in {64, 128}, world <= 32. */
int pa_splitkv_exchange_combine(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                                void* const* d_peer_bufs, int rank, int world, int rows,
                                int head_dim, uint32_t* d_epochs, float* d_out, float* d_lse_out,
                                int* d_status, pa_stream_t stream);

/* The two halves of pa_splitkv_exchange_combine on their own: _send stores this rank's rows into every peer's
 * buffer and never waits; _recv polls until every rank's packets of the current step are in this rank's buffer,
 * combines and advances the epochs. */
int pa_splitkv_exchange_send(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                             void* const* d_peer_bufs, int rank, int world, int rows, int head_dim,
                             const uint32_t* d_epochs, pa_stream_t stream);
int pa_splitkv_exchange_recv(void* const* d_peer_bufs, int rank, int world, int rows, int head_dim,
                             uint32_t* d_epochs, float* d_out, float* d_lse_out, int* d_status,
                             pa_stream_t stream);

/* --------------------------------- NCCL form of the exchange (the north star's baseline) */
/* NCCL is loaded at run time (dlopen "libnccl.so.2"); PA_ERR_UNSUPPORTED when it is absent.
 * pa_nccl_unique_id: rank 0 creates the 128-byte ncclUniqueId, the application distributes it;
 * pa_nccl_init: ncclCommInitRank on the current device; pa_nccl_destroy.
 * pa_nccl_allgather_combine: ncclAllGather of this rank's (m [rows], l [rows], O [rows, D]) into
 * d_gather_ws (>= pa_nccl_gather_bytes) in the [n_parts][rows] layout of pa_lse_combine, then the
 * combine kernel, all on `stream`. */
int pa_nccl_unique_id(unsigned char* id128);
int pa_nccl_init(const unsigned char* id128, int rank, int world, void** comm);
int pa_nccl_destroy(void* comm);
size_t pa_nccl_gather_bytes(int world, int rows, int head_dim);
int pa_nccl_allgather_combine(void* comm, int world, const float* d_part_m, const float* d_part_l,
                              const float* d_part_o, int rows, int head_dim, void* d_gather_ws,
                              size_t gather_bytes, float* d_out, float* d_lse_out, pa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PA_B200_H_ */
