// integration/attention_cuda_b200.cpp -- the reference-side binding: AttentionCUDA::forward over libpa_b200.so.
//
// Drop-in replacement for attention/attention_cuda.cu (the dtype dispatch, :41-95) and
// attention/attention_tile_launcher.hpp:35-90 (the only place that launches the two kernels) of the reference.
// Same 17-argument static signature as attention/attention_config.hpp:5-26.  It needs the eight accessors of
// integration/ref_accessors.hpp on KVTileCache / PageTable and nothing else from the reference.
//
// Compiles against the mirrored declarations (tests/test_integration_stub.py: g++ -c, then linked against
// libpa_b200.so so every pa_* symbol used here resolves).  In the reference tree: replace the #include below by
// "../kv_cache/kv_tile_cache.hpp" and add this file to setup.py / CMakeLists.txt in place of attention_cuda.cu.
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <type_traits>

#include "pa_b200.h"
#include "ref_accessors.hpp"  // in the reference tree: "../kv_cache/kv_tile_cache.hpp" (+ the accessors)

// half pools: KVTileCache<__half> in the reference (kv_tile_cache.cpp:128); any 2-byte element type binds here
struct pa_half_t { uint16_t bits; };

namespace {

// Decode scratch: owned by the caller of the C-ABI (this translation unit), one buffer per cache object and
// stream would be the production choice; the reference runs everything on the default stream from one thread
// (bindings.cpp:8-15 holds the GIL), so one grow-only buffer matches its threading model.
struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
};
Scratch g_scratch;

extern "C" int cudaMalloc(void**, size_t);  // declared here so the stub compiles without the CUDA headers
extern "C" int cudaFree(void*);

void* scratch(size_t need) {
    if (need > g_scratch.bytes) {
        if (g_scratch.ptr) cudaFree(g_scratch.ptr);
        g_scratch.ptr = nullptr;
        g_scratch.bytes = 0;
        if (cudaMalloc(&g_scratch.ptr, need) != 0) throw std::runtime_error("pa_b200: scratch allocation failed");
        g_scratch.bytes = need;
    }
    return g_scratch.ptr;
}

void check(int st) {
    if (st != PA_OK) throw std::runtime_error(pa_error_string(st));  // -> pybind11 -> RuntimeError (SURVEY 8b "Errors")
}

}  // namespace

// attention/attention_tile_launcher.hpp:35-90, for any pool element type the cache is instantiated with.
template <typename T>
void attention_forward_paged(const float* q, float* out, int B, int H, int D, int T_len, const int* beam_ids,
                             KVTileCache<T>* kv, const float* rotary_emb, bool use_overlap, float temperature,
                             int top_k, float top_p, float* rerank_scores) {
    if (!kv) throw std::runtime_error("AttentionCUDA::forward: kv_cache is required (paged path)");
    const PageTable& pt = kv->page_table();
    const size_t ws_bytes = pa_decode_workspace_bytes(B, H, D, pt.num_tiles(), kv->tile_size());
    void* ws = scratch(ws_bytes);
    if (top_k > 0 || top_p < 1.0f) {
        // the reference's in-attention filter (softmax_lut.cpp:233-256): the explicit three-stage kernels
        const size_t fbytes = pa_attention_filtered_workspace_bytes(B, H, T_len);
        void* fws = scratch(fbytes > ws_bytes ? fbytes : ws_bytes);
        check(pa_paged_attention_filtered(q, out, kv->key_buffer(), kv->value_buffer(), nullptr, nullptr,
                                          std::is_same<T, float>::value ? 2 : 0, pt.device_data(), pt.num_beams(), H,
                                          pt.num_tiles(), kv->total_pages(), beam_ids, nullptr, B, T_len, D,
                                          kv->tile_size(), temperature, rotary_emb, top_k, top_p, nullptr, nullptr, fws,
                                          fbytes, nullptr));
        return;
    }
    int st;
    if (std::is_same<T, float>::value) {  // KVTileCache<float>: the instantiation AttentionCUDA::forward takes
        auto fn = use_overlap ? pa_paged_decode_f32_overlap : pa_paged_decode_f32;
        st = fn(q, out, reinterpret_cast<const float*>(kv->key_buffer()), reinterpret_cast<const float*>(kv->value_buffer()),
                pt.device_data(), pt.num_beams(), H, pt.num_tiles(), kv->total_pages(), beam_ids, nullptr, B, T_len, D,
                kv->tile_size(), temperature, rotary_emb, rerank_scores, ws, ws_bytes, nullptr);
    } else {                              // KVTileCache<half>
        static_assert(std::is_same<T, float>::value || sizeof(T) == 2, "fp32 or fp16 pools");
        auto fn = use_overlap ? pa_paged_decode_f16_overlap : pa_paged_decode_f16;
        st = fn(q, out, kv->key_buffer(), kv->value_buffer(), pt.device_data(), pt.num_beams(), H, pt.num_tiles(),
                kv->total_pages(), beam_ids, nullptr, B, T_len, D, kv->tile_size(), temperature, rotary_emb,
                rerank_scores, ws, ws_bytes, nullptr);
    }
    check(st);
}

// attention/attention_config.hpp:5-26 -- the class the decoder block calls (decoder_block.hpp:45-58).
class AttentionCUDA {
public:
    static void forward(const float* q, float* out, int B, int H, int D, int T, const int* beam_ids = nullptr,
                        KVTileCache<float>* kv_cache = nullptr, const float* rotary_emb = nullptr, bool is_prefill = true,
                        bool use_fp16 = false, bool use_overlap = false, float temperature = 1.0f, int top_k = 1,
                        float top_p = 1.0f, float* rerank_scores = nullptr, bool debug = false);
};

void AttentionCUDA::forward(const float* q, float* out, int B, int H, int D, int T, const int* beam_ids,
                            KVTileCache<float>* kv_cache, const float* rotary_emb, bool is_prefill, bool use_fp16,
                            bool use_overlap, float temperature, int top_k, float top_p, float* rerank_scores, bool debug) {
    (void)is_prefill;  // decode form: q / out [B, H, D] (the kernels' indexing, ...fused.cu:40)
    (void)use_fp16;    // the pool's element type decides, not a flag
    (void)debug;
    // attention_tile_launcher.hpp:48 defaults top_k to 1, which with the kernel's positional filter (...fused.cu:77)
    // would zero every key but token 0; the CPU kernel's default is "off" (cpu_attention_kernel.hpp:20).  1 = off here.
    attention_forward_paged<float>(q, out, B, H, D, T, beam_ids, kv_cache, rotary_emb, use_overlap, temperature,
                                   top_k <= 1 ? 0 : top_k, top_p, rerank_scores);
}

// KVTileCache<half> callers (the second instantiation, kv_tile_cache.cpp:128)
template void attention_forward_paged<pa_half_t>(const float*, float*, int, int, int, int, const int*, KVTileCache<pa_half_t>*,
                                                 const float*, bool, float, int, float, float*);

// kv_cache/kv_tile_cache.hpp:29-34 `get_write_ptr` + row write -> one batched append.
void kv_append_rows(KVTileCache<float>* kv, const float* new_k, const float* new_v, const int* d_row_beams,
                    const int* d_positions, int rows, int H) {
    const PageTable& pt = kv->page_table();
    check(pa_kv_append_f32(const_cast<float*>(kv->key_buffer()), const_cast<float*>(kv->value_buffer()), pt.device_data(),
                           pt.num_beams(), H, pt.num_tiles(), kv->total_pages(), kv->tile_size(), kv->head_dim(), new_k,
                           new_v, d_row_beams, d_positions, rows, nullptr));
}

// attention_cpu/dnnl_matmul_int8.hpp:6-14 -> the tcgen05 GEMM; `false` on any failure, as the reference (cpp:72-75).
bool dnnl_matmul_int8_b200(const int8_t* d_A, const int8_t* d_B, int8_t* d_C, int BATCH, int M, int N, int K, float scaleA,
                           float scaleB, float scaleC, const float* d_bias, int act, void* d_ws, size_t ws_bytes) {
    return pa_gemm_i8(d_A, d_B, d_C, nullptr, BATCH, M, N, K, scaleA, scaleB, scaleC, d_bias, act, d_ws, ws_bytes, nullptr) == PA_OK;
}
