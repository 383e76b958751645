// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" trampolines over the reference's OWN translation units that
// compile as shipped: attention_cpu/int8_quant.cpp,
// attention_cpu/softmax_lut.cpp (-fpermissive) and kv_cache/kv_tile_cache_cpu.cpp.
// The reference sources are compiled where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libref_cpu.so; nothing is copied into this
// repo.  This file only adapts std::vector signatures to plain pointers so
// Python (ctypes) can call them.  Used to pin oracle_cpu.c and, in bench.py's
// cpu_baseline / --impl reference legs, as the timed reference CPU code.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "attention_cpu/int8_quant.hpp"
#include "attention_cpu/softmax_lut.hpp"
#include "kv_cache/kv_tile_cache_cpu.hpp"

#define REF_API extern "C" __attribute__((visibility("default")))

// ---- attention_cpu/int8_quant.cpp -------------------------------------
REF_API void ref_quantize_to_int8(const float* x, int64_t n, float scale, int8_t* out) {
    std::vector<float> in(x, x + n);
    auto q = quantize_to_int8(in, scale);
    std::memcpy(out, q.data(), (size_t)n);
}
REF_API void ref_batch_quantize(const float* x, const float* scales, int rows, int dim, int8_t* out) {
    std::vector<float> in(x, x + (size_t)rows * dim), sc(scales, scales + rows);
    auto q = batch_quantize(in, sc, dim);
    std::memcpy(out, q.data(), (size_t)rows * dim);
}
REF_API float ref_compute_absmax(const float* x, int64_t n) {
    return compute_absmax(std::vector<float>(x, x + n));
}
REF_API void ref_dequantize_from_int8(const int8_t* q, int64_t n, float scale, float* out) {
    auto f = dequantize_from_int8(std::vector<int8_t>(q, q + n), scale);
    std::memcpy(out, f.data(), sizeof(float) * (size_t)n);
}
REF_API void ref_batch_dequantize(const int8_t* q, const float* scales, int rows, int dim, float* out) {
    std::vector<int8_t> in(q, q + (size_t)rows * dim);
    std::vector<float> sc(scales, scales + rows);
    auto f = batch_dequantize(in, sc, dim);
    std::memcpy(out, f.data(), sizeof(float) * (size_t)rows * dim);
}
REF_API float ref_compute_minmax_scale(const float* x, int64_t n) {
    return compute_minmax_scale(std::vector<float>(x, x + n));
}

// ---- attention_cpu/softmax_lut.cpp -------------------------------------
REF_API void ref_build_exp_lut(int resolution, float max_x, float* lut) {
    auto v = build_exp_lut(resolution, max_x);
    std::memcpy(lut, v.data(), sizeof(float) * v.size());
}
// n must be a multiple of 8 (the reference's 8-wide store overruns otherwise).
REF_API void ref_softmax_lut(const int32_t* logits, int64_t n, float scale, const float* lut,
                             int resolution, float* out) {
    std::vector<int32_t> in(logits, logits + n);
    std::vector<float> l(lut, lut + resolution);
    auto p = softmax_lut(in, scale, l);
    std::memcpy(out, p.data(), sizeof(float) * (size_t)n);
}
REF_API void ref_fused_softmax_lut_inplace(const int32_t* logits, int64_t n, float scale,
                                           const float* lut, int resolution, float* out) {
    std::vector<int32_t> in(logits, logits + n);
    std::vector<float> l(lut, lut + resolution), o;
    fused_softmax_lut_inplace(in, scale, l, o);
    std::memcpy(out, o.data(), sizeof(float) * (size_t)n);
}
REF_API void ref_softmax_batch_parallel(const int32_t* logits, int rows, int64_t n, float scale,
                                        const float* lut, int resolution, float* out) {
    std::vector<std::vector<int32_t>> batch(rows);
    for (int r = 0; r < rows; ++r) batch[r].assign(logits + (size_t)r * n, logits + (size_t)(r + 1) * n);
    std::vector<float> l(lut, lut + resolution);
    std::vector<std::vector<float>> probs;
    softmax_batch_parallel(batch, scale, l, probs);
    for (int r = 0; r < rows; ++r) std::memcpy(out + (size_t)r * n, probs[r].data(), sizeof(float) * (size_t)n);
}
// len must be a multiple of 8.
REF_API void ref_softmax_lut_vec(const float* scores, int len, float temperature, float* out) {
    std::vector<float> s(scores, scores + len);
    softmax_lut_vec(s.data(), len, temperature, out, nullptr);
}
REF_API void ref_softmax_lut_tile(const float* scores, int len, float temperature, float* out) {
    std::vector<float> s(scores, scores + len);
    softmax_lut_tile(s.data(), len, temperature, out);
}
REF_API void ref_apply_topk_topp_filter(float* probs, int len, int top_k, float top_p,
                                        int eos_token_id, float eos_thresh) {
    apply_topk_topp_filter(probs, len, top_k, top_p, eos_token_id, eos_thresh);
}
REF_API void ref_apply_top_k(float* probs, int len, int k) {
    std::vector<float> p(probs, probs + len);
    apply_top_k(p, k);
    std::memcpy(probs, p.data(), sizeof(float) * (size_t)len);
}
REF_API void ref_apply_top_p(float* probs, int len, float pth) {
    std::vector<float> p(probs, probs + len);
    apply_top_p(p, pth);
    std::memcpy(probs, p.data(), sizeof(float) * (size_t)len);
}

// ---- kv_cache/kv_tile_cache_cpu.cpp ------------------------------------
// tile_size counts ELEMENTS per stored tile (kv_tile_cache_cpu.cpp:22-23).
REF_API void* ref_kvcpu_f32_new(int max_size, int tile_size) { return new KVTileCacheCPU<float>(max_size, tile_size); }
REF_API void ref_kvcpu_f32_free(void* h) { delete static_cast<KVTileCacheCPU<float>*>(h); }
REF_API void ref_kvcpu_f32_put(void* h, int b, int hd, int t, const float* data) {
    static_cast<KVTileCacheCPU<float>*>(h)->put(b, hd, t, data);
}
REF_API const float* ref_kvcpu_f32_get(void* h, int b, int hd, int t) {
    return static_cast<KVTileCacheCPU<float>*>(h)->get(b, hd, t);
}
REF_API int ref_kvcpu_f32_save(void* h, const char* path) {
    try { static_cast<KVTileCacheCPU<float>*>(h)->save(path); return 0; } catch (...) { return 1; }
}
REF_API int ref_kvcpu_f32_load(void* h, const char* path) {
    try { static_cast<KVTileCacheCPU<float>*>(h)->load(path); return 0; } catch (...) { return 1; }
}
REF_API void* ref_kvcpu_i8_new(int max_size, int tile_size) { return new KVTileCacheCPU<int8_t>(max_size, tile_size); }
REF_API void ref_kvcpu_i8_free(void* h) { delete static_cast<KVTileCacheCPU<int8_t>*>(h); }
REF_API void ref_kvcpu_i8_put(void* h, int b, int hd, int t, const int8_t* data) {
    static_cast<KVTileCacheCPU<int8_t>*>(h)->put(b, hd, t, data);
}
REF_API const int8_t* ref_kvcpu_i8_get(void* h, int b, int hd, int t) {
    return static_cast<KVTileCacheCPU<int8_t>*>(h)->get(b, hd, t);
}
