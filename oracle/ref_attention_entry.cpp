// ref_attention_entry.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" trampolines over (a) the reference's cpu_paged_attention_forward<float>
// (attention_cpu/cpu_attention_kernel.cpp:36-129, compiled from the patched temporary copy made by
// oracle/build_ref_attention.py) and (b) the decoder's header-only LayerNorm<float>, MLP<float>,
// TokenEmbedding<float> (decoder/layer_norm.hpp, mlp.hpp, token_embedding.hpp, compiled unmodified).
// Nothing here computes: the functions only move caller arrays into the reference's own containers
// (its KVTileCacheCPU<float> tile store, its weight files) and call the reference code.
#include <omp.h>

#include <cstring>

#include "cpu_attention_kernel.hpp"  // the patched copy (-I <tmp>/attention_cpu comes first)
#include "decoder/layer_norm.hpp"
#include "decoder/mlp.hpp"
#include "decoder/token_embedding.hpp"

#define REF_API extern "C" __attribute__((visibility("default")))

// Paged pools + dense page table -> the reference's tile store, then its forward.
//   k_pool / v_pool: [total_pages][tile_size][D] f32; table: [num_beams][H][num_tiles] int32, -1 / out of
//   range = unmapped (tile absent from the store -> get() returns nullptr -> the kernel skips it, :73).
//   probs_out / logits_out: optional [B*H][T] (CPUAttentionOutput::attention_weights / logits, hpp:34-39).
// Runs on ONE thread: KVTileCacheCPU::get mutates its LRU list under a shared lock
// (kv_tile_cache_cpu.cpp:70-80), which is a data race under the kernel's own `omp parallel for`.
REF_API int ref_cpu_paged_attention_f32(const float* q, float* out, const float* k_pool, const float* v_pool,
                                        const int32_t* table, int num_beams, int num_tiles, int total_pages,
                                        const int32_t* beam_ids, int B, int H, int T, int D, int tile_size,
                                        float temperature, const float* rope, int top_k, float top_p, int causal,
                                        float* probs_out, float* logits_out) {
    try {
        const int elems = tile_size * D;
        KVTileCacheCPU<float> store(2 * num_beams * H * num_tiles + 2, elems);
        for (int b = 0; b < num_beams; ++b)
            for (int h = 0; h < H; ++h)
                for (int t = 0; t < num_tiles; ++t) {
                    const int page = table[((int64_t)b * H + h) * num_tiles + t];
                    if (page < 0 || page >= total_pages) continue;
                    store.put(b, h, 2 * t, k_pool + (int64_t)page * elems);
                    store.put(b, h, 2 * t + 1, v_pool + (int64_t)page * elems);
                }
        refshim::KVTileStore4<float> store4{&store};
        std::vector<int> beams;
        if (beam_ids) beams.assign(beam_ids, beam_ids + B);
        CPUAttentionInput<float> in;
        in.q = q;
        in.beam_ids = beam_ids ? &beams : nullptr;
        in.rotary_emb = rope;
        in.B = B; in.H = H; in.T = T; in.D = D;
        in.tile_size = tile_size;
        in.temperature = temperature;
        in.top_k = top_k;
        in.top_p = top_p;
        in.causal = causal != 0;
        in.kv_cache = &store4;
        std::vector<std::vector<float>> weights((size_t)B * H), logits((size_t)B * H);
        CPUAttentionOutput<float> o;
        o.out = out;
        o.attention_weights = probs_out ? &weights : nullptr;
        o.logits = logits_out ? &logits : nullptr;
        const int saved = omp_get_max_threads();
        omp_set_num_threads(1);
        cpu_paged_attention_forward<float>(in, o);
        omp_set_num_threads(saved);
        for (size_t r = 0; r < (size_t)B * H; ++r) {
            if (probs_out) std::memcpy(probs_out + r * T, weights[r].data(), sizeof(float) * (size_t)T);
            if (logits_out) std::memcpy(logits_out + r * T, logits[r].data(), sizeof(float) * (size_t)T);
        }
        return 0;
    } catch (...) {
        return 1;
    }
}

// decoder/layer_norm.hpp:20-37.  weights_path: gamma then beta, raw f32 (layer_norm.hpp:13-18).
REF_API int ref_layer_norm_f32(const char* weights_path, int hidden, float eps, const float* in, float* out, int rows) {
    try {
        LayerNorm<float> ln(hidden, eps);
        if (weights_path) ln.load_weights(weights_path);
        ln.forward(in, out, rows);
        return 0;
    } catch (...) {
        return 1;
    }
}

// decoder/mlp.hpp:23-41.  weights_path: fc1_w [hidden][inter], fc1_b, fc2_w [inter][hidden], fc2_b (mlp.hpp:14-21).
REF_API int ref_mlp_f32(const char* weights_path, int hidden, int inter, const float* in, float* out, int rows) {
    try {
        MLP<float> mlp(hidden, inter);
        if (weights_path) mlp.load_weights(weights_path);
        mlp.forward(in, out, rows);
        return 0;
    } catch (...) {
        return 1;
    }
}

// decoder/token_embedding.hpp:19-26.
REF_API int ref_token_embedding_f32(const char* weights_path, int vocab, int hidden, const int32_t* ids, int n, float* out) {
    try {
        TokenEmbedding<float> emb(vocab, hidden);
        if (weights_path) emb.load_weights(weights_path);
        std::vector<int> in(ids, ids + n);
        std::vector<float> e;
        emb.forward(in, e);
        std::memcpy(out, e.data(), sizeof(float) * e.size());
        return 0;
    } catch (...) {
        return 1;
    }
}
