// decoder_ops.cu -- the small device ops that keep CUDADecoder / INT8Decoder.generate on the
// GPU between two attention calls ("minimal GPU versions only to keep
// generate end-to-end"): token embedding, LayerNorm, the fp32 MLP linear layers, tied-embedding
// logits and greedy argmax.  All are HBM-bound streaming kernels at decode batch sizes (a few
// rows against a weight matrix read exactly once).
//
//   decoder/token_embedding.hpp:19-26   -> embedding_kernel
//   decoder/layer_norm.hpp:20-37        -> layer_norm_kernel
//   decoder/mlp.hpp:23-41 (float)       -> linear_kn_kernel (W [K,N] row-major, j*N+i)
//   decoder/cuda_decoder.cu:7-14, decoder/int8_decoder.cpp:97-104 -> argmax_kernel
#include "pa_common.cuh"

namespace pa {

constexpr int kRowChunk = 8;  // activation rows processed per pass over the weights

// ---- embedding --------------------------------------------------------------------------
template <typename WT>
__global__ void embedding_kernel(const WT* __restrict__ E, const int32_t* __restrict__ ids, int rows, int hidden,
                                 int vocab, float qscale, float* __restrict__ out) {
    const int r = blockIdx.x;
    int id = ids[r];
    const bool ok = id >= 0 && id < vocab;
    const WT* src = E + (int64_t)(ok ? id : 0) * hidden;
    for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
        float v = 0.f;
        if (ok) {
            if (sizeof(WT) == 1) v = __fdiv_rn((float)src[j], qscale);  // int8_quant.cpp:41 dequantise
            else v = (float)src[j];
        }
        out[(int64_t)r * hidden + j] = v;
    }
}

// ---- LayerNorm: one warp per row, two passes (mean, then biased variance) ----------------
__global__ void layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, int rows, int hidden, float eps,
                                  float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float* xr = x + (int64_t)warp * hidden;
    float s = 0.f;
    for (int j = lane; j < hidden; j += 32) s += xr[j];
    const float mean = warp_sum(s) / (float)hidden;
    float v = 0.f;
    for (int j = lane; j < hidden; j += 32) {
        const float d = xr[j] - mean;
        v = fmaf(d, d, v);
    }
    const float var = warp_sum(v) / (float)hidden;
    const float inv_std = (float)(1.0 / sqrt((double)(var + eps)));  // layer_norm.hpp:33
    float* o = out + (int64_t)warp * hidden;
    for (int j = lane; j < hidden; j += 32) o[j] = (xr[j] - mean) * inv_std * gamma[j] + beta[j];
}

// ---- linear, W [K, N] row-major: out[r, n] = act(bias[n] + sum_k x[r,k] W[k,n]) ----------
// Block = 8 warps x 32 columns: warp w streams rows k = w, w+8, ... of a 128-byte wide column
// strip (one coalesced line per k), keeps up to kRowChunk accumulators per lane, and the 8
// warps are summed through shared memory.  grid.x = N/32 strips, grid.y = row chunks.
__global__ void __launch_bounds__(256) linear_kn_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                        const float* __restrict__ bias, int rows, int K, int N,
                                                        int act, float* __restrict__ out) {
    __shared__ float red[8][kRowChunk][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 32 + lane;
    const int r0 = blockIdx.y * kRowChunk;
    const int nr = min(kRowChunk, rows - r0);
    float acc[kRowChunk];
#pragma unroll
    for (int r = 0; r < kRowChunk; ++r) acc[r] = 0.f;
    if (n < N) {
        const float* xr = x + (int64_t)r0 * K;
#pragma unroll 4
        for (int k = warp; k < K; k += 8) {
            const float w = __ldcs(W + (int64_t)k * N + n);  // streamed once
#pragma unroll
            for (int r = 0; r < kRowChunk; ++r)
                if (r < nr) acc[r] = fmaf(__ldg(xr + (int64_t)r * K + k), w, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kRowChunk; ++r) red[warp][r][lane] = acc[r];
    __syncthreads();
    // 256 threads finish kRowChunk x 32 outputs
    const int rr = threadIdx.x >> 5, cc = threadIdx.x & 31;
    const int nn = blockIdx.x * 32 + cc;
    if (rr < nr && nn < N) {
        float s = bias ? bias[nn] : 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][rr][cc];
        if (act == PA_ACT_RELU) s = fmaxf(s, 0.f);
        out[(int64_t)(r0 + rr) * N + nn] = s;
    }
}

// ---- logits against the tied embedding E [vocab, hidden]: one warp per vocab row ----------
template <typename WT>
__global__ void __launch_bounds__(256) logits_kernel(const float* __restrict__ x, const WT* __restrict__ E, int rows,
                                                     int hidden, int vocab, float qscale, float* __restrict__ logits) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= vocab) return;
    const WT* e = E + (int64_t)v * hidden;
    for (int r0 = 0; r0 < rows; r0 += kRowChunk) {
        const int nr = min(kRowChunk, rows - r0);
        float acc[kRowChunk];
#pragma unroll
        for (int r = 0; r < kRowChunk; ++r) acc[r] = 0.f;
        for (int j = lane; j < hidden; j += 32) {
            const float w = (float)e[j];
#pragma unroll
            for (int r = 0; r < kRowChunk; ++r)
                if (r < nr) acc[r] = fmaf(__ldg(x + (int64_t)(r0 + r) * hidden + j), w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kRowChunk; ++r) {
            if (r < nr) {
                float s = warp_sum(acc[r]);
                if (sizeof(WT) == 1) s = __fdiv_rn(s, qscale);
                if (lane == 0) logits[(int64_t)(r0 + r) * vocab + v] = s;
            }
        }
    }
}

// ---- greedy sampling: argmax_i(logit_i / T) (cuda_decoder.cu:7-14) or (logit_i * T)
// (int8_decoder.cpp:97-104); std::max_element semantics = FIRST maximum. --------------------
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, int vocab, float t,
                                                      int divide, int32_t* __restrict__ out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const float* lr = logits + (int64_t)blockIdx.x * vocab;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    auto better = [](float v, int i, float bv, int bidx) { return v > bv || (v == bv && i < bidx); };
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        const float v = divide ? __fdiv_rn(lr[i], t) : __fmul_rn(lr[i], t);
        if (better(v, i, best, bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = best; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? sv[lane] : -INFINITY;
        bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) out[blockIdx.x] = (bi == 0x7fffffff) ? 0 : bi;
    }
}

// positions[r] += 1 and ids := next ids (keeps the decode step free of host work so it can be
// captured in a CUDA graph)
__global__ void advance_kernel(int32_t* __restrict__ positions, int32_t* __restrict__ ctx_lens, int rows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) {
        positions[r] += 1;
        if (ctx_lens) ctx_lens[r] += 1;
    }
}

}  // namespace pa

using namespace pa;

PA_API int pa_embedding_f32(const float* d_E, const int32_t* d_ids, int rows, int hidden, int vocab, float* d_out,
                            pa_stream_t stream) {
    PA_CHECK_ARG(d_E && d_ids && d_out && rows >= 0 && hidden > 0 && vocab > 0);
    if (rows == 0) return PA_OK;
    embedding_kernel<float><<<rows, 256, 0, as_stream(stream)>>>(d_E, d_ids, rows, hidden, vocab, 1.f, d_out);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_embedding_i8(const int8_t* d_E, float qscale, const int32_t* d_ids, int rows, int hidden, int vocab,
                           float* d_out, pa_stream_t stream) {
    PA_CHECK_ARG(d_E && d_ids && d_out && rows >= 0 && hidden > 0 && vocab > 0 && qscale != 0.f);
    if (rows == 0) return PA_OK;
    embedding_kernel<int8_t><<<rows, 256, 0, as_stream(stream)>>>(d_E, d_ids, rows, hidden, vocab, qscale, d_out);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_layer_norm_f32(const float* d_x, const float* d_gamma, const float* d_beta, int rows, int hidden,
                             float eps, float* d_out, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_gamma && d_beta && d_out && rows >= 0 && hidden > 0);
    if (rows == 0) return PA_OK;
    const int wpb = 4;
    layer_norm_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, as_stream(stream)>>>(d_x, d_gamma, d_beta, rows, hidden,
                                                                                 eps, d_out);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_linear_f32(const float* d_x, const float* d_W, const float* d_bias, int rows, int K, int N, int act,
                         float* d_out, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_W && d_out && rows >= 0 && K > 0 && N > 0);
    PA_CHECK_ARG(act == PA_ACT_NONE || act == PA_ACT_RELU);
    PA_CHECK_ARG(d_x != d_out);
    if (rows == 0) return PA_OK;
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((rows + kRowChunk - 1) / kRowChunk));
    linear_kn_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_x, d_W, d_bias, rows, K, N, act, d_out);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_logits_f32(const float* d_x, const float* d_E, int rows, int hidden, int vocab, float* d_logits,
                         pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_E && d_logits && rows >= 0 && hidden > 0 && vocab > 0);
    if (rows == 0) return PA_OK;
    logits_kernel<float><<<(vocab + 7) / 8, 256, 0, as_stream(stream)>>>(d_x, d_E, rows, hidden, vocab, 1.f, d_logits);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_logits_i8(const float* d_x, const int8_t* d_E, float qscale, int rows, int hidden, int vocab,
                        float* d_logits, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_E && d_logits && rows >= 0 && hidden > 0 && vocab > 0 && qscale != 0.f);
    if (rows == 0) return PA_OK;
    logits_kernel<int8_t><<<(vocab + 7) / 8, 256, 0, as_stream(stream)>>>(d_x, d_E, rows, hidden, vocab, qscale,
                                                                          d_logits);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_argmax_f32(const float* d_logits, int rows, int vocab, float temperature, int divide,
                         int32_t* d_out_ids, pa_stream_t stream) {
    PA_CHECK_ARG(d_logits && d_out_ids && rows >= 0 && vocab > 0);
    PA_CHECK_ARG(!divide || temperature != 0.f);
    if (rows == 0) return PA_OK;
    argmax_kernel<<<rows, 1024, 0, as_stream(stream)>>>(d_logits, vocab, temperature, divide, d_out_ids);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_advance_positions(int32_t* d_positions, int32_t* d_ctx_lens, int rows, pa_stream_t stream) {
    PA_CHECK_ARG(d_positions && rows >= 0);
    if (rows == 0) return PA_OK;
    advance_kernel<<<(rows + 127) / 128, 128, 0, as_stream(stream)>>>(d_positions, d_ctx_lens, rows);
    PA_RETURN_LAUNCH_STATUS();
}
