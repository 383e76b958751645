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
#include <cstdlib>

#include "pa_common.cuh"

namespace pa {

constexpr int kRowChunk = 8;     // rows per pass of the logits GEMV (one warp per vocab row)
constexpr int kLinRows = 16;     // activation rows per pass over the weights in linear_kn_kernel

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

// ---- LayerNorm: one CTA of 256 threads per row, two passes (mean, then biased variance) -----
__device__ __forceinline__ float block_sum_256(float v, float* sm) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = (lane < 8) ? sm[lane] : 0.f;
    r = warp_sum(r);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int rows, int hidden,
                                                         float eps, float* __restrict__ out) {
    __shared__ float sm[8];
    const int row = blockIdx.x;
    const float* xr = x + (int64_t)row * hidden;
    float s = 0.f;
    for (int j = threadIdx.x; j < hidden; j += 256) s += xr[j];
    const float mean = block_sum_256(s, sm) / (float)hidden;
    float v = 0.f;
    for (int j = threadIdx.x; j < hidden; j += 256) {
        const float d = xr[j] - mean;
        v = fmaf(d, d, v);
    }
    const float var = block_sum_256(v, sm) / (float)hidden;
    // layer_norm.hpp:33 `T inv_std = 1.0 / std::sqrt(var + epsilon_)`, T = float: float square root, double division
    const float inv_std = (float)(1.0 / (double)sqrtf(var + eps));
    float* o = out + (int64_t)row * hidden;
    for (int j = threadIdx.x; j < hidden; j += 256) o[j] = (xr[j] - mean) * inv_std * gamma[j] + beta[j];
}

// ---- LayerNorm + dynamic row quantisation fused (the INT8 decoder's LN2 -> int8_quant -> fc1 hand-off): the
// normalised row never goes to memory as f32 unless `out` is given; scale and int8 values are bit-identical to
// layer_norm_kernel followed by row_quantize_dynamic_kernel (same expressions, same order). ----
__global__ void __launch_bounds__(256) layer_norm_quantize_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, int rows, int hidden,
                                                                  float eps, float* __restrict__ out,
                                                                  float* __restrict__ scales, int8_t* __restrict__ q) {
    __shared__ float sm[8];
    const int row = blockIdx.x;
    const float* xr = x + (int64_t)row * hidden;
    float s = 0.f;
    for (int j = threadIdx.x; j < hidden; j += 256) s += xr[j];
    const float mean = block_sum_256(s, sm) / (float)hidden;
    float v = 0.f;
    for (int j = threadIdx.x; j < hidden; j += 256) {
        const float d = xr[j] - mean;
        v = fmaf(d, d, v);
    }
    const float var = block_sum_256(v, sm) / (float)hidden;
    const float inv_std = (float)(1.0 / (double)sqrtf(var + eps));
    float m = 0.f;
    for (int j = threadIdx.x; j < hidden; j += 256) {
        const float y = (xr[j] - mean) * inv_std * gamma[j] + beta[j];
        if (out) out[(int64_t)row * hidden + j] = y;
        m = fmaxf(m, fabsf(y));
    }
    m = warp_max(m);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    m = (lane < 8) ? sm[lane] : 0.f;
    m = warp_max(m);
    const float scale = __fdiv_rn(127.f, __fadd_rn(m, 1e-6f));
    if (threadIdx.x == 0) scales[row] = scale;
    int8_t* qr = q + (int64_t)row * hidden;
    for (int j = threadIdx.x; j < hidden; j += 256) {
        const float y = (xr[j] - mean) * inv_std * gamma[j] + beta[j];
        float r = roundf(__fmul_rn(y, scale));
        r = fminf(127.f, fmaxf(-128.f, r));
        qr[j] = (int8_t)(int)r;
    }
}

// ---- dynamic per-row activation quantisation in ONE kernel: compute_minmax_scale (int8_quant.cpp:59-64)
// then batch_quantize (int8_quant.cpp:15-28) of the same row; bit-identical to the two separate kernels. ----
__global__ void __launch_bounds__(256) row_quantize_dynamic_kernel(const float* __restrict__ x, int rows, int dim,
                                                                   float* __restrict__ scales, int8_t* __restrict__ q) {
    __shared__ float sm[8];
    const int row = blockIdx.x;
    const float* xr = x + (int64_t)row * dim;
    float m = 0.f;
    for (int d = threadIdx.x; d < dim; d += 256) m = fmaxf(m, fabsf(xr[d]));
    m = warp_max(m);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    m = (lane < 8) ? sm[lane] : 0.f;
    m = warp_max(m);
    const float scale = __fdiv_rn(127.f, __fadd_rn(m, 1e-6f));
    if (threadIdx.x == 0) scales[row] = scale;
    int8_t* qr = q + (int64_t)row * dim;
    for (int d = threadIdx.x; d < dim; d += 256) {
        float r = roundf(__fmul_rn(xr[d], scale));  // half away from zero (std::round)
        r = fminf(127.f, fmaxf(-128.f, r));
        qr[d] = (int8_t)(int)r;
    }
}

// ---- linear, W [K, N] row-major: out[r, n] = act(bias[n] + sum_k x[r,k] W[k,n]) ----------
// CTA = 8 warps x 32 columns: warp w streams rows k = k0 + w, k0 + w + 8, ... of a 128-byte wide column
// strip (one coalesced line per k) inside this CTA's K slice, keeps up to kRowChunk accumulators per
// lane, and the 8 warps are summed through shared memory.  grid = (N/32 strips, row chunks, K slices):
// with few strips (fc2 of a small model: N = 768 -> 24) the K dimension is sliced across CTAs so the
// whole chip streams the weights; slice partials go to a scratch buffer and a second tiny kernel adds
// them in a fixed order (deterministic, unlike float atomics).
__global__ void __launch_bounds__(256) linear_kn_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                        const float* __restrict__ bias, int rows, int K, int N,
                                                        int act, int kslice, float* __restrict__ out,
                                                        float* __restrict__ partial) {
    constexpr int KT = 256;                          // k rows staged per tile
    __shared__ __align__(16) float xs[KT][kLinRows];  // activations of the tile, [k][row]: one k = 4 x LDS.128
    __shared__ float red[8][kLinRows][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 32 + lane;
    const int r0 = blockIdx.y * kLinRows;
    const int nr = min(kLinRows, rows - r0);
    const int k0 = blockIdx.z * kslice, k1 = min(K, k0 + kslice);
    float acc[kLinRows];
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
    for (int kt = k0; kt < k1; kt += KT) {
        const int kn = min(KT, k1 - kt);
        __syncthreads();  // previous tile fully consumed
        if ((int)threadIdx.x < kn) {
#pragma unroll
            for (int r = 0; r < kLinRows; ++r)
                xs[threadIdx.x][r] = (r < nr) ? x[(int64_t)(r0 + r) * K + kt + threadIdx.x] : 0.f;
        }
        __syncthreads();
        if (n < N) {
#pragma unroll 4
            for (int kk = warp; kk < kn; kk += 8) {
                const float w = __ldcs(W + (int64_t)(kt + kk) * N + n);  // streamed once per row chunk
                const float4* xv = reinterpret_cast<const float4*>(xs[kk]);
#pragma unroll
                for (int r4 = 0; r4 < kLinRows / 4; ++r4) {
                    const float4 v = xv[r4];
                    acc[4 * r4 + 0] = fmaf(v.x, w, acc[4 * r4 + 0]);
                    acc[4 * r4 + 1] = fmaf(v.y, w, acc[4 * r4 + 1]);
                    acc[4 * r4 + 2] = fmaf(v.z, w, acc[4 * r4 + 2]);
                    acc[4 * r4 + 3] = fmaf(v.w, w, acc[4 * r4 + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) red[warp][r][lane] = acc[r];
    __syncthreads();
    // 256 threads finish kLinRows x 32 outputs (8 rows per pass)
    const int cc = threadIdx.x & 31;
    const int nn = blockIdx.x * 32 + cc;
    for (int rr = threadIdx.x >> 5; rr < kLinRows; rr += 8) {
        if (rr < nr && nn < N) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][rr][cc];
            if (partial) {
                partial[((int64_t)blockIdx.z * rows + r0 + rr) * N + nn] = s;
            } else {
                s += bias ? bias[nn] : 0.f;
                if (act == PA_ACT_RELU) s = fmaxf(s, 0.f);
                out[(int64_t)(r0 + rr) * N + nn] = s;
            }
        }
    }
}

__global__ void linear_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias, int rows, int N,
                                     int nslices, int act, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * N) return;
    float s = bias ? bias[i % N] : 0.f;
    for (int z = 0; z < nslices; ++z) s += partial[(int64_t)z * rows * N + i];
    if (act == PA_ACT_RELU) s = fmaxf(s, 0.f);
    out[i] = s;
}

// ---- logits against the tied embedding E [vocab, hidden]: one warp per vocab row, 16-byte loads.
// Optionally folds the greedy sampler in: best[r] = max over v of the 64-bit key
// (order-preserving bits of (logit / T or logit * T) << 32 | ~v), so the maximum key is the FIRST maximum
// (std::max_element, cuda_decoder.cu:13 / int8_decoder.cpp:103); argmax_decode_kernel turns it into ids. ----
__device__ __forceinline__ uint32_t orderable_f32(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <typename WT>
__global__ void __launch_bounds__(256) logits_kernel(const float* __restrict__ x, const WT* __restrict__ E, int rows,
                                                     int hidden, int vocab, float qscale, float* __restrict__ logits,
                                                     unsigned long long* __restrict__ best, float t, int divide) {
    constexpr int EPL = 16 / (int)sizeof(WT);  // elements per 16-byte load: 4 (f32) or 16 (int8)
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= vocab) return;
    const WT* e = E + (int64_t)v * hidden;
    const bool vec_ok = (hidden % EPL == 0) && (((uintptr_t)E & 15) == 0);
    for (int r0 = 0; r0 < rows; r0 += kRowChunk) {
        const int nr = min(kRowChunk, rows - r0);
        float acc[kRowChunk];
#pragma unroll
        for (int r = 0; r < kRowChunk; ++r) acc[r] = 0.f;
        if (vec_ok) {
            for (int j = lane * EPL; j < hidden; j += 32 * EPL) {
                float w[EPL];
                const uint4 raw = ldg_stream_128(e + j);
                if (sizeof(WT) == 4) {
                    w[0] = __uint_as_float(raw.x); w[1 % EPL] = __uint_as_float(raw.y);
                    w[2 % EPL] = __uint_as_float(raw.z); w[3 % EPL] = __uint_as_float(raw.w);
                } else {
                    const uint32_t ww[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {  // exact int8 -> f32 through the mantissa of 2^23 (no I2F)
                        const uint32_t xw = ww[i] ^ 0x80808080u;
                        w[(4 * i + 0) % EPL] = __uint_as_float(__byte_perm(xw, 0x4B000000u, 0x7440)) - 8388736.f;
                        w[(4 * i + 1) % EPL] = __uint_as_float(__byte_perm(xw, 0x4B000000u, 0x7441)) - 8388736.f;
                        w[(4 * i + 2) % EPL] = __uint_as_float(__byte_perm(xw, 0x4B000000u, 0x7442)) - 8388736.f;
                        w[(4 * i + 3) % EPL] = __uint_as_float(__byte_perm(xw, 0x4B000000u, 0x7443)) - 8388736.f;
                    }
                }
#pragma unroll
                for (int r = 0; r < kRowChunk; ++r) {
                    if (r < nr) {
                        const float* xr = x + (int64_t)(r0 + r) * hidden + j;
#pragma unroll
                        for (int i = 0; i < EPL; ++i) acc[r] = fmaf(__ldg(xr + i), w[i], acc[r]);
                    }
                }
            }
        } else {
            for (int j = lane; j < hidden; j += 32) {
                const float w = (float)e[j];
#pragma unroll
                for (int r = 0; r < kRowChunk; ++r)
                    if (r < nr) acc[r] = fmaf(__ldg(x + (int64_t)(r0 + r) * hidden + j), w, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRowChunk; ++r) {
            if (r < nr) {
                float s = warp_sum(acc[r]);
                if (sizeof(WT) == 1) s = __fdiv_rn(s, qscale);
                if (lane == 0) {
                    logits[(int64_t)(r0 + r) * vocab + v] = s;
                    if (best) {
                        const float sv = divide ? __fdiv_rn(s, t) : __fmul_rn(s, t);
                        if (sv == sv)  // NaN never wins (std::max_element keeps the first element)
                            atomicMax(best + r0 + r, ((unsigned long long)orderable_f32(sv) << 32) | (0xffffffffu - (uint32_t)v));
                    }
                }
            }
        }
    }
}

__global__ void argmax_decode_kernel(unsigned long long* __restrict__ best, int rows, int32_t* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) {
        const unsigned long long k = best[r];
        out[r] = (k == 0ull) ? 0 : (int32_t)(0xffffffffu - (uint32_t)(k & 0xffffffffull));
        best[r] = 0ull;  // ready for the next step
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

// ---- apply_rotary_embedding (attention/attention_kernel_utils.cuh:20-35): pairwise rotation of q and k
// rows [rows, H, D] with the interleaved table rotary_emb[token * D + d] = cos, [.. + d + 1] = sin, token =
// positions[row].  One thread per (row, head, pair); same operation order as the reference (two products and
// one add/sub per output, no FMA contraction) so the result is bit-exact. ----
__global__ void rope_kernel(float* __restrict__ q, float* __restrict__ k, const float* __restrict__ rope,
                            const int32_t* __restrict__ positions, int64_t n_pairs, int H, int D, int T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int half = D >> 1;
    const int64_t rh = i / half;  // row * H + head
    const int d = (int)(i - rh * half) * 2;
    const int row = (int)(rh / H);
    const int tok = positions[row];
    if (tok < 0 || tok >= T) return;  // outside the table: left unrotated
    const float c = rope[(int64_t)tok * D + d], s = rope[(int64_t)tok * D + d + 1];
    const int64_t o = rh * D + d;
    if (q) {
        const float x0 = q[o], x1 = q[o + 1];
        q[o] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        q[o + 1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
    }
    if (k) {
        const float x0 = k[o], x1 = k[o + 1];
        k[o] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        k[o + 1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
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
    layer_norm_kernel<<<rows, 256, 0, as_stream(stream)>>>(d_x, d_gamma, d_beta, rows, hidden, eps, d_out);
    PA_RETURN_LAUNCH_STATUS();
}

// ---- linear for MANY rows (batched decode / prefill): register-tiled fp32 GEMM -----------------------------------
// The strip kernel above re-streams the weights once per 16 rows and spends one shared-memory read per four FMAs
// (measured ~17 TFLOP/s: two thirds of a CUDADecoder step at the C2 shape).  Here a CTA of 256 threads owns a 64-row x
// 128-column tile: thread (warp w, lane l) keeps an 8 x 4 accumulator block (rows 8w..8w+7, columns 4l..4l+3), the K
// loop stages W [32][128] and x [64][32] tiles with cp.async (3 stages, 16-byte copies; the W rows are whole 512-byte
// runs of the reference's [K, N] layout) and per 4 k-steps issues 4 + 8 LDS.128 for 128 FMAs, so the FMA pipe, not
// shared memory, bounds it.  fp32 throughout (the reference's CUDADecoder<float> arithmetic; only the summation order
// differs).  Needs N % 4 == 0 and K % 4 == 0 (16-byte alignment of the rows); other shapes keep the strip kernel.
constexpr int kGemmBM = 64, kGemmBN = 128, kGemmKT = 32, kGemmStages = 3;
constexpr int kGemmXStride = kGemmKT + 4;  // floats; keeps every x row 16-byte aligned, staggers banks for the copies

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(256) linear_gemm_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                          const float* __restrict__ bias, int rows, int K, int N, int act,
                                                          int kslice, float* __restrict__ out, float* __restrict__ partial) {
    extern __shared__ __align__(16) float gsm[];
    float* ws = gsm;                                             // [stages][KT][BN]
    float* xs = gsm + kGemmStages * kGemmKT * kGemmBN;           // [stages][BM][XStride]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kGemmBN, r0 = blockIdx.y * kGemmBM;
    const int k0 = blockIdx.z * kslice, k1 = min(K, k0 + kslice);
    const int ntiles = (k1 - k0 + kGemmKT - 1) / kGemmKT;

    auto load_tile = [&](int t, int st) {
        const int kt = k0 + t * kGemmKT;
        // W tile: 32 rows x 128 floats = 1024 16-byte pieces, 4 per thread
        float* wd = ws + st * kGemmKT * kGemmBN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int p = threadIdx.x + i * 256;
            const int kr = p >> 5, c4 = (p & 31) * 4;
            const bool ok = (kt + kr < k1) && (n0 + c4 < N);
            cp_async16(smem_u32(wd + kr * kGemmBN + c4), W + (int64_t)(ok ? kt + kr : 0) * N + (ok ? n0 + c4 : 0), ok);
        }
        // x tile: 64 rows x 32 floats = 512 pieces, 2 per thread
        float* xd = xs + st * kGemmBM * kGemmXStride;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p = threadIdx.x + i * 256;
            const int r = p >> 3, c4 = (p & 7) * 4;
            const bool ok = (r0 + r < rows) && (kt + c4 < k1);
            cp_async16(smem_u32(xd + r * kGemmXStride + c4), x + (int64_t)(ok ? r0 + r : 0) * K + (ok ? kt + c4 : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;

    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < ntiles) load_tile(s, s);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int t = 0; t < ntiles; ++t) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kGemmStages - 2) : "memory");
        __syncthreads();  // tile t has landed for everyone; tile t-1's buffer is free
        if (t + kGemmStages - 1 < ntiles) load_tile(t + kGemmStages - 1, (t + kGemmStages - 1) % kGemmStages);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const float* wt = ws + (t % kGemmStages) * kGemmKT * kGemmBN + lane * 4;
        const float* xt = xs + (t % kGemmStages) * kGemmBM * kGemmXStride + warp * 8 * kGemmXStride;
#pragma unroll
        for (int kk = 0; kk < kGemmKT; kk += 4) {
            float4 wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4*>(wt + (kk + j) * kGemmBN);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 xv = *reinterpret_cast<const float4*>(xt + r * kGemmXStride + kk);  // broadcast
                const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[r][0] = fmaf(xa[j], wv[j].x, acc[r][0]);
                    acc[r][1] = fmaf(xa[j], wv[j].y, acc[r][1]);
                    acc[r][2] = fmaf(xa[j], wv[j].z, acc[r][2]);
                    acc[r][3] = fmaf(xa[j], wv[j].w, acc[r][3]);
                }
            }
        }
    }
    const int n = n0 + lane * 4;
    if (n >= N) return;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!partial && bias) b4 = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = r0 + warp * 8 + r;
        if (row >= rows) continue;
        float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        if (partial) {
            *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.z * rows + row) * N + n) = v;
        } else {
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            if (act == PA_ACT_RELU) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            *reinterpret_cast<float4*>(out + (int64_t)row * N + n) = v;
        }
    }
}

// K-slice geometry of pa_linear_f32 (shared with pa_linear_workspace_bytes).
static bool linear_uses_gemm(const void* x, const void* W, const void* out, const void* bias, int rows, int K, int N) {
    // the register-tiled kernel needs 16-byte aligned rows; few rows stay on the strip kernel (weights streamed once
    // either way, and its K slicing covers the chip better for tiny problems)
    if (const char* env = getenv("PA_LINEAR_GEMM")) {
        if (atoi(env) == 0) return false;
    } else if ((int64_t)rows * K * N < 500000000ll) {
        // small layers (GPT-2-small MLP at batch 64: 0.15 GMAC, weights L2-resident): too few 128-column tiles to fill
        // the chip without heavy K slicing; the strip kernel measured 14 % faster there
        return false;
    }
    return rows >= 16 && N % 4 == 0 && K % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)W % 16 == 0 &&
           (uintptr_t)out % 16 == 0 && (uintptr_t)bias % 16 == 0;
}

// linear_tf32x3.cu: 3xTF32 tcgen05 kernel over the same [rows, K] x [K, N] operands (stream-K or K-sliced grid + sum kernel)
size_t pa_linear_tc_workspace_bytes(int rows, int K, int N, int sm_count);
int pa_linear_tc_run(const float* d_x, const float* d_W, const float* d_bias, int rows, int K, int N, int act,
                     float* d_out, void* d_ws, size_t ws_bytes, int packed, int sm_count, cudaStream_t st);
size_t pa_linear_tc_pack_floats(int K, int N);
int pa_linear_tc_pack(const float* d_W, float* d_Wp, int K, int N, cudaStream_t st);

// The tensor-core kernel serves what the register-tiled kernel serves (>= 16 rows, 16-byte aligned rows of x, W, out)
// wherever the layer is large enough for the tile kernels at all; PA_LINEAR_TC=0 keeps the fp32 SIMT kernels
// (bit-level fp32 arithmetic), PA_LINEAR_TC=1 forces it for every eligible shape.
static bool linear_uses_tc(const void* x, const void* W, const void* out, const void* bias, int rows, int K, int N) {
    if (const char* env = getenv("PA_LINEAR_TC")) {
        if (atoi(env) == 0) return false;
    } else if ((int64_t)rows * K * N < 50000000ll) {
        return false;   // tiny layers: launch-bound either way, keep the fp32 kernels
    }
    return rows >= 16 && N % 4 == 0 && K % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)W % 16 == 0 &&
           (uintptr_t)out % 16 == 0 && (uintptr_t)bias % 16 == 0;
}

static int linear_slices(int rows, int K, int N, int sm_count, int* kslice_out, bool gemm) {
    const int strips = gemm ? (N + kGemmBN - 1) / kGemmBN : (N + 31) / 32;
    const int chunks = gemm ? (rows + kGemmBM - 1) / kGemmBM : (rows + kLinRows - 1) / kLinRows;
    // slice K until ~2 CTAs (strip kernel) / ~1 CTA (tile kernel) per SM stream the weights; every slice keeps
    // >= 64 k-rows (8 per warp) / >= 128 (4 K tiles)
    const int want = gemm ? sm_count : 2 * sm_count;
    int nslices = (want + strips * chunks - 1) / (strips * chunks);
    const int min_k = gemm ? 128 : 64;
    if (nslices > K / min_k) nslices = K / min_k;
    if (nslices < 1) nslices = 1;
    int kslice = (K + nslices - 1) / nslices;
    if (gemm) kslice = (kslice + kGemmKT - 1) / kGemmKT * kGemmKT;  // whole K tiles per slice (16-byte aligned starts)
    if (kslice_out) *kslice_out = kslice;
    return (K + kslice - 1) / kslice;
}

PA_API size_t pa_linear_workspace_bytes(int rows, int K, int N) {
    if (rows <= 0 || K <= 0 || N <= 0) return 0;
    const DeviceInfo& di = device_info();
    const int sm = di.ok ? di.sm_count : 148;
    // any of the kernels may run (the choice also depends on pointer alignment): size for the largest need
    const int a = linear_slices(rows, K, N, sm, nullptr, false);
    const int b = (rows >= 16 && N % 4 == 0 && K % 4 == 0) ? linear_slices(rows, K, N, sm, nullptr, true) : 1;  // (env may force it)
    const int nslices = a > b ? a : b;
    const size_t simt = nslices > 1 ? (size_t)nslices * rows * N * sizeof(float) : 0;
    const size_t tc = K % 4 == 0 ? pa_linear_tc_workspace_bytes(rows, K, N, sm) : 0;
    return simt > tc ? simt : tc;
}

// The K-slice partials live in the CALLER's workspace (no library-owned scratch: a pointer captured in a CUDA
// graph stays valid, streams never share it).  Without a large enough workspace the layer runs unsliced.
PA_API int pa_linear_f32(const float* d_x, const float* d_W, const float* d_bias, int rows, int K, int N, int act,
                         float* d_out, void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_W && d_out && rows >= 0 && K > 0 && N > 0);
    PA_CHECK_ARG(act == PA_ACT_NONE || act == PA_ACT_RELU);
    PA_CHECK_ARG(d_x != d_out);
    if (rows == 0) return PA_OK;
    const DeviceInfo& di = device_info();
    if (!di.ok) return PA_ERR_NO_DEVICE;
    if (linear_uses_tc(d_x, d_W, d_out, d_bias, rows, K, N))
        return pa_linear_tc_run(d_x, d_W, d_bias, rows, K, N, act, d_out, d_workspace, workspace_bytes, 0, di.sm_count,
                                as_stream(stream));
    const bool gemm = linear_uses_gemm(d_x, d_W, d_out, d_bias, rows, K, N);
    int kslice = K;
    int nslices = linear_slices(rows, K, N, di.sm_count, &kslice, gemm);
    float* partial = nullptr;
    if (nslices > 1) {
        if (d_workspace && workspace_bytes >= (size_t)nslices * rows * N * sizeof(float) && (uintptr_t)d_workspace % 16 == 0) {
            partial = static_cast<float*>(d_workspace);
        } else {
            nslices = 1;
            kslice = K;
        }
    }
    cudaError_t e;
    if (gemm) {
        const size_t smem = (size_t)kGemmStages * (kGemmKT * kGemmBN + kGemmBM * kGemmXStride) * sizeof(float);
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr_set[dev & 63]) {
            e = cudaFuncSetAttribute(linear_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set[dev & 63] = true;
        }
        dim3 grid((unsigned)((N + kGemmBN - 1) / kGemmBN), (unsigned)((rows + kGemmBM - 1) / kGemmBM), (unsigned)nslices);
        linear_gemm_kernel<<<grid, 256, smem, as_stream(stream)>>>(d_x, d_W, d_bias, rows, K, N, act, kslice, d_out, partial);
    } else {
        const int strips = (N + 31) / 32, chunks = (rows + kLinRows - 1) / kLinRows;
        dim3 grid((unsigned)strips, (unsigned)chunks, (unsigned)nslices);
        linear_kn_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_x, d_W, d_bias, rows, K, N, act, kslice, d_out, partial);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (nslices > 1) {
        const int64_t n = (int64_t)rows * N;
        linear_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(partial, d_bias, rows, N,
                                                                                        nslices, act, d_out);
    }
    PA_RETURN_LAUNCH_STATUS();
}

// ---- packed weights: the tensor-core kernel's own layout (every tile's K block one contiguous 16 KB run) ----
PA_API size_t pa_linear_pack_bytes(int K, int N) {
    if (K <= 0 || N <= 0) return 0;
    return pa_linear_tc_pack_floats(K, N) * sizeof(float);
}

PA_API int pa_linear_pack_f32(const float* d_W, float* d_W_packed, int K, int N, pa_stream_t stream) {
    PA_CHECK_ARG(d_W && d_W_packed && K > 0 && N > 0 && (uintptr_t)d_W_packed % 16 == 0);
    return pa_linear_tc_pack(d_W, d_W_packed, K, N, as_stream(stream));
}

PA_API int pa_linear_f32_packed(const float* d_x, const float* d_W_packed, const float* d_bias, int rows, int K, int N,
                                int act, float* d_out, void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_W_packed && d_out && rows >= 0 && K > 0 && N > 0);
    PA_CHECK_ARG(act == PA_ACT_NONE || act == PA_ACT_RELU);
    PA_CHECK_ARG(d_x != d_out);
    // the kernel's operand rules: 16-byte aligned rows of x (TMA), 16-byte aligned packed blocks
    if (K % 4 != 0 || (uintptr_t)d_x % 16 != 0 || (uintptr_t)d_W_packed % 16 != 0) return PA_ERR_UNSUPPORTED;
    if (rows == 0) return PA_OK;
    const DeviceInfo& di = device_info();
    if (!di.ok) return PA_ERR_NO_DEVICE;
    return pa_linear_tc_run(d_x, d_W_packed, d_bias, rows, K, N, act, d_out, d_workspace, workspace_bytes, 1, di.sm_count,
                            as_stream(stream));
}

PA_API int pa_logits_f32(const float* d_x, const float* d_E, int rows, int hidden, int vocab, float* d_logits,
                         pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_E && d_logits && rows >= 0 && hidden > 0 && vocab > 0);
    if (rows == 0) return PA_OK;
    logits_kernel<float><<<(vocab + 7) / 8, 256, 0, as_stream(stream)>>>(d_x, d_E, rows, hidden, vocab, 1.f, d_logits,
                                                                         nullptr, 1.f, 0);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_logits_i8(const float* d_x, const int8_t* d_E, float qscale, int rows, int hidden, int vocab,
                        float* d_logits, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_E && d_logits && rows >= 0 && hidden > 0 && vocab > 0 && qscale != 0.f);
    if (rows == 0) return PA_OK;
    logits_kernel<int8_t><<<(vocab + 7) / 8, 256, 0, as_stream(stream)>>>(d_x, d_E, rows, hidden, vocab, qscale,
                                                                          d_logits, nullptr, 1.f, 0);
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

// logits + greedy sample in one pass over the embedding table: d_best [rows] u64 scratch (zero-initialised
// once; reset by the call), d_out_ids [rows].  d_E is f32 (elem_bytes 4) or int8 (elem_bytes 1, dequantised
// as q / qscale).  divide: 1 = argmax(logit / T) (cuda_decoder.cu:7-14), 0 = argmax(logit * T)
// (int8_decoder.cpp:97-104).
PA_API int pa_logits_argmax(const float* d_x, const void* d_E, int elem_bytes, float qscale, int rows, int hidden,
                            int vocab, float temperature, int divide, float* d_logits,
                            unsigned long long* d_best, int32_t* d_out_ids, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_E && d_logits && d_best && d_out_ids && rows >= 0 && hidden > 0 && vocab > 0);
    PA_CHECK_ARG((elem_bytes == 4) || (elem_bytes == 1 && qscale != 0.f));
    PA_CHECK_ARG(!divide || temperature != 0.f);
    if (rows == 0) return PA_OK;
    cudaStream_t st = as_stream(stream);
    if (elem_bytes == 4)
        logits_kernel<float><<<(vocab + 7) / 8, 256, 0, st>>>(d_x, static_cast<const float*>(d_E), rows, hidden, vocab,
                                                              1.f, d_logits, d_best, temperature, divide);
    else
        logits_kernel<int8_t><<<(vocab + 7) / 8, 256, 0, st>>>(d_x, static_cast<const int8_t*>(d_E), rows, hidden,
                                                               vocab, qscale, d_logits, d_best, temperature, divide);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    argmax_decode_kernel<<<(rows + 127) / 128, 128, 0, st>>>(d_best, rows, d_out_ids);
    PA_RETURN_LAUNCH_STATUS();
}

// compute_minmax_scale + batch_quantize of every row in one kernel (int8_quant.cpp:59-64, 15-28).
PA_API int pa_layer_norm_quantize_i8(const float* d_x, const float* d_gamma, const float* d_beta, int rows, int hidden,
                                     float eps, float* d_out_f32, float* d_scales, int8_t* d_q, pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_gamma && d_beta && d_scales && d_q && rows >= 0 && hidden > 0);
    PA_CHECK_ARG((const void*)d_x != (const void*)d_q && d_x != d_out_f32);  // the row is read three times
    if (rows == 0) return PA_OK;
    layer_norm_quantize_kernel<<<rows, 256, 0, as_stream(stream)>>>(d_x, d_gamma, d_beta, rows, hidden, eps, d_out_f32,
                                                                     d_scales, d_q);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_row_quantize_dynamic_i8(const float* d_x, int rows, int dim, float* d_scales, int8_t* d_q,
                                      pa_stream_t stream) {
    PA_CHECK_ARG(d_x && d_scales && d_q && rows >= 0 && dim > 0);
    if (rows == 0) return PA_OK;
    row_quantize_dynamic_kernel<<<rows, 256, 0, as_stream(stream)>>>(d_x, rows, dim, d_scales, d_q);
    PA_RETURN_LAUNCH_STATUS();
}

// attention/attention_kernel_utils.cuh:20-35 apply_rotary_embedding over rows: d_q and/or d_k [rows, H, D] f32
// rotated in place with d_rope [T, D] (interleaved cos, sin) at token d_positions[row] (apply_on_k = d_k != NULL).
PA_API int pa_apply_rope_f32(float* d_q, float* d_k, const float* d_rope, const int32_t* d_positions, int rows,
                             int num_heads, int head_dim, int T, pa_stream_t stream) {
    PA_CHECK_ARG((d_q || d_k) && d_rope && d_positions && rows >= 0 && num_heads > 0 && head_dim > 0 && T > 0);
    PA_CHECK_ARG(head_dim % 2 == 0);
    if (rows == 0) return PA_OK;
    const int64_t n = (int64_t)rows * num_heads * (head_dim / 2);
    rope_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(d_q, d_k, d_rope, d_positions, n, num_heads,
                                                                            head_dim, T);
    PA_RETURN_LAUNCH_STATUS();
}
