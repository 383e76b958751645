// sampling.cu -- temperature softmax, top-k / top-p filter with the EOS rule, and inverse-CDF
// sampling over vocabulary logits on the device.
//
//   attention_cpu/softmax_lut.cpp:203-231  softmax_lut_vec: p = exp((x - max)/T) / (sum + 1e-6)
//   attention_cpu/softmax_lut.cpp:233-256  apply_topk_topp_filter: sort (prob, index) pairs descending
//       (std::greater on the pair: ties -> LARGER index first), zero every entry whose rank >= top_k or
//       whose preceding cumulative probability >= top_p, no renormalisation; then the EOS rule
//   attention/top_k_top_p_filter.cuh:55-111  the reference's GPU sampler (serial top-k in thread 0 and a
//       host rand() in device code): replaced by a radix select and a caller-supplied uniform number
//
// One CTA per row.  Ranks are never materialised: with the 64-bit key (float bits of p << 32 | index)
// "rank < top_k" is "key >= k-th largest key" (MSB-first radix select, 8 bits per pass), and "cumulative
// probability of all higher-ranked entries < top_p" is "key >= boundary key", found by the same descent
// over per-bucket probability sums.
#include <climits>

#include "pa_common.cuh"

namespace pa {

constexpr int kSampThreads = 1024;

__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = is_max ? warp_max(v) : warp_sum(v);
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : (is_max ? -INFINITY : 0.f);
    r = is_max ? warp_max(r) : warp_sum(r);
    __syncthreads();
    return r;  // every thread
}

// softmax_lut.cpp:203-231
__global__ void __launch_bounds__(kSampThreads) softmax_temperature_kernel(const float* __restrict__ logits, int V,
                                                                           float temperature, float* __restrict__ probs) {
    __shared__ float sm[32];
    const float* x = logits + (int64_t)blockIdx.x * V;
    float* p = probs + (int64_t)blockIdx.x * V;
    float mx = -1e9f;  // the reference's initial maximum
    for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, x[i]);
    mx = block_reduce(mx, sm, true);
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const float e = expf(__fdiv_rn(x[i] - mx, temperature));
        p[i] = e;
        s += e;
    }
    s = block_reduce(s, sm, false);
    const float inv = __fdiv_rn(1.f, s + 1e-6f);
    for (int i = threadIdx.x; i < V; i += blockDim.x) p[i] *= inv;
}

// attention_cpu/softmax_lut.cpp:60-82 fused_softmax_lut_inplace (= the row body of softmax_batch_parallel,
// :85-100): x = ((float)logit - max) * scale clamped to [-10, 10], idx = (int)((x + 10) * (res-1)/20)
// (truncation), p = lut[idx] / (sum + 1e-6).  BIT-EXACT: the table look-ups run in parallel, the sum is
// formed by one thread in the reference's sequential order (rows are attention rows, <= a few thousand
// entries), so `inv` and every product round exactly as on the CPU.
__global__ void __launch_bounds__(kSampThreads) softmax_lut_kernel(const int32_t* __restrict__ logits, int N, float scale,
                                                                   const float* __restrict__ lut, int resolution,
                                                                   float* __restrict__ probs) {
    __shared__ int smax[32];
    __shared__ float s_inv;
    const int32_t* x = logits + (int64_t)blockIdx.x * N;
    float* p = probs + (int64_t)blockIdx.x * N;
    int mx = INT_MIN;
    for (int i = threadIdx.x; i < N; i += blockDim.x) mx = max(mx, x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = smax[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = max(mx, smax[w]);
    const float max_x = 10.0f;
    const float inv_range = __fdiv_rn((float)(resolution - 1), __fmul_rn(2.f, max_x));
    const float fmx = (float)mx;  // static_cast<float>(logits[i]) - max_val: int32 max promoted to float
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float v = __fmul_rn(__fsub_rn((float)x[i], fmx), scale);
        v = fmaxf(-max_x, fminf(max_x, v));
        const int idx = (int)__fmul_rn(__fadd_rn(v, max_x), inv_range);
        p[i] = lut[idx];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.0f;
        for (int i = 0; i < N; ++i) sum = __fadd_rn(sum, p[i]);
        s_inv = __fdiv_rn(1.0f, __fadd_rn(sum, 1e-6f));
    }
    __syncthreads();
    const float inv = s_inv;
    for (int i = threadIdx.x; i < N; i += blockDim.x) p[i] = __fmul_rn(p[i], inv);
}

__device__ __forceinline__ uint64_t key_of(float p, int i) {
    return ((uint64_t)__float_as_uint(fmaxf(p, 0.f)) << 32) | (uint32_t)i;
}

// Keys >= the returned key are exactly the `k` largest (1 <= k <= V).
__device__ uint64_t select_kth_largest(const float* p, int V, int k, int* cnt) {
    uint64_t prefix = 0;
    int remaining = k;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) cnt[i] = 0;
        __syncthreads();
        const uint64_t hi_mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (int i = threadIdx.x; i < V; i += blockDim.x) {
            const uint64_t key = key_of(p[i], i);
            if ((key & hi_mask) == prefix) atomicAdd(&cnt[(key >> shift) & 0xff], 1);
        }
        __syncthreads();
        // highest bucket b with (count of buckets above b) < remaining <= (that + cnt[b])
        int b = 255, above = 0;
        while (b > 0 && above + cnt[b] < remaining) {
            above += cnt[b];
            --b;
        }
        remaining -= above;
        prefix |= (uint64_t)b << shift;
        __syncthreads();
    }
    return prefix;
}

// Smallest KEPT key under the top-p rule: rank i is kept iff the mass of all higher-ranked entries is
// < top_p (top_p > 0, so rank 0 is always kept and the boundary exists).  The kept set is a prefix of the
// ranking, so the boundary lies in the lowest non-empty bucket whose FIRST entry is still kept; descend
// into it, carrying the mass above.
__device__ uint64_t select_top_p_boundary(const float* p, int V, float top_p, int* cnt, float* sum) {
    uint64_t prefix = 0;
    float mass_above = 0.f;  // probability of all keys above the current prefix range
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            cnt[i] = 0;
            sum[i] = 0.f;
        }
        __syncthreads();
        const uint64_t hi_mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (int i = threadIdx.x; i < V; i += blockDim.x) {
            const uint64_t key = key_of(p[i], i);
            if ((key & hi_mask) == prefix) {
                const int bk = (int)((key >> shift) & 0xff);
                atomicAdd(&cnt[bk], 1);
                atomicAdd(&sum[bk], p[i]);
            }
        }
        __syncthreads();
        int best = -1;
        float best_m = mass_above, m = mass_above;  // m: mass above the first entry of bucket bk
        for (int bk = 255; bk >= 0; --bk) {
            if (cnt[bk] > 0) {
                if (m < top_p) {
                    best = bk;
                    best_m = m;
                } else {
                    break;
                }
            }
            m += sum[bk];
        }
        if (best < 0) best = 0;  // unreachable for top_p > 0 (kept for safety)
        mass_above = best_m;
        prefix |= (uint64_t)best << shift;
        __syncthreads();
    }
    return prefix;
}

// apply_topk_topp_filter (softmax_lut.cpp:233-256), in place on probs [rows, V].
__global__ void __launch_bounds__(kSampThreads) topk_topp_filter_kernel(float* __restrict__ probs, int V, int top_k,
                                                                        float top_p, int eos_token_id, float eos_thresh) {
    __shared__ int cnt[256];
    __shared__ float sum[256];
    float* p = probs + (int64_t)blockIdx.x * V;
    uint64_t thr = 0;  // keep keys >= thr
    bool drop_all = false;
    if (top_k > 0 && top_k < V) thr = select_kth_largest(p, V, top_k, cnt);
    if (top_p < 1.0f) {
        if (!(top_p > 0.f)) {
            drop_all = true;  // cum (0) >= top_p already at rank 0: every entry is zeroed
        } else {
            const uint64_t tp = select_top_p_boundary(p, V, top_p, cnt, sum);
            if (tp > thr) thr = tp;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < V; i += blockDim.x)
        if (drop_all || key_of(p[i], i) < thr) p[i] = 0.f;
    __syncthreads();
    // EOS hard threshold (:252-255)
    if (eos_token_id >= 0 && eos_token_id < V && p[eos_token_id] > eos_thresh)
        for (int i = threadIdx.x; i < V; i += blockDim.x)
            if (i != eos_token_id) p[i] = 0.f;
}

// Inverse-CDF sample in index order over (unnormalised) probabilities: the first index whose inclusive prefix
// sum exceeds u * total.  Rows with total == 0 return the argmax-free fallback 0.
__global__ void __launch_bounds__(kSampThreads) sample_kernel(const float* __restrict__ probs, int V,
                                                              const float* __restrict__ uniform,
                                                              int32_t* __restrict__ out_ids) {
    __shared__ float part[kSampThreads];
    __shared__ int chosen;
    const float* p = probs + (int64_t)blockIdx.x * V;
    const int per = (V + blockDim.x - 1) / blockDim.x;
    const int i0 = threadIdx.x * per, i1 = min(V, i0 + per);
    float s = 0.f;
    for (int i = i0; i < i1; ++i) s += p[i];
    part[threadIdx.x] = s;
    if (threadIdx.x == 0) chosen = -1;
    __syncthreads();
    if (threadIdx.x == 0) {  // 1024 partial sums: a serial scan is ~1 us and keeps the order well defined
        float total = 0.f;
        for (int t = 0; t < (int)blockDim.x; ++t) total += part[t];
        const float target = uniform[blockIdx.x] * total;
        float run = 0.f;
        int t = 0;
        for (; t < (int)blockDim.x - 1; ++t) {
            if (run + part[t] > target) break;
            run += part[t];
        }
        // scan inside chunk t
        int pick = -1, last_nz = -1;
        for (int i = t * per; i < min(V, (t + 1) * per); ++i) {
            if (p[i] > 0.f) last_nz = i;
            run += p[i];
            if (run > target && p[i] > 0.f) {
                pick = i;
                break;
            }
        }
        if (pick < 0) {  // rounding at the very end of the distribution: last non-zero entry overall
            pick = last_nz;
            for (int i = V - 1; pick < 0 && i >= 0; --i)
                if (p[i] > 0.f) pick = i;
        }
        chosen = pick < 0 ? 0 : pick;
    }
    __syncthreads();
    if (threadIdx.x == 0) out_ids[blockIdx.x] = chosen;
}

}  // namespace pa

using namespace pa;

PA_API int pa_softmax_temperature(const float* d_logits, int rows, int vocab, float temperature, float* d_probs,
                                  pa_stream_t stream) {
    PA_CHECK_ARG(d_logits && d_probs && rows >= 0 && vocab > 0 && temperature != 0.f);
    if (rows == 0) return PA_OK;
    softmax_temperature_kernel<<<rows, kSampThreads, 0, as_stream(stream)>>>(d_logits, vocab, temperature, d_probs);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_softmax_lut_i32(const int32_t* d_logits, int rows, int n, float scale, const float* d_lut,
                              int resolution, float* d_probs, pa_stream_t stream) {
    PA_CHECK_ARG(d_logits && d_lut && d_probs && rows >= 0 && n > 0 && resolution > 1);
    if (rows == 0) return PA_OK;
    softmax_lut_kernel<<<rows, kSampThreads, 0, as_stream(stream)>>>(d_logits, n, scale, d_lut, resolution, d_probs);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_topk_topp_filter(float* d_probs, int rows, int vocab, int top_k, float top_p, int eos_token_id,
                               float eos_thresh, pa_stream_t stream) {
    PA_CHECK_ARG(d_probs && rows >= 0 && vocab > 0);
    if (rows == 0) return PA_OK;
    topk_topp_filter_kernel<<<rows, kSampThreads, 0, as_stream(stream)>>>(d_probs, vocab, top_k, top_p, eos_token_id,
                                                                         eos_thresh);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_sample_from_probs(const float* d_probs, int rows, int vocab, const float* d_uniform,
                                int32_t* d_out_ids, pa_stream_t stream) {
    PA_CHECK_ARG(d_probs && d_uniform && d_out_ids && rows >= 0 && vocab > 0);
    if (rows == 0) return PA_OK;
    sample_kernel<<<rows, kSampThreads, 0, as_stream(stream)>>>(d_probs, vocab, d_uniform, d_out_ids);
    PA_RETURN_LAUNCH_STATUS();
}
