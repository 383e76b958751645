// paged_decode.cu -- single-query (decode) attention over the paged KV cache.
//
// Replaces attention/paged_flash_attention_kernel_fused.cu:5-90 (a1) and
// attention/paged_flash_attention_kernel_fused_overlap.cu:5-91 (a2) with the math of
// attention_cpu/cpu_attention_kernel.cpp:36-129 (global softmax).
//
// The op is HBM-bound (~1 flop/B): every K/V byte is read once.  Design:
//   * the unit of work is a 16-token sub-tile of one page: K [16][D] + V [16][D],
//     contiguous in the pools, located through the dense int32 page table;
//   * a warp owns a unit: lane = 8*g + c reads token rows 4p+g (p = 0..3) and the
//     16 B dim chunk c of each row, so each quarter-warp touches 128 contiguous
//     bytes (coalesced in global memory, conflict-free in shared memory);
//     QK^T is reduced over the 8 chunk lanes with 3 xor-shuffles per token row
//     and the online softmax (running m, l) lives per lane group -- no shared
//     memory or block barrier in the steady state;
//   * direct kernel (a1): grid (rows, splits), 128-bit ld.global.nc loads
//     straight into registers, split-KV partials + LSE combine kernel;
//   * overlap kernel (a2): persistent, one CTA per SM, every warp is an
//     independent streaming engine with its own multi-stage ring of TMA bulk
//     copies (cp.async.bulk -> UBLKCP, mbarrier complete_tx), ~190 KB of page
//     loads in flight per SM while the math runs; chunks of <= 16 units are
//     handed out through a global atomic counter (dynamic balance; a static
//     equal split lost ~4% to SM-to-SM rate differences, profiles/), each chunk
//     emits an (m, l, O) partial that a small combine kernel merges per row.
#include <cstdlib>

#include "mma_utils.cuh"
#include "pa_common.cuh"
#include "xchg.cuh"

namespace pa {

constexpr int kUnitTok = 16;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct DecodeArgs {
    const float* q;
    float* out;
    float* lse_out;
    float* part_m;  // non-null => emit un-normalised partials (multi-GPU split-KV)
    float* part_l;
    float* part_o;
    const uint8_t* k_pool;
    const uint8_t* v_pool;
    const float* k_scales;
    const float* v_scales;
    const int32_t* table;
    const int32_t* beam_ids;
    const int32_t* ctx_lens;
    const float* rope;
    float* ws_m;  // split / stream-K partials, log2 domain
    float* ws_l;
    float* ws_o;
    int num_beams, H, num_tiles, total_pages;
    int B, T, tile_size;
    int num_splits;   // direct kernel
    int evict_first;  // overlap kernel L2 hint
    float qscale;     // log2(e) / temperature
    // fused inter-GPU exchange (split-KV of one sequence across ranks; p2p_combine.cu layout)
    uint8_t* const* xch_peers;  // device array [xch_world] of exchange-buffer base pointers; null = off
    uint32_t* xch_epochs;       // [rows] step counters
    int* xch_status;
    int xch_rank, xch_world;
    int all_rows_in_ws;  // combine kernel: merge rows of a single chunk too (group kernel)
    unsigned int* row_done;  // streaming kernel: per-row count of finished chunks (the last one merges the row)
    const int* row_prefix;    // [B+1] chunk prefix per row, precomputed in global memory (group kernel, ragged)
    const int* group_prefix;  // [groups+1] chunk prefix per beam group (max over its rows)
};

template <int D, int KV>
struct Cfg {
    static constexpr int ES = KV == 0 ? 2 : (KV == 1 ? 1 : 4);  // bytes per element: fp16, int8, fp32 pools
    static constexpr int ROWB = D * ES;                        // bytes per token row
    static constexpr int VB = (ROWB / 8 >= 16) ? 16 : ROWB / 8;  // bytes per lane vector
    static constexpr int NV = ROWB / (8 * VB);                 // vectors per lane per row
    static constexpr int EV = VB / ES;                         // elements per vector
    static constexpr int E = NV * EV;                          // = D / 8 elements per lane
    static constexpr int WPV = VB / 4;
    static constexpr int W = NV * WPV;                         // 32-bit words per lane per row
    static constexpr int UNIT_BYTES = kUnitTok * ROWB;         // one K (or V) unit
    static constexpr int SCALE_BYTES = KV == 1 ? kUnitTok * 4 : 0;
    static constexpr int STAGE_BYTES = 2 * UNIT_BYTES + 2 * SCALE_BYTES;
    // head dim owned by (chunk lane c, element e)
    __device__ static __forceinline__ int dim_of(int c, int e) {
        return (e / EV) * 8 * EV + c * EV + (e % EV);
    }
    __device__ static __forceinline__ int byte_off(int c, int v) { return v * 8 * VB + c * VB; }
};

template <int D, int KV>
struct Acc {
    float m, l;
    float o[Cfg<D, KV>::E];
    __device__ __forceinline__ void reset() {
        m = -INFINITY;
        l = 0.f;
#pragma unroll
        for (int e = 0; e < Cfg<D, KV>::E; ++e) o[e] = 0.f;
    }
};

template <int KV, int W>
__device__ __forceinline__ void words_to_float(const uint32_t (&w)[W], float* f) {
    if (KV == 2) {  // fp32 pools (KVTileCache<float>, kv_tile_cache.cpp:127): the words are the values
#pragma unroll
        for (int i = 0; i < W; ++i) f[i] = __uint_as_float(w[i]);
    } else if (KV == 0) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    } else {
        // int8 -> f32 without the conversion unit (I2F runs at quarter rate and bounded this kernel at
        // 2.3 TB/s).  Flip the sign bit (u = b + 128 in [0, 255]) and drop the byte into mantissa bits
        // 8..15 of 2^15 with ONE PRMT: the float is exactly 32768 + u = b + 32896.  The constant is
        // folded out by the caller: sum_e q_e b_e = sum_e q_e f_e - 32896 sum_e q_e, so the inner loop
        // is one PRMT + one FFMA per element (costs ~8 mantissa bits of the partial sums; the offset is
        // removed per 16-token unit so it never accumulates).
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const uint32_t x = w[i] ^ 0x80808080u;
            f[4 * i + 0] = __uint_as_float(__byte_perm(x, 0x47000000u, 0x7404));
            f[4 * i + 1] = __uint_as_float(__byte_perm(x, 0x47000000u, 0x7414));
            f[4 * i + 2] = __uint_as_float(__byte_perm(x, 0x47000000u, 0x7424));
            f[4 * i + 3] = __uint_as_float(__byte_perm(x, 0x47000000u, 0x7434));
        }
    }
}

// One 16-token unit: 4 token rows per lane group.  kf/vf: this lane's words of rows
// 4p+g.  ksc/vsc: reciprocal scales of those rows (int8 only).  Scores are in log2
// units (q pre-multiplied by log2(e)/temperature).
template <int D, int KV, bool FULL>
__device__ __forceinline__ void unit_update_impl(const uint32_t (&kf)[4][Cfg<D, KV>::W],
                                                 uint32_t (&vf)[4][Cfg<D, KV>::W], const float (&ksc)[4],
                                                 const float (&vsc)[4], const float (&q)[Cfg<D, KV>::E],
                                                 float qoff, int nvalid, int g, Acc<D, KV>& a) {
    using C = Cfg<D, KV>;
    float s[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float f[C::E];
        words_to_float<KV, C::W>(kf[p], f);
        float acc = qoff;  // int8: qoff = -32896 * sum_e q_e (see words_to_float); fp16: 0
#pragma unroll
        for (int e = 0; e < C::E; ++e) acc = fmaf(q[e], f[e], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (KV == 1) acc *= ksc[p];
        s[p] = (FULL || 4 * p + g < nvalid) ? acc : -INFINITY;
    }
    const float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    // Partial unit: a lane group whose rows are all past the context end keeps its state (no early
    // return: the warp votes below).
    const bool dead = !FULL && !(mx > -INFINITY);
    const float m_new = dead ? a.m : fmaxf(a.m, mx);
    const float corr = dead ? 1.f : fast_exp2(a.m - m_new);  // a.m == -inf -> 0
    float pw[4];
    float ps = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        pw[p] = dead ? 0.f : fast_exp2(s[p] - m_new);  // masked rows -> 0
        ps += pw[p];
    }
    a.l = fmaf(a.l, corr, ps);
    a.m = m_new;
    if (__any_sync(0xffffffffu, corr != 1.f)) {  // the running max settles after a few units
#pragma unroll
        for (int e = 0; e < C::E; ++e) a.o[e] *= corr;
    }
    float wsum = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (!FULL && 4 * p + g >= nvalid) continue;  // never touch bytes past the context end
        float f[C::E];
        words_to_float<KV, C::W>(vf[p], f);
        const float w = (KV == 1) ? pw[p] * vsc[p] : pw[p];
        if (KV == 1) wsum += w;
#pragma unroll
        for (int e = 0; e < C::E; ++e) a.o[e] = fmaf(w, f[e], a.o[e]);
    }
    if (KV == 1) {  // sum_p w_p b = sum_p w_p f - 32896 sum_p w_p
        const float off = -32896.f * wsum;
#pragma unroll
        for (int e = 0; e < C::E; ++e) a.o[e] += off;
    }
}

// Full 16-token units (all but possibly the last unit of a row) take the branch-free path.
template <int D, int KV>
__device__ __forceinline__ void unit_update(const uint32_t (&kf)[4][Cfg<D, KV>::W],
                                            uint32_t (&vf)[4][Cfg<D, KV>::W], const float (&ksc)[4],
                                            const float (&vsc)[4], const float (&q)[Cfg<D, KV>::E],
                                            float qoff, int nvalid, int g, Acc<D, KV>& a) {
    if (nvalid == kUnitTok) unit_update_impl<D, KV, true>(kf, vf, ksc, vsc, q, qoff, nvalid, g, a);
    else unit_update_impl<D, KV, false>(kf, vf, ksc, vsc, q, qoff, nvalid, g, a);
}

// Merge the 4 lane groups of a warp (xor 8, 16); afterwards every lane holds the
// warp-wide (m, l, o) for its dim chunk.
template <int D, int KV>
__device__ __forceinline__ void warp_merge(Acc<D, KV>& a) {
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, a.m, off);
        const float lo = __shfl_xor_sync(0xffffffffu, a.l, off);
        const float mn = fmaxf(a.m, mo);
        const float wa = (a.m == -INFINITY) ? 0.f : fast_exp2(a.m - mn);
        const float wb = (mo == -INFINITY) ? 0.f : fast_exp2(mo - mn);
        a.l = a.l * wa + lo * wb;
#pragma unroll
        for (int e = 0; e < Cfg<D, KV>::E; ++e) {
            const float oo = __shfl_xor_sync(0xffffffffu, a.o[e], off);
            a.o[e] = a.o[e] * wa + oo * wb;
        }
        a.m = mn;
    }
}

enum EmitKind { kEmitFinal = 0, kEmitWorkspace = 1 };

// Final / partial write of one row's merged (M, L, O[d]) by thread d (< D).
__device__ __forceinline__ void emit_row(const DecodeArgs& a, int kind, int64_t row, int64_t slot,
                                         int D, int d, float M, float L, float O) {
    if (kind == kEmitWorkspace) {
        a.ws_o[slot * D + d] = O;
        if (d == 0) {
            a.ws_m[slot] = M;
            a.ws_l[slot] = L;
        }
    } else if (a.part_m) {
        a.part_o[row * D + d] = O;
        if (d == 0) {
            a.part_m[row] = M * kLn2;  // natural-log units at the API
            a.part_l[row] = L;
        }
    } else {
        a.out[row * D + d] = O / (L + 1e-6f);  // softmax_lut.cpp:224 epsilon (App. A D4)
        if (d == 0 && a.lse_out) a.lse_out[row] = (L > 0.f) ? (M + log2f(L)) * kLn2 : -INFINITY;
    }
}

// Cross-warp merge through shared memory.  red: [NW][D+2].  Must be called by all
// threads of the CTA (contains one __syncthreads).
template <int D, int KV, int NW>
__device__ __forceinline__ void cta_merge_emit(float* red, Acc<D, KV>& acc, int warp, int lane,
                                               const DecodeArgs& a, int kind, int64_t row,
                                               int64_t slot) {
    using C = Cfg<D, KV>;
    warp_merge<D, KV>(acc);
    if (lane < 8) {
        float* r = red + warp * (D + 2);
#pragma unroll
        for (int e = 0; e < C::E; ++e) r[C::dim_of(lane, e)] = acc.o[e];
        if (lane == 0) {
            r[D] = acc.m;
            r[D + 1] = acc.l;
        }
    }
    __syncthreads();
    const int d = threadIdx.x;
    if (d < D) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) M = fmaxf(M, red[w * (D + 2) + D]);
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float mw = red[w * (D + 2) + D];
            const float wt = (mw == -INFINITY) ? 0.f : fast_exp2(mw - M);
            L = fmaf(red[w * (D + 2) + D + 1], wt, L);
            O = fmaf(red[w * (D + 2) + d], wt, O);
        }
        emit_row(a, kind, row, slot, D, d, M, L, O);
    }
}

// q row -> registers, pre-scaled; optional pairwise RoPE (cpu_attention_kernel.cpp:13-19).
template <int D, int KV>
__device__ __forceinline__ float load_q(const DecodeArgs& a, int64_t row, int c,
                                        float (&q)[Cfg<D, KV>::E]) {
    using C = Cfg<D, KV>;
    const float* qr = a.q + row * D;
#pragma unroll
    for (int e = 0; e < C::E; e += 2) {
        const int d0 = C::dim_of(c, e);  // even, and dim_of(c, e+1) == d0 + 1
        float x0 = qr[d0], x1 = qr[d0 + 1];
        if (a.rope) {
            const float cs = a.rope[d0], sn = a.rope[d0 + 1];
            const float r0 = x0 * cs - x1 * sn, r1 = x0 * sn + x1 * cs;
            x0 = r0;
            x1 = r1;
        }
        q[e] = x0 * a.qscale;
        q[e + 1] = x1 * a.qscale;
    }
    if (KV != 1) return 0.f;
    // int8 pages: fold the +256 offset of the PRMT conversion (words_to_float) out of the inner loop
    float qs = 0.f;
#pragma unroll
    for (int e = 0; e < C::E; ++e) qs += q[e];
    return -32896.f * qs;
}

__device__ __forceinline__ int row_ctx(const DecodeArgs& a, int b) {
    int ctx = a.ctx_lens ? a.ctx_lens[b] : a.T;
    const int cap = a.num_tiles * a.tile_size;
    return ctx < 0 ? 0 : (ctx > cap ? cap : ctx);
}

// page_table.hpp:44-49 + kv_tile_cache.hpp:23: -1 when unmapped / out of range.
__device__ __forceinline__ int lookup_page(const DecodeArgs& a, int beam, int h, int tile) {
    const int64_t idx = ((int64_t)beam * a.H + h) * a.num_tiles + tile;
    if (idx < 0 || idx >= (int64_t)a.num_beams * a.H * a.num_tiles) return -1;
    const int page = __ldg(a.table + idx);
    return (page < 0 || page >= a.total_pages) ? -1 : page;
}

// ---- row merge by ONE warp -------------------------------------------------------------------------------------
// Reads `nc` partials (slots s0 .. s0+nc, log2-domain m) straight from L2 (ld.global.cg: they were written by other
// SMs); lane l ends up with O for dims l*VEC .. l*VEC+VEC and the (M, L) of the merged range.
__device__ __forceinline__ float ldcg_f32(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
template <int D>
__device__ __forceinline__ void merge_row_warp(const DecodeArgs& a, int64_t s0, int nc, int lane, float& M, float& L,
                                               float (&O)[D / 32]) {
    // Online merge in groups of G partials: all loads of a group (the G partial rows of 512 B, and m / l of the
    // group on lanes < G) are issued together, so a group costs ONE L2 round trip; the running maximum is raised
    // per group and the accumulator rescaled (a row of 37 chunks -- one GPU's share of a 128K-token sequence --
    // merges in 3 round trips instead of 37 dependent ones).
    constexpr int VEC = D / 32, G = 16;
    M = -INFINITY;
    float Lw = 0.f;  // lane-partial sum of l (lanes < G)
#pragma unroll
    for (int e = 0; e < VEC; ++e) O[e] = 0.f;
    for (int base = 0; base < nc; base += G) {
        const int cnt = min(G, nc - base);
        float v[G][VEC];
#pragma unroll
        for (int t = 0; t < G; ++t) {
            if (t < cnt) {
                const float* src = a.ws_o + (s0 + base + t) * D + lane * VEC;
                if (VEC == 4) {
                    const float4 x = __ldcg(reinterpret_cast<const float4*>(src));
                    v[t][0] = x.x; v[t][1] = x.y; v[t][2 % VEC] = x.z; v[t][3 % VEC] = x.w;
                } else {
                    const float2 x = __ldcg(reinterpret_cast<const float2*>(src));
                    v[t][0] = x.x; v[t][1] = x.y;
                }
            }
        }
        float mj = -INFINITY, lj = 0.f;
        if (lane < cnt) {
            mj = ldcg_f32(a.ws_m + s0 + base + lane);
            lj = ldcg_f32(a.ws_l + s0 + base + lane);
        }
        const float Mn = fmaxf(M, warp_max(mj));
        const float corr = (M == -INFINITY) ? 0.f : fast_exp2(M - Mn);  // Mn == -inf only if everything so far is empty
        const float wt = (mj == -INFINITY) ? 0.f : fast_exp2(mj - Mn);
        Lw = fmaf(Lw, corr, lj * wt);
#pragma unroll
        for (int e = 0; e < VEC; ++e) O[e] *= corr;
#pragma unroll
        for (int t = 0; t < G; ++t) {
            if (t < cnt) {
                const float w = __shfl_sync(0xffffffffu, wt, t);
#pragma unroll
                for (int e = 0; e < VEC; ++e) O[e] = fmaf(v[t][e], w, O[e]);
            }
        }
        M = Mn;
    }
    L = warp_sum(Lw);
}

// Final / partial / exchange emit of a merged row held as above (M in log2 units).
template <int D>
__device__ __forceinline__ void emit_merged_row(const DecodeArgs& a, int64_t row, int lane, float M, float L,
                                                const float (&O)[D / 32]) {
    constexpr int VEC = D / 32;
    if (a.xch_peers) {  // send only (never blocks); the row is received and combined at the end of the kernel
        xchg::send_row<D>(a.xch_peers, a.xch_epochs, a.xch_rank, a.xch_world, (int64_t)a.B * a.H, row, O, M * kLn2, L, lane);
        return;
    }
    if (a.part_m) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a.part_o[row * D + lane * VEC + e] = O[e];
        if (lane == 0) {
            a.part_m[row] = M * kLn2;  // natural-log units at the API
            a.part_l[row] = L;
        }
        return;
    }
    const float inv = 1.f / (L + 1e-6f);  // softmax_lut.cpp:224 epsilon (App. A D4)
#pragma unroll
    for (int e = 0; e < VEC; ++e) a.out[row * D + lane * VEC + e] = O[e] * inv;
    if (lane == 0 && a.lse_out) a.lse_out[row] = (L > 0.f) ? (M + log2f(L)) * kLn2 : -INFINITY;
}

// Cold tails of the streaming kernel, kept OUT of line (and only in its TAIL instance): inlined -- or merely present
// as calls -- their registers cost the 16-warp int8 instance (128-register cap) its allocation in the hot loop
// (1.49 -> 2.6 ms at C4).
template <int D>
__device__ __noinline__ void tail_merge_row(const DecodeArgs& a, int64_t s0, int nc, int64_t row, int lane) {
    float M, L, O[D / 32];
    merge_row_warp<D>(a, s0, nc, lane, M, L, O);
    emit_merged_row<D>(a, row, lane, M, L, O);
}
template <int D>
__device__ __noinline__ void tail_empty_row(const DecodeArgs& a, int64_t row, int lane) {
    float O[D / 32];
#pragma unroll
    for (int e = 0; e < D / 32; ++e) O[e] = 0.f;
    emit_merged_row<D>(a, row, lane, -INFINITY, 0.f, O);
}
template <int D>
__device__ __noinline__ void tail_recv_rows(const DecodeArgs& a, int64_t first, int64_t stride, int lane) {
    const int64_t nrows = (int64_t)a.B * a.H;
    for (int64_t row = first; row < nrows; row += stride)
        xchg::recv_row<D>(a.xch_peers, a.xch_epochs, a.xch_rank, a.xch_world, nrows, row, a.out, a.lse_out, a.xch_status,
                          lane);
}

// ------------------------------------------------------------------ direct (a1)
template <int D, int KV>
__global__ void __launch_bounds__(128) paged_decode_direct_kernel(const DecodeArgs a) {
    using C = Cfg<D, KV>;
    constexpr int NW = 4;
    __shared__ float red[NW * (D + 2)];
    // A grid launched with programmatic stream serialisation after this one (splitkv_merge_exchange_kernel) may
    // become resident as soon as every CTA here has started; it waits for this grid's completion itself.
    asm volatile("griddepcontrol.launch_dependents;");
    const int64_t row = blockIdx.x;
    const int split = blockIdx.y;
    const int b = (int)(row / a.H), h = (int)(row % a.H);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, c = lane & 7;
    const int ctx = row_ctx(a, b);
    const int units = (ctx + kUnitTok - 1) / kUnitTok;
    const int per = (units + a.num_splits - 1) / a.num_splits;
    const int u0 = split * per;
    const int u1 = min(units, u0 + per);
    const int beam = a.beam_ids ? a.beam_ids[b] : b;
    const int upt = a.tile_size / kUnitTok;

    float q[C::E];
    const float qoff = load_q<D, KV>(a, row, c, q);
    Acc<D, KV> acc;
    acc.reset();

    for (int u = u0 + warp; u < u1; u += NW) {
        const int tile = u / upt, sub = u - tile * upt;
        const int page = lookup_page(a, beam, h, tile);
        if (page < 0) continue;  // ...fused.cu:32 / cpu_attention_kernel.cpp:73
        const int nvalid = min(kUnitTok, ctx - u * kUnitTok);
        const int64_t tok0 = (int64_t)page * a.tile_size + sub * kUnitTok;
        const uint8_t* kb = a.k_pool + tok0 * C::ROWB;
        const uint8_t* vb = a.v_pool + tok0 * C::ROWB;
        uint32_t kf[4][C::W], vf[4][C::W];
        float ksc[4] = {1.f, 1.f, 1.f, 1.f}, vsc[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int t = 4 * p + g;
#pragma unroll
            for (int v = 0; v < C::NV; ++v) {
                const uint8_t* ka = kb + t * C::ROWB + C::byte_off(c, v);
                const uint8_t* va = vb + t * C::ROWB + C::byte_off(c, v);
                if (C::VB == 16) {
                    uint4 x = ldg_stream_128(ka), y = ldg_stream_128(va);
                    kf[p][v * 4 + 0] = x.x; kf[p][v * 4 + 1] = x.y; kf[p][v * 4 + 2] = x.z; kf[p][v * 4 + 3] = x.w;
                    vf[p][v * 4 + 0] = y.x; vf[p][v * 4 + 1] = y.y; vf[p][v * 4 + 2] = y.z; vf[p][v * 4 + 3] = y.w;
                } else {
                    uint2 x = ldg_stream_64(ka), y = ldg_stream_64(va);
                    kf[p][v * 2 + 0] = x.x; kf[p][v * 2 + 1] = x.y;
                    vf[p][v * 2 + 0] = y.x; vf[p][v * 2 + 1] = y.y;
                }
            }
            if (KV == 1) {
                ksc[p] = __frcp_rn(__ldg(a.k_scales + tok0 + t));
                vsc[p] = __frcp_rn(__ldg(a.v_scales + tok0 + t));
            }
        }
        unit_update<D, KV>(kf, vf, ksc, vsc, q, qoff, nvalid, g, acc);
    }

    const bool final_row = (a.num_splits == 1) && !a.xch_peers;  // across GPUs the row always goes through the workspace
    cta_merge_emit<D, KV, NW>(red, acc, warp, lane, a, final_row ? kEmitFinal : kEmitWorkspace, row,
                              row * a.num_splits + split);
}

// ---------------------------------------------------------------- overlap (a2)
// Work is cut into CHUNKS of `cu` consecutive 16-token units of one row.  Chunks are
// ordered b-major, then h, then position; row (b, h) owns nchunks(b) = ceil(units(b) / cu)
// consecutive chunk ids.  Every WARP is an independent streaming engine: it pulls chunk ids
// from a global atomic counter, keeps its own ring of S TMA bulk-copy stages in flight
// (running up to S units ahead, across chunk boundaries), and writes one (m, l, O) partial
// per chunk -- or the final row when the row is a single chunk.  No block barrier after
// the prologue; load balance is dynamic.
struct ChunkMap {
    const int* prefix;  // shared: prefix[b] = sum_{b'<b} nchunks(b'); null => uniform
    int B, H, NC, cu;   // NC: chunks per row when uniform
    __device__ __forceinline__ int64_t total() const {
        return prefix ? (int64_t)prefix[B] * H : (int64_t)B * H * NC;
    }
    __device__ __forceinline__ int nchunks_of(int b) const { return prefix ? prefix[b + 1] - prefix[b] : NC; }
    __device__ __forceinline__ int64_t row_start(int b, int h) const {
        return prefix ? (int64_t)prefix[b] * H + (int64_t)h * (prefix[b + 1] - prefix[b])
                      : ((int64_t)b * H + h) * NC;
    }
    __device__ __forceinline__ void locate(int64_t id, int& b, int& h, int& j, int& nc) const {
        if (!prefix) {
            const int64_t per_b = (int64_t)H * NC;
            b = (int)(id / per_b);
            const int rem = (int)(id - (int64_t)b * per_b);
            h = rem / NC;
            j = rem - h * NC;
            nc = NC;
        } else {
            int lo = 0, hi = B;  // largest b with prefix[b]*H <= id
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)prefix[mid] * H <= id) lo = mid; else hi = mid;
            }
            b = lo;
            nc = prefix[b + 1] - prefix[b];
            const int64_t rem = id - (int64_t)prefix[b] * H;
            h = (int)(rem / nc);
            j = (int)(rem - (int64_t)h * nc);
        }
    }
};

__device__ __forceinline__ int units_of_ctx(int ctx) { return (ctx + kUnitTok - 1) / kUnitTok; }

// TAIL: the instance that can merge rows in-kernel / exchange them across GPUs (a.row_done != null).  The plain
// instance carries none of that code (see the note at tail_merge_row).
template <int D, int KV, int NW, int S, bool TAIL>
__global__ void __launch_bounds__(NW * 32, 1) paged_decode_overlap_kernel(const DecodeArgs a, int cu,
                                                                          unsigned int* counter) {
    using C = Cfg<D, KV>;
    constexpr int QN = 16;  // per-warp chunk-id ring (producer runs <= S chunks ahead)
    asm volatile("griddepcontrol.launch_dependents;");  // the merge kernel may become resident (it waits for this grid)
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)NW * S * C::STAGE_BYTES);
    int* meta = reinterpret_cast<int*>(bars + NW * S);
    int64_t* cq = reinterpret_cast<int64_t*>(meta + NW * S + ((NW * S) & 1));
    int* prefix = reinterpret_cast<int*>(cq + NW * QN);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, c = lane & 7;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NW * S; ++i) mbar_init(smem_u32(bars + i), 1);
        mbar_fence_init();
    }
    ChunkMap cm;
    cm.B = a.B;
    cm.H = a.H;
    cm.cu = cu;
    cm.NC = (units_of_ctx(row_ctx(a, 0)) + cu - 1) / cu;
    cm.prefix = nullptr;
    if (a.ctx_lens) {
        for (int i = threadIdx.x; i < a.B; i += blockDim.x)
            prefix[i + 1] = (units_of_ctx(row_ctx(a, i)) + cu - 1) / cu;
        __syncthreads();
        if (threadIdx.x == 0) {
            prefix[0] = 0;
            for (int i = 0; i < a.B; ++i) prefix[i + 1] += prefix[i];
        }
        cm.prefix = prefix;
    }
    __syncthreads();

    const int64_t total = cm.total();
    const int64_t total_warps = (int64_t)gridDim.x * NW;
    const int64_t gw = (int64_t)blockIdx.x * NW + warp;
    // kernel parameters used per unit, hoisted out of the constant bank
    const int tile_size = a.tile_size, upt = tile_size / kUnitTok, Hh = a.H, num_tiles = a.num_tiles;
    const int total_pages = a.total_pages, num_beams = a.num_beams;
    const uint8_t* const k_pool = a.k_pool;
    const uint8_t* const v_pool = a.v_pool;
    const float* const k_scales = a.k_scales;
    const float* const v_scales = a.v_scales;
    const bool evict_first = a.evict_first != 0;
    const uint64_t policy = l2_policy_evict_first();

    const uint32_t my_stage0 = smem_u32(stage_base + (size_t)warp * S * C::STAGE_BYTES);
    const uint32_t my_bar0 = smem_u32(bars + warp * S);
    int* my_meta = meta + warp * S;
    int64_t* my_cq = cq + warp * QN;

    // ---- producer state (warp-uniform; lane 0 issues).  Everything that changes per unit is kept
    // incrementally (no divisions / 64-bit index math in the steady state). ----
    bool p_has = false, p_first = true;
    const int32_t* p_trow = nullptr;  // page-table row of (beam, head); null when the beam is out of range
    int p_tile = 0, p_sub = 0, p_rem = 0, p_left = 0, p_page = -1;
    uint32_t p_st = 0, p_chunks = 0;

    auto p_fetch = [&]() {
        int page = p_trow ? __ldg(p_trow + p_tile) : -1;  // page_table.hpp:44-49
        p_page = ((unsigned)page < (unsigned)total_pages) ? page : -1;  // kv_tile_cache.hpp:23
    };
    auto p_next_chunk = [&]() -> bool {
        int64_t id;
        if (p_first) {
            id = gw;  // first chunk is static: no atomic on the critical path of the prologue
            p_first = false;
        } else {
            if (total <= total_warps) return false;  // one static chunk per warp: nothing to fetch
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(counter, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            id = total_warps + (int64_t)t;
        }
        if (id >= total) return false;
        int pb, ph, j, nc;
        cm.locate(id, pb, ph, j, nc);
        const int ctx = row_ctx(a, pb);
        const int u0 = j * cu;
        p_left = min(units_of_ctx(ctx), u0 + cu) - u0;
        p_rem = ctx - u0 * kUnitTok;
        p_tile = u0 / upt;
        p_sub = u0 - p_tile * upt;
        const int beam = a.beam_ids ? a.beam_ids[pb] : pb;
        p_trow = ((unsigned)beam < (unsigned)num_beams) ? a.table + ((int64_t)beam * Hh + ph) * num_tiles : nullptr;
        if (lane == 0) my_cq[p_chunks % QN] = id;
        ++p_chunks;
        __syncwarp();
        return true;
    };

    p_has = p_next_chunk();
    if (p_has) p_fetch();

    auto produce = [&]() {
        if (!p_has) return;
        // elect.sync, not `lane == 0`: the bulk copies keep their operands in uniform registers and are issued
        // back to back (under lane == 0 ptxas wraps every UBLKCP in an R2UR + BRA.U.ANY loop; tests/test_sass.py)
        if (elect_one()) {
            const int nvalid = (p_page >= 0) ? min(kUnitTok, p_rem) : 0;
            my_meta[p_st] = nvalid;
            const uint32_t bar = my_bar0 + p_st * 8;
            if (nvalid > 0) {
                const int64_t tok0 = (int64_t)p_page * tile_size + p_sub * kUnitTok;
                const uint32_t dst = my_stage0 + p_st * C::STAGE_BYTES;
                fence_proxy_async();
                mbar_arrive_expect_tx(bar, C::STAGE_BYTES);
                if (evict_first) {
                    bulk_g2s(dst, k_pool + tok0 * C::ROWB, C::UNIT_BYTES, bar, policy);
                    bulk_g2s(dst + C::UNIT_BYTES, v_pool + tok0 * C::ROWB, C::UNIT_BYTES, bar, policy);
                } else {
                    bulk_g2s_nohint(dst, k_pool + tok0 * C::ROWB, C::UNIT_BYTES, bar);
                    bulk_g2s_nohint(dst + C::UNIT_BYTES, v_pool + tok0 * C::ROWB, C::UNIT_BYTES, bar);
                }
                if (KV == 1) {
                    bulk_g2s_nohint(dst + 2 * C::UNIT_BYTES, k_scales + tok0, C::SCALE_BYTES, bar);
                    bulk_g2s_nohint(dst + 2 * C::UNIT_BYTES + C::SCALE_BYTES, v_scales + tok0, C::SCALE_BYTES, bar);
                }
            } else {
                mbar_arrive(bar);
            }
        }
        p_st = (p_st + 1 == S) ? 0u : p_st + 1;
        p_rem -= kUnitTok;
        if (++p_sub == upt) {
            p_sub = 0;
            ++p_tile;
        }
        if (--p_left == 0) p_has = p_next_chunk();
        if (p_has) p_fetch();
    };

#pragma unroll 1
    for (int s = 0; s < S; ++s) produce();

    // ---- consumer ----
    uint32_t c_st = 0, c_par = 0, c_chunks = 0;
    Acc<D, KV> acc;
    float q[C::E];

    while (c_chunks < p_chunks) {  // the producer is always >= 1 chunk ahead unless it is done
        const int64_t id = my_cq[c_chunks % QN];
        ++c_chunks;
        int b, h, j, nc;
        cm.locate(id, b, h, j, nc);
        const int64_t row = (int64_t)b * Hh + h;
        const int u0 = j * cu;
        const int u1 = min(units_of_ctx(row_ctx(a, b)), u0 + cu);
        const float qoff = load_q<D, KV>(a, row, c, q);
        acc.reset();
#pragma unroll 1
        for (int u = u0; u < u1; ++u) {
            mbar_wait(my_bar0 + c_st * 8, c_par);
            const int nvalid = my_meta[c_st];
            if (nvalid > 0) {
                const uint32_t sb = my_stage0 + c_st * C::STAGE_BYTES;
                uint32_t kf[4][C::W], vf[4][C::W];
                float ksc[4] = {1.f, 1.f, 1.f, 1.f}, vsc[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int t = 4 * p + g;
#pragma unroll
                    for (int v = 0; v < C::NV; ++v) {
                        const uint32_t ka = sb + t * C::ROWB + C::byte_off(c, v);
                        const uint32_t va = ka + C::UNIT_BYTES;
                        if (C::VB == 16) {
                            uint4 x = lds_128(ka), y = lds_128(va);
                            kf[p][v * 4 + 0] = x.x; kf[p][v * 4 + 1] = x.y; kf[p][v * 4 + 2] = x.z; kf[p][v * 4 + 3] = x.w;
                            vf[p][v * 4 + 0] = y.x; vf[p][v * 4 + 1] = y.y; vf[p][v * 4 + 2] = y.z; vf[p][v * 4 + 3] = y.w;
                        } else {
                            uint2 x = lds_64(ka), y = lds_64(va);
                            kf[p][v * 2 + 0] = x.x; kf[p][v * 2 + 1] = x.y;
                            vf[p][v * 2 + 0] = y.x; vf[p][v * 2 + 1] = y.y;
                        }
                    }
                    if (KV == 1) {
                        float ks, vs;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ks) : "r"(sb + 2 * C::UNIT_BYTES + t * 4));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(vs) : "r"(sb + 2 * C::UNIT_BYTES + C::SCALE_BYTES + t * 4));
                        ksc[p] = fast_rcp(ks);
                        vsc[p] = fast_rcp(vs);
                    }
                }
                unit_update<D, KV>(kf, vf, ksc, vsc, q, qoff, nvalid, g, acc);
            }
            __syncwarp();
            if (++c_st == S) {
                c_st = 0;
                c_par ^= 1u;
            }
            produce();  // refill the stage just drained (may pull the next chunk id)
        }
        // Chunk done: warp-level merge, then lanes 0..7 write their dim chunks.
        warp_merge<D, KV>(acc);
        const bool final_row = (nc == 1) && !a.xch_peers;  // across GPUs every row goes through the workspace
        if (final_row) {
            if (lane < 8) {
                if (!a.part_m) {
                    const float inv = 1.f / (acc.l + 1e-6f);
#pragma unroll
                    for (int e = 0; e < C::E; ++e) a.out[row * D + C::dim_of(lane, e)] = acc.o[e] * inv;
                    if (lane == 0 && a.lse_out)
                        a.lse_out[row] = (acc.l > 0.f) ? (acc.m + log2f(acc.l)) * kLn2 : -INFINITY;
                } else {
#pragma unroll
                    for (int e = 0; e < C::E; ++e) a.part_o[row * D + C::dim_of(lane, e)] = acc.o[e];
                    if (lane == 0) {
                        a.part_m[row] = acc.m * kLn2;
                        a.part_l[row] = acc.l;
                    }
                }
            }
            continue;
        }
        if (lane < 8) {
#pragma unroll
            for (int e = 0; e < C::E; ++e) a.ws_o[id * D + C::dim_of(lane, e)] = acc.o[e];
            if (lane == 0) {
                a.ws_m[id] = acc.m;
                a.ws_l[id] = acc.l;
            }
        }
        if (TAIL && a.row_done) {
            // The warp that completes the LAST chunk of a row merges the row here (no second launch): partial
            // visible device-wide -> count -> the last arrival reads all nc partials back from L2.
            __threadfence();
            __syncwarp();
            unsigned int prev = 0;
            if (lane == 0) prev = atomicAdd(a.row_done + row, 1u);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev + 1u == (unsigned)nc) {
                __threadfence();
                if (lane == 0) a.row_done[row] = 0u;  // self-resetting: counters handed over zeroed stay zeroed
                tail_merge_row<D>(a, id - j, nc, row, lane);
            }
        }
    }
    // Rows without a single chunk (context length 0) are emitted as "no keys": out = 0, lse = -inf.
    if (TAIL && a.row_done && (cm.prefix || cm.NC == 0)) {
        const int64_t nrows0 = (int64_t)a.B * Hh;
        for (int64_t row = gw; row < nrows0; row += total_warps) {
            if (cm.nchunks_of((int)(row / Hh)) == 0) tail_empty_row<D>(a, row, lane);
        }
    }
    // ---- split-KV across GPUs: receive + combine.  Every row's partial has been (or will be) SENT by the warp
    // that merged it, on this rank and on the peers; sends never block, so polling here cannot deadlock. ----
    if (TAIL && a.xch_peers) tail_recv_rows<D>(a, gw, total_warps, lane);
}

// ---- split-KV grid kernel across GPUs: merge + exchange as a PROGRAMMATICALLY DEPENDENT launch ----------------------
// One CTA per row, launched with programmatic stream serialisation right behind paged_decode_direct_kernel: its CTAs
// are resident before the decode grid has finished (the launch latency and ramp of a second kernel disappear) and
// block in griddepcontrol.wait until the split partials are complete and visible.  Warp w merges splits
// [16w, 16w + 16) (two L2 round trips), the four results meet in shared memory, warp 0 sends the row to every peer
// and receives / combines the peers' rows.
template <int D>
__global__ void __launch_bounds__(128) splitkv_merge_exchange_kernel(const DecodeArgs a) {
    constexpr int VEC = D / 32;
    __shared__ float red[4 * (D + 2)];
    const int64_t row = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int s_lo = warp * 16, n_mine = max(0, min(16, a.num_splits - s_lo));
    float M, L, O[VEC];
    merge_row_warp<D>(a, row * a.num_splits + s_lo, n_mine, lane, M, L, O);
#pragma unroll
    for (int e = 0; e < VEC; ++e) red[warp * (D + 2) + lane * VEC + e] = O[e];
    if (lane == 0) {
        red[warp * (D + 2) + D] = M;
        red[warp * (D + 2) + D + 1] = L;
    }
    __syncthreads();
    if (warp != 0) return;
    float Mg = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) Mg = fmaxf(Mg, red[w * (D + 2) + D]);
    float Lg = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) O[e] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const float mw = red[w * (D + 2) + D];
        const float wt = (mw == -INFINITY) ? 0.f : fast_exp2(mw - Mg);
        Lg = fmaf(red[w * (D + 2) + D + 1], wt, Lg);
#pragma unroll
        for (int e = 0; e < VEC; ++e) O[e] = fmaf(red[w * (D + 2) + lane * VEC + e], wt, O[e]);
    }
    const int64_t nrows = (int64_t)a.B * a.H;
    xchg::send_row<D>(a.xch_peers, a.xch_epochs, a.xch_rank, a.xch_world, nrows, row, O, Mg * kLn2, Lg, lane);
    xchg::recv_row<D>(a.xch_peers, a.xch_epochs, a.xch_rank, a.xch_world, nrows, row, a.out, a.lse_out, a.xch_status, lane);
}

// Merge the per-chunk partials of every row (rows of a single chunk were finished by the main
// kernel unless the exchange is on; rows with no keys are zeroed).  A CTA of 8 warps handles
// 8 / WPR rows: WPR warps share a row's chunks (strided), each lane owns D/32 output dims, chunk
// weights are computed by lanes in parallel and broadcast with shuffles, so a row of 256 chunks
// (C5: 16K tokens per GPU) costs ~32 dependent 512-byte loads per warp instead of 512 serial
// scalar loads per thread.  Used by the split-KV grid kernel and the beam-group kernel; the streaming kernel
// merges its rows itself (merge_row_warp).
template <int D>
__global__ void __launch_bounds__(256) combine_chunks_kernel(const DecodeArgs a, int cu, int wpr) {
    constexpr int VEC = D / 32;
    extern __shared__ int prefix_sm[];
    __shared__ float red[8][D + 2];
    // launched with programmatic stream serialisation behind the kernel that writes the partials: resident early,
    // blocks here until that grid has completed and its writes are visible (a no-op for an ordinary launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    ChunkMap cm;
    cm.B = a.B;
    cm.H = a.H;
    cm.cu = cu;
    cm.NC = (units_of_ctx(row_ctx(a, 0)) + cu - 1) / cu;
    cm.prefix = nullptr;
    if (a.num_splits > 1) {
        cm.NC = a.num_splits;  // split-KV grid kernel: every row owns num_splits slots (empty splits hold m = -inf)
    } else if (a.row_prefix) {
        cm.prefix = a.row_prefix;
    } else if (a.ctx_lens) {
        if (threadIdx.x == 0) {
            prefix_sm[0] = 0;
            for (int i = 0; i < a.B; ++i)
                prefix_sm[i + 1] = prefix_sm[i] + (units_of_ctx(row_ctx(a, i)) + cu - 1) / cu;
        }
        __syncthreads();
        cm.prefix = prefix_sm;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rpc = 8 / wpr;
    const int wsub = warp % wpr;
    const int64_t nrows = (int64_t)a.B * a.H;
    const int64_t row = (int64_t)blockIdx.x * rpc + warp / wpr;
    const bool row_ok = row < nrows;
    int nc = 0;
    int64_t s0 = 0;
    if (row_ok) {
        const int b = (int)(row / a.H), h = (int)(row % a.H);
        nc = cm.nchunks_of(b);
        s0 = cm.row_start(b, h);
    }
    const bool skip = !row_ok || (nc == 1 && !a.all_rows_in_ws);  // nc == 1: written by the main kernel
    // chunks of this warp: j = wsub + wpr * i, i < n_mine
    const int n_mine = (skip || nc <= wsub) ? 0 : (nc - wsub + wpr - 1) / wpr;
    float mloc = -INFINITY;
    for (int i = lane; i < n_mine; i += 32) mloc = fmaxf(mloc, a.ws_m[s0 + wsub + (int64_t)wpr * i]);
    const float Mw = warp_max(mloc);
    float Lw = 0.f;
    float O[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) O[e] = 0.f;
    for (int base = 0; base < n_mine; base += 32) {
        const int i = base + lane;
        float wt = 0.f;
        if (i < n_mine) {
            const int64_t j = s0 + wsub + (int64_t)wpr * i;
            const float mj = a.ws_m[j];
            wt = (mj == -INFINITY) ? 0.f : fast_exp2(mj - Mw);
            Lw = fmaf(a.ws_l[j], wt, Lw);
        }
        const int cnt = min(32, n_mine - base);
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            const float w = __shfl_sync(0xffffffffu, wt, t);
            const int64_t j = s0 + wsub + (int64_t)wpr * (base + t);
            const float* src = a.ws_o + j * D + lane * VEC;
            if (VEC == 4) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(src));
                O[0] = fmaf(v.x, w, O[0]); O[1] = fmaf(v.y, w, O[1]);
                O[2 % VEC] = fmaf(v.z, w, O[2 % VEC]); O[3 % VEC] = fmaf(v.w, w, O[3 % VEC]);
            } else {
                const float2 v = __ldcs(reinterpret_cast<const float2*>(src));
                O[0] = fmaf(v.x, w, O[0]); O[1] = fmaf(v.y, w, O[1]);
            }
        }
    }
    Lw = warp_sum(Lw);
    float M = Mw, L = Lw;
    if (wpr > 1) {  // cross-warp merge (uniform branch: wpr is a kernel argument)
#pragma unroll
        for (int e = 0; e < VEC; ++e) red[warp][lane * VEC + e] = O[e];
        if (lane == 0) {
            red[warp][D] = Mw;
            red[warp][D + 1] = Lw;
        }
        __syncthreads();
        if (wsub == 0) {
            const int w0 = warp;
            M = -INFINITY;
            for (int w = 0; w < wpr; ++w) M = fmaxf(M, red[w0 + w][D]);
            L = 0.f;
#pragma unroll
            for (int e = 0; e < VEC; ++e) O[e] = 0.f;
            for (int w = 0; w < wpr; ++w) {
                const float mw = red[w0 + w][D];
                const float wt = (mw == -INFINITY) ? 0.f : fast_exp2(mw - M);
                L = fmaf(red[w0 + w][D + 1], wt, L);
#pragma unroll
                for (int e = 0; e < VEC; ++e) O[e] = fmaf(red[w0 + w][lane * VEC + e], wt, O[e]);
            }
        }
    }
    if (skip || wsub != 0) return;
    // ---- emit: this warp holds the merged row (M log2-domain, L, O[VEC] for dims lane*VEC..) ----
    if (a.xch_peers) {  // split-KV across GPUs behind the streaming kernel: send this GPU's row, combine the peers' rows
        xchg::send_row<D>(a.xch_peers, a.xch_epochs, a.xch_rank, a.xch_world, nrows, row, O, M * kLn2, L, lane);
        xchg::recv_row<D>(a.xch_peers, a.xch_epochs, a.xch_rank, a.xch_world, nrows, row, a.out, a.lse_out, a.xch_status, lane);
        return;
    }
    if (a.part_m) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a.part_o[row * D + lane * VEC + e] = O[e];
        if (lane == 0) {
            a.part_m[row] = M * kLn2;
            a.part_l[row] = L;
        }
        return;
    }
    const float inv = 1.f / (L + 1e-6f);  // softmax_lut.cpp:224 epsilon (App. A D4)
#pragma unroll
    for (int e = 0; e < VEC; ++e) a.out[row * D + lane * VEC + e] = O[e] * inv;
    if (lane == 0 && a.lse_out) a.lse_out[row] = (L > 0.f) ? (M + log2f(L)) * kLn2 : -INFINITY;
}

// Public LSE combine (natural-log m), n_parts x rows layout.
__global__ void lse_combine_kernel(const float* __restrict__ pm, const float* __restrict__ pl,
                                   const float* __restrict__ po, int n_parts, int64_t rows, int D,
                                   float* __restrict__ out, float* __restrict__ lse_out) {
    const int64_t row = blockIdx.x;
    float M = -INFINITY;
    for (int i = 0; i < n_parts; ++i) M = fmaxf(M, pm[i * rows + row]);
    float L = 0.f;
    for (int i = 0; i < n_parts; ++i) {
        const float m = pm[i * rows + row];
        L += (m == -INFINITY) ? 0.f : pl[i * rows + row] * __expf(m - M);
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float O = 0.f;
        for (int i = 0; i < n_parts; ++i) {
            const float m = pm[i * rows + row];
            const float w = (m == -INFINITY) ? 0.f : __expf(m - M);
            O = fmaf(po[(i * rows + row) * D + d], w, O);
        }
        out[row * D + d] = O / (L + 1e-6f);
    }
    if (threadIdx.x == 0 && lse_out) lse_out[row] = (L > 0.f) ? M + logf(L) : -INFINITY;
}

// ------------------------------------------------------- beam-group kernel (C3)
// Beams of a group share the page ids of their common prefix (beam_ids -> table rows,
// ...fused.cu:22).  Here a GROUP of W <= 4 consecutive rows is the unit of work: for every 16-token
// unit the W page ids are compared; rows with equal ids share ONE staged copy of the K/V unit
// (the bytes cross HBM once per group instead of once per beam), private / copy-on-write pages are
// staged per distinct id.  With W queries per K/V byte the op is a skinny GEMM, so it runs on the
// tensor cores: mma.sync.m16n8k16 (fp16 in, fp32 accumulate) with
//     S[16 x 16 tok] = Qs[16 x 128] . K^T      rows 0-3 = fp16 hi part of the W queries,
//     O[16 x 128]   += P[16 x 16 tok] . V      rows 4-7 = fp16 lo part (q - hi), rows 8-15 unused
// so hi+lo rows summed give fp32-accurate scores from fp16 tensor-core operands; P is split into
// hi/lo rows the same way.  K/V units are staged by TMA tensor copies (cp.async.bulk.tensor.2d,
// SWIZZLE_128B, boxes of 16 tokens x 64 dims) into per-warp rings so ldmatrix is conflict-free.
// Each chunk writes W (m, l, O) partials into the same workspace slots the overlap kernel uses;
// combine_chunks_kernel merges them.
constexpr int kGroupMaxW = 4;
constexpr int kGroupStageBytes = 8192;  // K lo/hi + V lo/hi boxes of [16][64] fp16

struct GroupArgs {
    int W;        // rows per group
    int groups;   // B / W
};

template <int NW, int S>
__global__ void __launch_bounds__(NW * 32, 1)
paged_decode_group_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const DecodeArgs a, const GroupArgs ga, int cu, unsigned int* counter) {
    constexpr int D = 128;
    constexpr int QN = 16;
    asm volatile("griddepcontrol.launch_dependents;");  // the merge kernel may become resident (it waits for this grid)
    extern __shared__ __align__(1024) uint8_t smem_g[];
    const uint32_t smem_base = (smem_u32(smem_g) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_g + (smem_base - smem_u32(smem_g));
    uint64_t* bars = reinterpret_cast<uint64_t*>(gen_base + (size_t)NW * S * kGroupStageBytes);
    int* meta = reinterpret_cast<int*>(bars + NW * S);
    int64_t* cq = reinterpret_cast<int64_t*>(meta + NW * S + ((NW * S) & 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g8 = lane >> 2, j4 = lane & 3;
    const int W = ga.W;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NW * S; ++i) mbar_init(smem_u32(bars + i), 1);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV));
    }
    __syncthreads();

    // chunk space: (group, head, j).  Uniform context (a.ctx_lens == NULL): NC chunks of `cu` units per row.
    // Ragged (per-row context lengths, e.g. causal prefill where the W rows of a group are consecutive
    // query positions): row b has row_prefix[b+1]-row_prefix[b] chunks, group g as many as its longest row.
    const bool ragged = a.row_prefix != nullptr;
    const int units = units_of_ctx(row_ctx(a, 0));
    const int NC = (units + cu - 1) / cu;
    const int64_t total = ragged ? (int64_t)a.group_prefix[ga.groups] * a.H : (int64_t)ga.groups * a.H * NC;
    const int64_t total_warps = (int64_t)gridDim.x * NW;
    const int64_t gw = (int64_t)blockIdx.x * NW + warp;
    const int upt = a.tile_size / kUnitTok;
    // chunk id -> (group, head, chunk index, chunks of the group)
    auto locate_group = [&](int64_t id, int& g_, int& h_, int& j_, int& ncg) {
        if (!ragged) {
            // position-major, LAST chunks first: the tail of a beam's context is private (W staged copies
            // per unit) while the prefix is shared (one copy), so the expensive chunks are dispatched first
            // and the cheap ones fill the end of the run (longest-first balance).
            const int64_t per_j = (int64_t)ga.groups * a.H;
            j_ = NC - 1 - (int)(id / per_j);
            const int rem = (int)(id % per_j);
            g_ = rem / a.H;
            h_ = rem - g_ * a.H;
            ncg = NC;
        } else {
            int lo = 0, hi = ga.groups;  // largest g with group_prefix[g] * H <= id
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)a.group_prefix[mid] * a.H <= id) lo = mid; else hi = mid;
            }
            g_ = lo;
            ncg = a.group_prefix[lo + 1] - a.group_prefix[lo];
            const int64_t rem = id - (int64_t)a.group_prefix[lo] * a.H;
            h_ = (int)(rem / ncg);
            j_ = ncg - 1 - (int)(rem - (int64_t)h_ * ncg);
        }
    };

    const uint32_t my_stage0 = smem_base + (uint32_t)warp * S * kGroupStageBytes;
    const uint32_t my_bar0 = smem_u32(bars + warp * S);
    int* my_meta = meta + warp * S;
    int64_t* my_cq = cq + warp * QN;

    // ---------------- producer (warp-uniform state; lanes < W look pages up in parallel) -------
    bool p_has = false, p_first = true;
    int p_g = 0, p_h = 0, p_u = 0, p_uend = 0, p_ctx = 0;
    int p_beam = 0;          // lane w < W: table row of beam w of the current group
    uint32_t issued = 0, p_chunks = 0;
    int p_pending_mask = 0;  // beams of the current unit not yet staged
    int p_page = -1;         // lane w: page of beam w for the current unit

    auto p_next_chunk = [&]() -> bool {
        int64_t id;
        if (p_first) {
            id = gw;
            p_first = false;
        } else {
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(counter, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            id = total_warps + (int64_t)t;
        }
        if (id >= total) return false;
        int j, ncg;
        locate_group(id, p_g, p_h, j, ncg);
        int cw = 0;
        if (lane < W) {
            const int b = p_g * W + lane;
            p_beam = a.beam_ids ? a.beam_ids[b] : b;
            cw = row_ctx(a, b);
        }
        p_ctx = (int)warp_max((float)cw);  // the group's longest context (exact: < 2^24 tokens)
        p_u = j * cu;
        p_uend = min(units_of_ctx(p_ctx), p_u + cu);
        if (lane == 0) my_cq[p_chunks % QN] = id;
        ++p_chunks;
        __syncwarp();
        return true;
    };
    auto p_fetch_unit = [&]() {  // page ids of the W beams for unit p_u
        p_page = (lane < W) ? lookup_page(a, p_beam, p_h, p_u / upt) : -1;
        p_pending_mask = (int)(__ballot_sync(0xffffffffu, lane < W && p_page >= 0));
    };
    p_has = p_next_chunk();
    if (p_has) p_fetch_unit();

    // Stage ONE copy: the lowest pending beam's page, shared by every pending beam with the same id.
    auto produce = [&]() {
        if (!p_has) return;
        const uint32_t st = issued % S;
        const uint32_t bar = my_bar0 + st * 8;
        int mask = 0, page = -1;
        if (p_pending_mask) {
            const int lead = __ffs(p_pending_mask) - 1;
            page = __shfl_sync(0xffffffffu, p_page, lead);
            mask = (int)(__ballot_sync(0xffffffffu, lane < W && p_page == page)) & p_pending_mask;
            p_pending_mask &= ~mask;
        }
        const bool last_of_unit = (p_pending_mask == 0);
        const bool last_of_chunk = last_of_unit && (p_u + 1 >= p_uend);
        const int nvalid = mask ? min(kUnitTok, p_ctx - p_u * kUnitTok) : 0;
        if (elect_one()) {
            // meta: [0,5) tokens of the unit inside the group's context, [5,13) beams sharing this copy,
            // [13] last stage of the chunk, [14,32) unit index (per-beam context limits in ragged mode)
            my_meta[st] = nvalid | (mask << 5) | (last_of_chunk ? (1 << 13) : 0) | (p_u << 14);
            if (nvalid > 0) {
                const int row0 = page * a.tile_size + (p_u % upt) * kUnitTok;  // token row in the pool
                const uint32_t dst = my_stage0 + st * kGroupStageBytes;
                fence_proxy_async();
                mbar_arrive_expect_tx(bar, kGroupStageBytes);
                tma_load_2d(dst, &tmK, 0, row0, bar);
                tma_load_2d(dst + 2048, &tmK, 64, row0, bar);
                tma_load_2d(dst + 4096, &tmV, 0, row0, bar);
                tma_load_2d(dst + 6144, &tmV, 64, row0, bar);
            } else {
                mbar_arrive(bar);
            }
        }
        ++issued;
        if (last_of_unit) {
            ++p_u;
            if (p_u >= p_uend) p_has = p_next_chunk();
            if (p_has) p_fetch_unit();
        }
    };

#pragma unroll 1
    for (int s = 0; s < S; ++s) produce();

    // ---------------- consumer ---------------------------------------------------------------
    uint32_t consumed = 0, c_chunks = 0;
    const int beam_of_lane = g8 & 3;          // MMA row g8: rows 0-3 hi, 4-7 lo of beam g8 & 3
    const bool lo_half = g8 >= 4;
    while (c_chunks < p_chunks) {
        const int64_t id = my_cq[c_chunks % QN];
        ++c_chunks;
        int cg, ch, cj, ncg;
        locate_group(id, cg, ch, cj, ncg);
        // context length of this lane's beam (tokens at or past it are masked: causal limit in prefill)
        const int c_ctx = (beam_of_lane < W) ? row_ctx(a, cg * W + beam_of_lane) : 0;
        // Q fragments: row g8 of the 16 x 128 operand, fp16 hi (rows 0-3) / lo (rows 4-7) parts
        uint32_t qa[8][2];
        {
            const bool valid_beam = beam_of_lane < W;
            const int64_t qrow = ((int64_t)(cg * W + (valid_beam ? beam_of_lane : 0)) * a.H + ch) * D;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int d0 = s * 16 + hh * 8 + j4 * 2;
                    float x0 = 0.f, x1 = 0.f;
                    if (valid_beam) {
                        x0 = a.q[qrow + d0];
                        x1 = a.q[qrow + d0 + 1];
                        if (a.rope) {
                            const float cs = a.rope[d0], sn = a.rope[d0 + 1];
                            const float r0 = x0 * cs - x1 * sn, r1 = x0 * sn + x1 * cs;
                            x0 = r0;
                            x1 = r1;
                        }
                        x0 *= a.qscale;
                        x1 *= a.qscale;
                    }
                    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                    if (lo_half) {
                        x0 -= __half2float(h0);
                        x1 -= __half2float(h1);
                        qa[s][hh] = pack_half2(x0, x1);
                    } else {
                        __half2 hv = __halves2half2(h0, h1);
                        qa[s][hh] = *reinterpret_cast<uint32_t*>(&hv);
                    }
                }
            }
        }
        float m_run = -INFINITY, l_run = 0.f;  // of beam_of_lane (replicated over the 4 j4 lanes: l is per-lane partial)
        float o[16][4];
#pragma unroll
        for (int t = 0; t < 16; ++t) { o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f; }

        bool done = false;
#pragma unroll 1
        while (!done) {
            const uint32_t st = consumed % S;
            mbar_wait(my_bar0 + st * 8, (consumed / S) & 1);
            const int mt = my_meta[st];
            const int nvalid = mt & 0x1f, mask = (mt >> 5) & 0xff;
            done = (mt >> 13) & 1;
            const int lim = c_ctx - (int)((unsigned)mt >> 14) * kUnitTok;  // tokens of this unit inside MY context
            if (nvalid > 0) {
                const uint32_t sb = my_stage0 + st * kGroupStageBytes;
                if (nvalid < kUnitTok) {
                    // rows past the context end may hold anything (even NaN bit patterns): zero the V rows
                    for (int i = lane; i < (kUnitTok - nvalid) * 16; i += 32) {
                        const int r = nvalid + i / 16, c = i % 16;  // 16 chunks of 16 B per 256-byte row (2 boxes)
                        const uint32_t addr = sb + 4096 + (c >> 3) * 2048 + r * 128 + (((c & 7) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
                    }
                    __syncwarp();
                }
                // ---- S = Qs . K^T : 2 n-tiles of 8 tokens ----
                float sacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                {
                    const int mi = lane >> 3;
                    const int r = (lane & 7) + (mi >> 1) * 8;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int c = 2 * (s & 3) + (mi & 1);
                        const uint32_t addr = sb + (s >> 2) * 2048 + r * 128 + ((c ^ (r & 7)) << 4);
                        uint32_t b0, b1, b2, b3;
                        ldmatrix_x4(addr, b0, b1, b2, b3);
                        mma_16816(sacc[0], qa[s][0], qa[s][1], b0, b1);
                        mma_16816(sacc[1], qa[s][0], qa[s][1], b2, b3);
                    }
                }
                // hi + lo rows -> full-precision scores of beam_of_lane for tokens t*8 + j4*2 + {0,1}
                float sc[4];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float v = sacc[t][e];
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        const int tok = t * 8 + j4 * 2 + e;
                        const bool ok = tok < nvalid && tok < lim && ((mask >> beam_of_lane) & 1);
                        sc[t * 2 + e] = ok ? v : -INFINITY;
                    }
                }
                float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float m_new = fmaxf(m_run, mx);
                // beams without this unit (mask bit 0) or with no valid key keep their state: corr = 1
                const float corr = (m_new == -INFINITY) ? 1.f : fast_exp2(m_run - m_new);
                float p[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = (sc[i] == -INFINITY) ? 0.f : fast_exp2(sc[i] - m_new);
                l_run = fmaf(l_run, corr, (p[0] + p[1]) + (p[2] + p[3]));
                m_run = m_new;
                if (__any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll
                    for (int t = 0; t < 16; ++t) { o[t][0] *= corr; o[t][1] *= corr; }
                }
                // P fragments: rows 0-3 hi, rows 4-7 lo
                uint32_t pa0, pa2;
                {
                    const __half2 h01 = __floats2half2_rn(p[0], p[1]), h23 = __floats2half2_rn(p[2], p[3]);
                    if (lo_half) {
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        pa0 = pack_half2(p[0] - f01.x, p[1] - f01.y);
                        pa2 = pack_half2(p[2] - f23.x, p[3] - f23.y);
                    } else {
                        pa0 = *reinterpret_cast<const uint32_t*>(&h01);
                        pa2 = *reinterpret_cast<const uint32_t*>(&h23);
                    }
                }
                // ---- O += P . V : 16 n-tiles of 8 dims ----
                {
                    const int mi = lane >> 3;
                    const int r = (lane & 7) + (mi & 1) * 8;
#pragma unroll
                    for (int n2 = 0; n2 < 8; ++n2) {
                        const int c = 2 * (n2 & 3) + (mi >> 1);
                        const uint32_t addr = sb + 4096 + (n2 >> 2) * 2048 + r * 128 + ((c ^ (r & 7)) << 4);
                        uint32_t b0, b1, b2, b3;
                        ldmatrix_x4_trans(addr, b0, b1, b2, b3);
                        mma_16816(o[2 * n2], pa0, pa2, b0, b1);
                        mma_16816(o[2 * n2 + 1], pa0, pa2, b2, b3);
                    }
                }
            }
            __syncwarp();
            ++consumed;
            produce();
        }
        // ---- chunk done: per-beam (m, l, O) partial -> workspace slot ((b*H + h)*NC + j) ----
        float l_tot = l_run;
        l_tot += __shfl_xor_sync(0xffffffffu, l_tot, 1);
        l_tot += __shfl_xor_sync(0xffffffffu, l_tot, 2);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            o[t][0] += __shfl_xor_sync(0xffffffffu, o[t][0], 16);
            o[t][1] += __shfl_xor_sync(0xffffffffu, o[t][1], 16);
        }
        // workspace slot of (row b, head, chunk j): rows own consecutive slots (ChunkMap::row_start); in ragged
        // mode a row shorter than its group simply has no chunk j.
        const int brow = cg * W + beam_of_lane;
        int nc_b = NC;
        int64_t slot0 = ((int64_t)brow * a.H + ch) * NC;
        if (ragged && beam_of_lane < W) {
            nc_b = a.row_prefix[brow + 1] - a.row_prefix[brow];
            slot0 = (int64_t)a.row_prefix[brow] * a.H + (int64_t)ch * nc_b;
        }
        if (!lo_half && beam_of_lane < W && cj < nc_b) {
            const int64_t slot = slot0 + cj;
            float* dst = a.ws_o + slot * D;
#pragma unroll
            for (int t = 0; t < 16; ++t)
                *reinterpret_cast<float2*>(dst + t * 8 + j4 * 2) = make_float2(o[t][0], o[t][1]);
            if (j4 == 0) {
                a.ws_m[slot] = m_run;
                a.ws_l[slot] = l_tot;
            }
        }
    }
}

// Chunk prefixes for the ragged group kernel, computed once per launch into the workspace:
// row_prefix[b+1] - row_prefix[b] = ceil(units(ctx_b) / cu); group_prefix over max of the W rows of a group.
__global__ void __launch_bounds__(1024) group_prefix_kernel(const DecodeArgs a, int W, int cu, int* __restrict__ row_prefix,
                                                            int* __restrict__ group_prefix) {
    __shared__ int part[1024];
    const int G = a.B / W;
    for (int pass = 0; pass < 2; ++pass) {
        const int n = pass == 0 ? a.B : G;
        int* out = pass == 0 ? row_prefix : group_prefix;
        const int per = (n + 1023) / 1024;
        const int i0 = threadIdx.x * per, i1 = min(n, i0 + per);
        auto count = [&](int i) {
            if (pass == 0) return (units_of_ctx(row_ctx(a, i)) + cu - 1) / cu;
            int m = 0;
            for (int w = 0; w < W; ++w) m = max(m, (units_of_ctx(row_ctx(a, i * W + w)) + cu - 1) / cu);
            return m;
        };
        int s = 0;
        for (int i = i0; i < i1; ++i) s += count(i);
        part[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int t = 0; t < 1024; ++t) {
                const int v = part[t];
                part[t] = run;
                run += v;
            }
            out[n] = run;
        }
        __syncthreads();
        int run = part[threadIdx.x];
        for (int i = i0; i < i1; ++i) {
            out[i] = run;
            run += count(i);
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------------- host side
constexpr int kOvWarps = 8;
template <int D, int KV>
struct OvCfg {
    // fp16: 8 streaming warps x 3 stages of 8 KB.  int8: twice the math per byte (and the kernel is
    // issue/latency bound, not HBM bound), so 16 warps x 3 stages of 4 KB: same bytes in flight, twice
    // the thread-level parallelism.  PA_OV_I8_WARPS overrides at build time for experiments.
#ifndef PA_OV_I8_WARPS
#define PA_OV_I8_WARPS 16
#endif
    // fp32 pools (KV == 2): 16 KB (D = 128) / 8 KB (D = 64) per stage -> 4 / 8 warps x 3 stages = 192 KB.
    static constexpr int NW = KV == 2 ? (D == 128 ? 4 : 8) : ((KV == 1 && D == 128) ? PA_OV_I8_WARPS : 8);
    static constexpr int S = KV == 2 ? 3 : ((D == 128) ? (KV == 0 ? 3 : (NW == 16 ? 3 : 6)) : (KV == 0 ? 6 : 8));
};

// combine_chunks_kernel as a programmatically dependent launch: its launch latency and ramp hide under the tail of the
// producer grid (which executes griddepcontrol.launch_dependents as its first instruction).
template <int D>
static cudaError_t launch_combine_pdl(const DecodeArgs& a, int cu, int wpr, int grid, size_t smem, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, combine_chunks_kernel<D>, a, cu, wpr);
}

// Few rows with a long context each (the C5 family: one sequence's 32 heads) stay on the grid kernel up to 4x that:
// measured at 32 rows x 16K / 24K / 32K / 48K / 64K tokens: 49.2 / 68.4 / 87.3 / 125.3 / 162.7 us (5.5 ... 6.6 TB/s)
// against 57.7 / 79.2 / 102.4 / 133.5 / 172.1 us for the streaming kernel.
static bool grid_kernel_range(int64_t rows, int64_t units_per_row) {
    const int64_t units = rows * units_per_row;
    return units <= 65536 || (rows <= 64 && units <= 262144);
}

static int units_of_ctx_host(int T, int cap) {
    int ctx = T < 0 ? 0 : (T > cap ? cap : T);
    return (ctx + kUnitTok - 1) / kUnitTok;
}

static int choose_splits(int64_t rows, int max_units, int sm_count) {
    if (const char* e = getenv("PA_DECODE_SPLITS")) {   // experiments
        const int v = atoi(e);
        if (v >= 1 && v <= 64) return v;
    }
    // Few rows with a long context each (one sequence split over GPUs: 32 rows x 1024+ units): ONE resident wave (4 CTAs
    // per SM) of long splits beats many short ones -- measured on the 8-GPU C5 share (32 x 1024 units), fused step:
    // 64 splits 48.6 us, 32: 47.4, 16-18: 46.8, 12: 47.7, 8: 53.6, and 20-24 (a second, partial wave) 54-57.  So: as
    // many splits as fit one wave, at least 64 units each; very long rows (> 192 units per split) take two or more waves
    // of 128-unit splits instead (4096 units: 32 splits 160.7 us, 16: 162.4, 64: 162.6).
    if (rows <= 64 && max_units >= 512) {
        int ns = (int)((4 * (int64_t)sm_count) / rows);
        if (ns > max_units / 64) ns = max_units / 64;
        if (ns > 64) ns = 64;
        if (ns >= 1 && (max_units + ns - 1) / ns > 192) ns = max_units / 128 < 64 ? max_units / 128 : 64;
        if (ns >= 1) return ns;
    }
    const int64_t target = (int64_t)sm_count * 15;  // ~5 resident CTAs/SM x 3 waves
    int64_t ns = (target + rows - 1) / rows;
    const int cap_units = max_units / 4 > 0 ? max_units / 4 : 1;  // >= 4 units (1 per warp) per split
    if (ns > cap_units) ns = cap_units;
    if (ns > 64) ns = 64;
    if (ns < 1) ns = 1;
    return (int)ns;
}

// Units per chunk of the overlap kernel (nw = streaming warps per CTA of the instance that will run).
//  * short jobs with few rows (one GPU's share of a long sequence, C5: 32 rows x 1024 units): ONE chunk per
//    warp, handed out statically (the first chunk of a warp is its global warp index) -- no atomics, no
//    per-chunk overheads, every warp streams the same number of units and the row merge happens in-kernel;
//  * otherwise 16 units (128 KiB of fp16 K+V at D=128) when that still yields >= 4 chunks per resident warp,
//    fewer for small problems so every warp gets work; chunks are dispatched dynamically.
static int choose_cu(int64_t rows, int max_units, int sm_count, int nw, bool prefer_static = false) {
    const int64_t warps = (int64_t)sm_count * nw;
    const int64_t total = rows * (int64_t)max_units;
    // prefer_static (inter-GPU exchange: rows are merged in-kernel, which costs a device-wide fence per finished
    // chunk): one chunk per warp whatever the size -- a static split loses some % to SM-to-SM rate differences on long
    // jobs (measured 5 % at 1 GB), less than ~14 fences + a 512-chunk row merge per warp would (25 %).
    const char* st_env = getenv("PA_DECODE_STATIC");  // experiments: 0 = never one static chunk per warp, 1 = whenever possible
    if (st_env) prefer_static = atoi(st_env) == 1;
    if (rows > 0 && rows <= warps && !(st_env && atoi(st_env) == 0) && (prefer_static || total <= warps * 64) && max_units > 0) {
        const int64_t per_row = warps / rows;
        const int64_t cu1 = (max_units + per_row - 1) / per_row;
        const int64_t chunks = rows * ((max_units + cu1 - 1) / cu1);
        if (chunks * 100 >= warps * 85) return (int)cu1;  // >= 85 % of the warps get (exactly) one chunk
    }
    int64_t cu = total / (4 * warps);
    if (cu > 16) cu = 16;
    if (cu < 2) cu = 2;
    int p = 2;
    while (p * 2 <= cu) p *= 2;
    return p;
}

// Units per chunk of the beam-group kernel.  A chunk is expensive to start and finish there (W query fragments in
// fp16 hi / lo, W partial rows written and merged later), so few large chunks win: measured at C3 (1024 (group, head)
// pairs of 128 units, 1184 resident warps) 8-unit chunks 307 us, 16: 272, 32: 260, 64: 249, 128 (ONE chunk per pair,
// a single wave on 86 % of the warps): 243 us.  Rule: split a pair into the largest power-of-two number of chunks
// that still fits ONE wave of warps; with more pairs than warps, the fewest chunks that give >= 3 waves to balance.
static int choose_group_cu(int64_t group_heads, int max_units, int sm_count) {
    const int64_t warps = (int64_t)sm_count * kOvWarps;
    if (group_heads <= 0 || max_units <= 0) return 2;
    int nc = 1;
    if (group_heads <= warps) {
        while ((int64_t)group_heads * nc * 2 <= warps && nc * 2 <= max_units) nc *= 2;
    } else {
        while ((int64_t)group_heads * nc < 3 * warps && nc * 2 <= max_units) nc *= 2;
    }
    int cu = (max_units + nc - 1) / nc;
    if (const char* env = getenv("PA_GROUP_CU")) {
        const int v = atoi(env);
        if (v >= 2) cu = v;
    }
    return cu < 2 ? 2 : cu;
}

static size_t ws_slots(int64_t rows, int max_units, int sm_count) {
    size_t ov = 0;
    for (int nw : {4, 8, 16}) {  // the streaming-kernel instances differ in warps per CTA; size for the largest need
        for (int st = 0; st < 2; ++st) {
            const int cu = choose_cu(rows, max_units, sm_count, nw, st == 1);
            const size_t v = (size_t)rows * ((max_units + cu - 1) / cu);
            if (v > ov) ov = v;
        }
    }
    size_t dr = (size_t)rows * choose_splits(rows, max_units, sm_count);
    // beam-group kernel: its chunk size depends on the number of (group, head) pairs = rows / beam width (1..4)
    size_t gr = 0;
    for (int w = 1; w <= kGroupMaxW; ++w) {
        int cug = choose_group_cu(rows / w > 0 ? rows / w : 1, max_units, sm_count);
        if (getenv("PA_GROUP_CU")) cug = 2;  // experiments may lower the chunk size down to 2 units: size for that
        const size_t v = (size_t)rows * ((max_units + cug - 1) / cug);
        if (v > gr) gr = v;
    }
    size_t mx = ov > dr ? ov : dr;
    if (gr > mx) mx = gr;
    return ((mx + 64) + 3) & ~(size_t)3;  // multiple of 4: ws_o stays 16-byte aligned
}

// Workspace layout: [chunk counter: 256 B][row_done: rows x u32, padded to 256 B][m: nslots][l: nslots][o: nslots * D]
// [row / group chunk prefixes of the ragged group kernel]
static size_t ws_header_bytes(int64_t rows) { return 256 + (((size_t)rows * 4 + 255) & ~(size_t)255); }

static size_t ws_bytes_needed(int B, int H, int D, int num_tiles, int tile_size, int sm_count) {
    const int max_units = (int)(((int64_t)num_tiles * tile_size + kUnitTok - 1) / kUnitTok);
    return ws_slots((int64_t)B * H, max_units, sm_count) * (size_t)(D + 2) * sizeof(float) +
           ws_header_bytes((int64_t)B * H) + 2 * ((size_t)B + 2) * sizeof(int);
}

template <int D, int KV>
static int launch_decode(DecodeArgs& a, bool overlap, void* ws, size_t ws_bytes, cudaStream_t st) {
    const DeviceInfo& di = device_info();
    if (!di.ok) return PA_ERR_NO_DEVICE;
    const int64_t rows = (int64_t)a.B * a.H;
    const int max_units = (int)(((int64_t)a.num_tiles * a.tile_size + kUnitTok - 1) / kUnitTok);
    if (!ws || ws_bytes < ws_bytes_needed(a.B, a.H, D, a.num_tiles, a.tile_size, di.sm_count))
        return PA_ERR_WORKSPACE;
    unsigned int* counter = static_cast<unsigned int*>(ws);
    float* w = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + ws_header_bytes(rows));
    const size_t nslots = ws_slots(rows, max_units, di.sm_count);
    a.ws_m = w;
    a.ws_l = w + nslots;
    a.ws_o = w + 2 * nslots;
    if (a.xch_peers) {
        // Inter-GPU exchange.  In the grid kernel's range (grid_kernel_range: C5 shares up to 1 GB of K/V per GPU) the
        // split-KV grid kernel streams fastest (one GPU's C5 share of 268 MB: 46 us vs 56 us for the streaming kernel),
        // so: grid kernel -> row partials in the workspace -> merge + exchange kernel chained by PROGRAMMATIC
        // DEPENDENT LAUNCH (resident before the decode grid ends, no launch gap).  Above that size: the streaming
        // kernel + the chunk-merge kernel (PDL) that exchanges in its emit step.  PA_PARTIAL_DIRECT=0 / 1 forces either.
        const char* env = getenv("PA_PARTIAL_DIRECT");
        overlap = env ? atoi(env) != 1 : !(grid_kernel_range(rows, max_units) && rows <= 1024);
        if (!overlap) {
            a.num_splits = choose_splits(rows, max_units, di.sm_count);  // <= 64 = 4 warps x 16 in the merge kernel
            dim3 grid((unsigned)rows, (unsigned)a.num_splits);
            paged_decode_direct_kernel<D, KV><<<grid, 128, 0, st>>>(a);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return (int)e;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)rows);
            cfg.blockDim = dim3(128);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            e = cudaLaunchKernelEx(&cfg, splitkv_merge_exchange_kernel<D>, a);
            return e == cudaSuccess ? PA_OK : (int)e;
        }
    }
    if (!overlap) {
        a.num_splits = choose_splits(rows, max_units, di.sm_count);
        dim3 grid((unsigned)rows, (unsigned)a.num_splits);
        paged_decode_direct_kernel<D, KV><<<grid, 128, 0, st>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        if (a.num_splits > 1) {
            a.all_rows_in_ws = 1;
            const int ns = a.num_splits;
            const int wpr = ns >= 64 ? 8 : (ns >= 32 ? 4 : (ns >= 16 ? 2 : 1));
            const int rpc = 8 / wpr;
            e = launch_combine_pdl<D>(a, 1, wpr, (int)((rows + rpc - 1) / rpc), 0, st);
            if (e != cudaSuccess) return (int)e;
        }
        return PA_OK;
    }
    constexpr int S = OvCfg<D, KV>::S;
    using C = Cfg<D, KV>;
    const int G = di.sm_count;
    constexpr int NWk = OvCfg<D, KV>::NW;
    // Across GPUs above the grid kernel's size range: the PLAIN streaming kernel, then the chunk-merge kernel chained by
    // programmatic dependent launch, whose emit step sends / receives / combines the row.  (r02 first merged the rows
    // inside a TAIL instance of the streaming kernel: at 2 GPUs x 1 GB the fence + counter per finished chunk cost 9 us
    // on top of the 8 us exchange -- 193.9 us against 183.4 for partial + stand-alone exchange kernel.
    // PA_DECODE_MERGE_KERNEL=0 still selects that form.)
    const char* merge_env = getenv("PA_DECODE_MERGE_KERNEL");
    const bool xch_tail = a.xch_peers && merge_env && atoi(merge_env) == 0;
    const int cu = choose_cu(rows, max_units, di.sm_count, NWk, xch_tail);
    const size_t prefix_bytes = a.ctx_lens ? (size_t)(a.B + 1) * sizeof(int) : 0;
    const size_t smem = (size_t)NWk * S * C::STAGE_BYTES + (size_t)NWk * S * (8 + 4) + 8 +
                        (size_t)NWk * 16 * sizeof(int64_t) + prefix_bytes;
    if (smem > (size_t)di.max_smem_optin) return PA_ERR_UNSUPPORTED;  // B too large for the prefix table
    // Rows are merged IN-KERNEL (the warp that finishes a row's last chunk merges it) only where that pays.  Measured
    // at C2 (27.7 chunks per warp): the device-wide fence + counter per finished chunk stall a streaming warp ~1.5 us
    // each = 5 % of the launch, more than the second launch costs (9 us of 640).  So: in-kernel where a warp
    // finishes ONE chunk (static mode: a GPU's share of a long sequence, where the second launch is ~10 % of the
    // step) and across GPUs (the exchange needs it); the separate merge kernel otherwise.
    // PA_DECODE_MERGE_KERNEL=0 / 1 forces in-kernel / separate merging where both are possible.
    const bool static_chunks = rows * (int64_t)((max_units + cu - 1) / cu) <= (int64_t)G * NWk;
    const bool fused_merge = a.xch_peers ? xch_tail : (merge_env ? atoi(merge_env) == 0 : static_chunks);
    if (a.xch_peers && !fused_merge) a.all_rows_in_ws = 1;
    a.row_done = fused_merge ? counter + 64 : nullptr;
    // Across GPUs with one static chunk per warp nothing in the workspace header is needed: the chunk counter is not
    // used and the row counters live behind the caller's zero-initialised, self-resetting epoch array (d_epochs
    // [2 * rows]) -- no memset node in front of the kernel (~2 us of a ~50 us step).
    const bool no_memset = xch_tail && static_chunks;
    if (no_memset) a.row_done = a.xch_epochs + rows;
    auto kern = fused_merge ? paged_decode_overlap_kernel<D, KV, NWk, S, true> : paged_decode_overlap_kernel<D, KV, NWk, S, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (!no_memset) e = cudaMemsetAsync(ws, 0, fused_merge ? ws_header_bytes(rows) : sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    kern<<<G, NWk * 32, smem, st>>>(a, cu, counter);
    e = cudaGetLastError();
    if (e != cudaSuccess || fused_merge) return e == cudaSuccess ? PA_OK : (int)e;
    // Rows of a single chunk were finished by the main kernel; everything else is merged here.
    const int nc_uniform = (units_of_ctx_host(a.T, a.num_tiles * a.tile_size) + cu - 1) / cu;
    const bool need_combine = a.ctx_lens != nullptr || nc_uniform != 1 || a.xch_peers != nullptr;
    if (need_combine) {
        const int nc_max = (max_units + cu - 1) / cu;
        const int wpr = nc_max >= 64 ? 8 : (nc_max >= 32 ? 4 : (nc_max >= 16 ? 2 : 1));
        const int rpc = 8 / wpr;
        const int cgrid = (int)((rows + rpc - 1) / rpc);
        e = launch_combine_pdl<D>(a, cu, wpr, cgrid, prefix_bytes, st);
    }
    return e == cudaSuccess ? PA_OK : (int)e;
}

static int launch_group(DecodeArgs& a, int W, void* ws, size_t ws_bytes, cudaStream_t st) {
    constexpr int D = 128, S = 3;
    const DeviceInfo& di = device_info();
    if (!di.ok) return PA_ERR_NO_DEVICE;
    const int64_t rows = (int64_t)a.B * a.H;
    const int max_units = (int)(((int64_t)a.num_tiles * a.tile_size + kUnitTok - 1) / kUnitTok);
    if (!ws || ws_bytes < ws_bytes_needed(a.B, a.H, D, a.num_tiles, a.tile_size, di.sm_count)) return PA_ERR_WORKSPACE;
    unsigned int* counter = static_cast<unsigned int*>(ws);
    float* w = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + ws_header_bytes(rows));
    const size_t nslots = ws_slots(rows, max_units, di.sm_count);
    a.ws_m = w;
    a.ws_l = w + nslots;
    a.ws_o = w + 2 * nslots;
    a.all_rows_in_ws = 1;
    GroupArgs ga{W, a.B / W};
    const int64_t gh = (int64_t)ga.groups * a.H;
    const int cu = choose_group_cu(gh, max_units, di.sm_count);
    CUtensorMap tmK, tmV;
    const uint64_t total_tokens = (uint64_t)a.total_pages * a.tile_size;
    if (!make_pool_map(&tmK, a.k_pool, total_tokens) || !make_pool_map(&tmV, a.v_pool, total_tokens))
        return PA_ERR_UNSUPPORTED;
    cudaError_t e;
    if (a.ctx_lens) {  // ragged: per-row context lengths -> chunk prefixes in the workspace
        int* rp = reinterpret_cast<int*>(w + nslots * (size_t)(D + 2));
        int* gp = rp + a.B + 2;
        group_prefix_kernel<<<1, 1024, 0, st>>>(a, W, cu, rp, gp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        a.row_prefix = rp;
        a.group_prefix = gp;
    }
    const size_t smem = (size_t)kOvWarps * S * kGroupStageBytes + (size_t)kOvWarps * S * (8 + 4) + 8 +
                        (size_t)kOvWarps * 16 * sizeof(int64_t) + 1024;
    auto kern = paged_decode_group_kernel<kOvWarps, S>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    kern<<<di.sm_count, kOvWarps * 32, smem, st>>>(tmK, tmV, a, ga, cu, counter);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int nc_max = (units_of_ctx_host(a.T, a.num_tiles * a.tile_size) + cu - 1) / cu;
    const int wpr = nc_max >= 64 ? 8 : (nc_max >= 32 ? 4 : (nc_max >= 16 ? 2 : 1));
    const int rpc = 8 / wpr;
    e = launch_combine_pdl<D>(a, cu, wpr, (int)((rows + rpc - 1) / rpc), 0, st);
    return e == cudaSuccess ? PA_OK : (int)e;
}

struct XchgParams {
    void* const* peers;
    uint32_t* epochs;
    int* status;
    int rank, world;
};

static int decode_entry(int kv, bool overlap, const float* q, float* out, float* part_m,
                        float* part_l, float* part_o, const void* k_pool, const void* v_pool,
                        const float* k_scales, const float* v_scales, const int32_t* table,
                        int num_beams, int H, int num_tiles, int total_pages,
                        const int32_t* beam_ids, const int32_t* ctx_lens, int B, int T, int D,
                        int tile_size, float temperature, const float* rope, float* lse_out,
                        void* ws, size_t ws_bytes, pa_stream_t stream, const XchgParams* xch = nullptr) {
    PA_CHECK_ARG(q && k_pool && v_pool && table);
    PA_CHECK_ARG(part_m ? (part_l && part_o) : (out != nullptr));
    PA_CHECK_ARG(num_beams > 0 && H > 0 && num_tiles > 0 && total_pages > 0 && B >= 0 && T >= 0);
    PA_CHECK_ARG(temperature != 0.f && tile_size > 0);
    PA_CHECK_ARG((uintptr_t)k_pool % 16 == 0 && (uintptr_t)v_pool % 16 == 0);
    if (kv == 1) PA_CHECK_ARG(k_scales && v_scales && (uintptr_t)k_scales % 16 == 0 && (uintptr_t)v_scales % 16 == 0);
    if ((D != 64 && D != 128) || tile_size % kUnitTok != 0) return PA_ERR_UNSUPPORTED;
    if (B == 0) return PA_OK;
    DecodeArgs a{};
    a.q = q; a.out = out; a.lse_out = lse_out;
    a.part_m = part_m; a.part_l = part_l; a.part_o = part_o;
    a.k_pool = static_cast<const uint8_t*>(k_pool);
    a.v_pool = static_cast<const uint8_t*>(v_pool);
    a.k_scales = k_scales; a.v_scales = v_scales;
    a.table = table; a.beam_ids = beam_ids; a.ctx_lens = ctx_lens; a.rope = rope;
    a.num_beams = num_beams; a.H = H; a.num_tiles = num_tiles; a.total_pages = total_pages;
    a.B = B; a.T = T; a.tile_size = tile_size;
    a.num_splits = 1;
    a.evict_first = beam_ids ? 0 : 1;  // shared-prefix pages are re-read by sibling beams: keep them in L2
    a.qscale = kLog2e / temperature;
    if (xch) {
        PA_CHECK_ARG(xch->peers && xch->epochs && xch->world > 0 && xch->world <= 32 &&
                     xch->rank >= 0 && xch->rank < xch->world && !part_m);
        a.xch_peers = reinterpret_cast<uint8_t* const*>(xch->peers);
        a.xch_epochs = xch->epochs;
        a.xch_status = xch->status;
        a.xch_rank = xch->rank;
        a.xch_world = xch->world;
    }
    cudaStream_t st = as_stream(stream);
    if (kv == 0) {
        return D == 128 ? launch_decode<128, 0>(a, overlap, ws, ws_bytes, st)
                        : launch_decode<64, 0>(a, overlap, ws, ws_bytes, st);
    }
    if (kv == 2) {
        return D == 128 ? launch_decode<128, 2>(a, overlap, ws, ws_bytes, st)
                        : launch_decode<64, 2>(a, overlap, ws, ws_bytes, st);
    }
    return D == 128 ? launch_decode<128, 1>(a, overlap, ws, ws_bytes, st)
                    : launch_decode<64, 1>(a, overlap, ws, ws_bytes, st);
}

}  // namespace pa

using namespace pa;

PA_API size_t pa_decode_workspace_bytes(int B, int num_heads, int head_dim, int num_tiles,
                                        int tile_size) {
    if (B < 0 || num_heads <= 0 || head_dim <= 0 || num_tiles <= 0 || tile_size <= 0) return 0;
    const DeviceInfo& di = device_info();
    return ws_bytes_needed(B, num_heads, head_dim, num_tiles, tile_size, di.ok ? di.sm_count : 148);
}

#define PA_DECODE_COMMON_PARAMS                                                                  \
    const int32_t *d_table, int num_beams, int num_heads, int num_tiles, int total_pages,        \
        const int32_t *d_beam_ids, const int32_t *d_ctx_lens, int B, int T, int head_dim,        \
        int tile_size, float temperature, const float *d_rope
#define PA_DECODE_COMMON_ARGS                                                                    \
    d_table, num_beams, num_heads, num_tiles, total_pages, d_beam_ids, d_ctx_lens, B, T,         \
        head_dim, tile_size, temperature, d_rope

PA_API int pa_paged_decode_f16(const float* d_q, float* d_out, const void* d_k_pool,
                               const void* d_v_pool, PA_DECODE_COMMON_PARAMS, float* d_lse_out,
                               void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    return decode_entry(0, false, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool, nullptr,
                        nullptr, PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace, workspace_bytes, stream);
}
PA_API int pa_paged_decode_f16_overlap(const float* d_q, float* d_out, const void* d_k_pool,
                                       const void* d_v_pool, PA_DECODE_COMMON_PARAMS,
                                       float* d_lse_out, void* d_workspace, size_t workspace_bytes,
                                       pa_stream_t stream) {
    return decode_entry(0, true, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool, nullptr,
                        nullptr, PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace, workspace_bytes, stream);
}
PA_API int pa_paged_decode_i8(const float* d_q, float* d_out, const int8_t* d_k_pool,
                              const int8_t* d_v_pool, const float* d_k_scales,
                              const float* d_v_scales, PA_DECODE_COMMON_PARAMS, float* d_lse_out,
                              void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    return decode_entry(1, false, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool,
                        d_k_scales, d_v_scales, PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace,
                        workspace_bytes, stream);
}
PA_API int pa_paged_decode_i8_overlap(const float* d_q, float* d_out, const int8_t* d_k_pool,
                                      const int8_t* d_v_pool, const float* d_k_scales,
                                      const float* d_v_scales, PA_DECODE_COMMON_PARAMS,
                                      float* d_lse_out, void* d_workspace, size_t workspace_bytes,
                                      pa_stream_t stream) {
    return decode_entry(1, true, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool,
                        d_k_scales, d_v_scales, PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace,
                        workspace_bytes, stream);
}
// fp32 pools: KVTileCache<float>, the reference's default instantiation (attention_config.hpp:15,
// kv_tile_cache.cpp:127).  Same kernels, 4-byte elements.
PA_API int pa_paged_decode_f32(const float* d_q, float* d_out, const float* d_k_pool, const float* d_v_pool,
                               PA_DECODE_COMMON_PARAMS, float* d_lse_out, void* d_workspace, size_t workspace_bytes,
                               pa_stream_t stream) {
    return decode_entry(2, false, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool, nullptr, nullptr,
                        PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace, workspace_bytes, stream);
}
PA_API int pa_paged_decode_f32_overlap(const float* d_q, float* d_out, const float* d_k_pool, const float* d_v_pool,
                                       PA_DECODE_COMMON_PARAMS, float* d_lse_out, void* d_workspace,
                                       size_t workspace_bytes, pa_stream_t stream) {
    return decode_entry(2, true, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool, nullptr, nullptr,
                        PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace, workspace_bytes, stream);
}

// Kernel choice for the partial entries, by problem size: below ~0.5 GB of K/V (one GPU's share of a 128K-token
// sequence is 268 MB) the split-KV grid kernel is faster than the persistent streaming kernel (measured on that share:
// 50 us incl. the merge launch vs 57 us -- the streaming kernel's static split waits for its slowest SM); above it the
// streaming kernel wins.  PA_PARTIAL_DIRECT=0 / 1 overrides.
static bool partial_uses_streaming(int B, int num_heads, int T) {
    const char* env = getenv("PA_PARTIAL_DIRECT");
    if (env) return atoi(env) != 1;
    return !grid_kernel_range((int64_t)B * num_heads, (T + kUnitTok - 1) / kUnitTok);
}

PA_API int pa_paged_decode_f16_partial(const float* d_q, float* d_part_m, float* d_part_l,
                                       float* d_part_o, const void* d_k_pool, const void* d_v_pool,
                                       PA_DECODE_COMMON_PARAMS, void* d_workspace,
                                       size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_part_m && d_part_l && d_part_o);
    return decode_entry(0, partial_uses_streaming(B, num_heads, T), d_q, nullptr, d_part_m, d_part_l, d_part_o, d_k_pool, d_v_pool,
                        nullptr, nullptr, PA_DECODE_COMMON_ARGS, nullptr, d_workspace,
                        workspace_bytes, stream);
}
PA_API int pa_paged_decode_i8_partial(const float* d_q, float* d_part_m, float* d_part_l, float* d_part_o,
                                      const int8_t* d_k_pool, const int8_t* d_v_pool, const float* d_k_scales,
                                      const float* d_v_scales, PA_DECODE_COMMON_PARAMS, void* d_workspace,
                                      size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_part_m && d_part_l && d_part_o);
    return decode_entry(1, partial_uses_streaming(B, num_heads, T), d_q, nullptr, d_part_m, d_part_l, d_part_o, d_k_pool, d_v_pool,
                        d_k_scales, d_v_scales, PA_DECODE_COMMON_ARGS, nullptr, d_workspace, workspace_bytes, stream);
}

PA_API int pa_paged_decode_f16_splitkv(const float* d_q, float* d_out, const void* d_k_pool,
                                       const void* d_v_pool, PA_DECODE_COMMON_PARAMS, float* d_lse_out,
                                       void* d_workspace, size_t workspace_bytes, void* const* d_peer_bufs,
                                       int rank, int world, uint32_t* d_epochs, int* d_status,
                                       pa_stream_t stream) {
    XchgParams x{d_peer_bufs, d_epochs, d_status, rank, world};
    return decode_entry(0, true, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool, nullptr,
                        nullptr, PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace, workspace_bytes, stream, &x);
}
PA_API int pa_paged_decode_i8_splitkv(const float* d_q, float* d_out, const int8_t* d_k_pool,
                                      const int8_t* d_v_pool, const float* d_k_scales, const float* d_v_scales,
                                      PA_DECODE_COMMON_PARAMS, float* d_lse_out, void* d_workspace,
                                      size_t workspace_bytes, void* const* d_peer_bufs, int rank, int world,
                                      uint32_t* d_epochs, int* d_status, pa_stream_t stream) {
    XchgParams x{d_peer_bufs, d_epochs, d_status, rank, world};
    return decode_entry(1, true, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool, d_k_scales,
                        d_v_scales, PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace, workspace_bytes, stream, &x);
}

PA_API int pa_paged_decode_f16_group(const float* d_q, float* d_out, const void* d_k_pool,
                                     const void* d_v_pool, PA_DECODE_COMMON_PARAMS, int beam_width,
                                     float* d_lse_out, void* d_workspace, size_t workspace_bytes,
                                     pa_stream_t stream) {
    PA_CHECK_ARG(d_q && d_out && d_k_pool && d_v_pool && d_table);
    PA_CHECK_ARG(num_beams > 0 && num_heads > 0 && num_tiles > 0 && total_pages > 0 && B >= 0 && T >= 0);
    PA_CHECK_ARG(temperature != 0.f && tile_size > 0 && beam_width >= 1);
    PA_CHECK_ARG((uintptr_t)d_k_pool % 128 == 0 && (uintptr_t)d_v_pool % 128 == 0);
    if (head_dim != 128 || tile_size % kUnitTok != 0 || beam_width > kGroupMaxW) return PA_ERR_UNSUPPORTED;
    if (B % beam_width != 0) return PA_ERR_INVALID_ARG;
    if (B == 0) return PA_OK;
    DecodeArgs a{};
    a.q = d_q; a.out = d_out; a.lse_out = d_lse_out;
    a.k_pool = static_cast<const uint8_t*>(d_k_pool);
    a.v_pool = static_cast<const uint8_t*>(d_v_pool);
    a.table = d_table; a.beam_ids = d_beam_ids; a.ctx_lens = d_ctx_lens; a.rope = d_rope;
    a.num_beams = num_beams; a.H = num_heads; a.num_tiles = num_tiles; a.total_pages = total_pages;
    a.B = B; a.T = T; a.tile_size = tile_size;
    a.num_splits = 1;
    a.qscale = kLog2e / temperature;
    return launch_group(a, beam_width, d_workspace, workspace_bytes, as_stream(stream));
}

// int8 pages for beam groups.  The tensor-core group kernel stages fp16 boxes for ldmatrix; int8 pages take the
// per-row streaming kernel through the beam_ids indirection instead: same results, and the shared-prefix pages of a
// group are served from L2 after their first read (L2::evict_first is off whenever beam_ids are given; sibling beams
// are adjacent rows, i.e. adjacent chunk ids, so they stream the same pages at about the same time) -- HBM sees the
// unique bytes roughly once, the SMs ingest the logical bytes.  beam_width only validates the grouping.
PA_API int pa_paged_decode_i8_group(const float* d_q, float* d_out, const int8_t* d_k_pool, const int8_t* d_v_pool,
                                    const float* d_k_scales, const float* d_v_scales, PA_DECODE_COMMON_PARAMS,
                                    int beam_width, float* d_lse_out, void* d_workspace, size_t workspace_bytes,
                                    pa_stream_t stream) {
    PA_CHECK_ARG(beam_width >= 1 && B >= 0);
    if (B % beam_width != 0) return PA_ERR_INVALID_ARG;
    return decode_entry(1, true, d_q, d_out, nullptr, nullptr, nullptr, d_k_pool, d_v_pool, d_k_scales, d_v_scales,
                        PA_DECODE_COMMON_ARGS, d_lse_out, d_workspace, workspace_bytes, stream);
}

PA_API int pa_lse_combine(const float* d_part_m, const float* d_part_l, const float* d_part_o,
                          int n_parts, int rows, int head_dim, float* d_out, float* d_lse_out,
                          pa_stream_t stream) {
    PA_CHECK_ARG(d_part_m && d_part_l && d_part_o && d_out && n_parts > 0 && rows >= 0 && head_dim > 0);
    if (rows == 0) return PA_OK;
    lse_combine_kernel<<<rows, 128, 0, as_stream(stream)>>>(d_part_m, d_part_l, d_part_o, n_parts,
                                                             rows, head_dim, d_out, d_lse_out);
    PA_RETURN_LAUNCH_STATUS();
}
