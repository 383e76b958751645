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
#include <cstdlib>
#include <cstring>

#include "mma_utils.cuh"
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

// ------------------------------------------------------------------ tensor-core prefill kernel
// Flash-attention forward over the paged cache for fp16 pages, head_dim 128.  One CTA = 64 consecutive
// query positions of one (row, head): 4 compute warps x 16 queries (mma.sync m16n8k16, fp16 operands,
// fp32 accumulate: S = Q K^T and O += P V per 16-token unit, online softmax in registers, the S
// accumulator fragments are re-packed in place as the A operand of P V) + 1 producer warp that walks the
// page table and stages K/V units with TMA tensor copies (SWIZZLE_128B, conflict-free ldmatrix) into a
// 4-stage ring shared by the 4 warps.  Causal limit per query; tiles are launched last-first (the last
// query tile of a row has the longest context).  q/out stay in the reference's [B, H, T, D] layout.
constexpr int kPfQ = 64;
constexpr int kPfStages = 4;
constexpr int kPfStageBytes = 8192;

struct PrefillArgs {
    const float* q;
    float* out;
    const int32_t* table;
    const int32_t* beam_ids;
    const int32_t* ctx_start;
    int num_beams, H, num_tiles, total_pages, B, Tq, tile_size;
    float qscale;  // log2(e) / temperature
    // int8 pages (KV == 1): raw pools + per-(page,row) scales, staged by plain bulk copies
    const int8_t* k_pool;
    const int8_t* v_pool;
    const float* k_scales;
    const float* v_scales;
};

constexpr int kPfRawStages = 4;
constexpr int kPfRawBytes = 2048 + 2048 + 64 + 64;  // int8 K unit, V unit, 16 + 16 f32 scales
constexpr int kPfF16StageI8 = kPfStageBytes + 128;  // fp16 K/V images + 16 + 16 reciprocal scales

// KV = 0: fp16 pages (TMA tensor copies straight into the swizzled fp16 stage).  KV = 1: int8 pages: the
// producer bulk-copies the raw unit (K, V, scales) into a second ring and a CONVERTER warp (warp 5) rewrites
// it as swizzled fp16 (exact: byte -> 1024 + u by PRMT, minus 1152 by HSUB2) plus reciprocal scales; the
// compute warps then apply the per-token K scale to the score columns and fold the V scale into P.
template <int KV>
__global__ void __launch_bounds__(192) prefill_fa_kernel(const __grid_constant__ CUtensorMap tmK,
                                                         const __grid_constant__ CUtensorMap tmV,
                                                         const PrefillArgs a) {
    constexpr int D = 128, S = kPfStages, R = kPfRawStages;
    constexpr int FST = KV == 0 ? kPfStageBytes : kPfF16StageI8;  // bytes per fp16 stage
    extern __shared__ __align__(1024) uint8_t smem_p[];
    const uint32_t base = (smem_u32(smem_p) + 1023u) & ~1023u;
    uint8_t* gen = smem_p + (base - smem_u32(smem_p));
    // layout: [S fp16 stages of 8 KB][KV==1: S x 128 B scale areas][KV==1: R raw stages][barriers][meta]
    const uint32_t scale0 = base + S * kPfStageBytes;
    const uint32_t raw0 = scale0 + (KV == 1 ? S * 128 : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(gen + (raw0 - base) + (KV == 1 ? R * kPfRawBytes : 0));
    int* meta = reinterpret_cast<int*>(bars + 2 * S + 2 * R);  // meta[S] (fp16 stages), meta[S..S+R) (raw stages)
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
    const uint32_t rfull0 = smem_u32(bars + 2 * S), rempty0 = smem_u32(bars + 2 * S + R);
    (void)FST;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(full0 + i * 8, 1);
            mbar_init(empty0 + i * 8, 4);  // one arrival per compute warp
        }
        if (KV == 1) {
            for (int i = 0; i < R; ++i) {
                mbar_init(rfull0 + i * 8, 1);
                mbar_init(rempty0 + i * 8, 1);  // the converter warp
            }
        }
        mbar_fence_init();
        if (KV == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV));
        }
    }
    __syncthreads();

    const int nqt = (a.Tq + kPfQ - 1) / kPfQ;
    const int bh_total = a.B * a.H;
    // CTA order: tile-major, longest tiles first (measured better than keeping the query tiles of one
    // (row, head) adjacent for L2 reuse: 0.187 vs 0.231 ms at B = 1, Tq = 2048)
    const int qt = nqt - 1 - (int)(blockIdx.x / bh_total);
    const int bh = (int)(blockIdx.x % bh_total);
    const int b = bh / a.H, h = bh - b * a.H;
    const int start = a.ctx_start ? a.ctx_start[b] : 0;
    const int q_last = min(a.Tq, (qt + 1) * kPfQ) - 1;
    const int cap = a.num_tiles * a.tile_size;
    const int kmax = min(cap, start + q_last + 1);  // keys [0, kmax) can be visible to this tile
    const int n_units = (kmax + 15) >> 4;
    const int upt = a.tile_size >> 4;

    if (warp == 4) {  // ---------------- producer
        if (elect_one()) {
            const int beam = a.beam_ids ? a.beam_ids[b] : b;
            const int32_t* trow = ((unsigned)beam < (unsigned)a.num_beams)
                                      ? a.table + ((int64_t)beam * a.H + h) * a.num_tiles : nullptr;
            int s = 0;
            uint32_t ph = 1;
            for (int u = 0; u < n_units; ++u) {
                int page = trow ? __ldg(trow + u / upt) : -1;
                if ((unsigned)page >= (unsigned)a.total_pages) page = -1;
                const int nvalid = page >= 0 ? min(16, kmax - u * 16) : 0;
                const int64_t row0 = (int64_t)page * a.tile_size + (u % upt) * 16;  // token row in the pool
                if (KV == 0) {
                    mbar_wait(empty0 + s * 8, ph);
                    meta[s] = nvalid;
                    if (page >= 0) {
                        const uint32_t dst = base + s * kPfStageBytes;
                        fence_proxy_async();
                        mbar_arrive_expect_tx(full0 + s * 8, kPfStageBytes);
                        tma_load_2d(dst, &tmK, 0, (int)row0, full0 + s * 8);
                        tma_load_2d(dst + 2048, &tmK, 64, (int)row0, full0 + s * 8);
                        tma_load_2d(dst + 4096, &tmV, 0, (int)row0, full0 + s * 8);
                        tma_load_2d(dst + 6144, &tmV, 64, (int)row0, full0 + s * 8);
                    } else {
                        mbar_arrive(full0 + s * 8);  // unmapped page: skipped (...fused.cu:32)
                    }
                    if (++s == S) {
                        s = 0;
                        ph ^= 1u;
                    }
                } else {
                    mbar_wait(rempty0 + s * 8, ph);
                    meta[S + s] = nvalid;
                    if (page >= 0) {
                        const uint32_t dst = raw0 + s * kPfRawBytes;
                        fence_proxy_async();
                        mbar_arrive_expect_tx(rfull0 + s * 8, kPfRawBytes);
                        bulk_g2s_nohint(dst, a.k_pool + row0 * D, 2048, rfull0 + s * 8);
                        bulk_g2s_nohint(dst + 2048, a.v_pool + row0 * D, 2048, rfull0 + s * 8);
                        bulk_g2s_nohint(dst + 4096, a.k_scales + row0, 64, rfull0 + s * 8);
                        bulk_g2s_nohint(dst + 4160, a.v_scales + row0, 64, rfull0 + s * 8);
                    } else {
                        mbar_arrive(rfull0 + s * 8);
                    }
                    if (++s == R) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
        return;
    }
    if (KV == 1 && warp == 5) {  // ---------------- converter: raw int8 unit -> swizzled fp16 stage
        int rs = 0, fs = 0;
        uint32_t rph = 0, fph = 1;
        const int t = lane >> 1, hf = lane & 1;  // token row, 64-dim half (= swizzle box)
        const __half2 off = __floats2half2_rn(1152.f, 1152.f);
        for (int u = 0; u < n_units; ++u) {
            mbar_wait(rfull0 + rs * 8, rph);
            mbar_wait(empty0 + fs * 8, fph);
            const int nvalid = meta[S + rs];
            if (nvalid > 0) {
                const uint32_t src = raw0 + rs * kPfRawBytes, dst = base + fs * kPfStageBytes;
#pragma unroll
                for (int kvsel = 0; kvsel < 2; ++kvsel) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {  // 16 int8 -> two 16-byte chunks of 8 halfs
                        const uint4 w = lds_128(src + kvsel * 2048 + t * 128 + hf * 64 + c4 * 16);
                        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
                        uint32_t h[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t x = ww[i] ^ 0x80808080u;  // u = b + 128
                            uint32_t p01 = __byte_perm(x, 0x64646464u, 0x4140);  // halfs (1024 + u0, 1024 + u1)
                            uint32_t p23 = __byte_perm(x, 0x64646464u, 0x4342);
                            __half2 a01 = __hsub2(*reinterpret_cast<__half2*>(&p01), off);
                            __half2 a23 = __hsub2(*reinterpret_cast<__half2*>(&p23), off);
                            h[2 * i] = *reinterpret_cast<uint32_t*>(&a01);
                            h[2 * i + 1] = *reinterpret_cast<uint32_t*>(&a23);
                        }
                        const uint32_t boxrow = dst + kvsel * 4096 + hf * 2048 + t * 128;
                        const int c = 2 * c4;
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(boxrow + (((c) ^ (t & 7)) << 4)),
                                     "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(boxrow + (((c + 1) ^ (t & 7)) << 4)),
                                     "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
                    }
                }
                // reciprocal scales: lanes 0-15 K rows, 16-31 V rows (int8_quant.cpp:46-57: x = q / scale)
                float sc;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sc) : "r"(src + 4096 + lane * 4));
                sc = fast_rcp(sc);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(scale0 + fs * 128 + lane * 4), "f"(sc) : "memory");
            }
            __syncwarp();
            if (lane == 0) {
                meta[fs] = nvalid;
                mbar_arrive(full0 + fs * 8);
                mbar_arrive(rempty0 + rs * 8);
            }
            if (++rs == R) {
                rs = 0;
                rph ^= 1u;
            }
            if (++fs == S) {
                fs = 0;
                fph ^= 1u;
            }
        }
        return;
    }

    // ---------------- compute warps
    const int g8 = lane >> 2, j4 = lane & 3;
    const int t0 = qt * kPfQ + warp * 16 + g8, t1 = t0 + 8;  // this thread's two query rows
    const int64_t qbase = ((int64_t)b * a.H + h) * a.Tq;
    uint32_t qa[8][4];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int d0 = s * 16 + hh * 8 + j4 * 2;
            float2 x0 = make_float2(0.f, 0.f), x1 = make_float2(0.f, 0.f);
            if (t0 < a.Tq) x0 = *reinterpret_cast<const float2*>(a.q + (qbase + t0) * D + d0);
            if (t1 < a.Tq) x1 = *reinterpret_cast<const float2*>(a.q + (qbase + t1) * D + d0);
            qa[s][hh * 2 + 0] = pack_half2(x0.x * a.qscale, x0.y * a.qscale);
            qa[s][hh * 2 + 1] = pack_half2(x1.x * a.qscale, x1.y * a.qscale);
        }
    }
    const int qpos0 = (t0 < a.Tq) ? start + t0 : -1, qpos1 = (t1 < a.Tq) ? start + t1 : -1;  // last visible key
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[16][4];
#pragma unroll
    for (int t = 0; t < 16; ++t) { o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f; }

    int s = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int u = 0; u < n_units; ++u) {
        mbar_wait(full0 + s * 8, ph);
        const int nvalid = meta[s];
        // units entirely above this warp's last query are masked for every row: skip the math
        const int kp0 = u * 16;
        if (nvalid > 0 && kp0 <= start + qt * kPfQ + warp * 16 + 15) {  // (warp-uniform condition)
            const uint32_t sb = base + s * kPfStageBytes;
            if (nvalid < 16) {  // rows past the context end may hold anything: zero those V rows
                for (int i = lane; i < (16 - nvalid) * 16; i += 32) {
                    const int r = nvalid + i / 16, c = i % 16;
                    const uint32_t addr = sb + 4096 + (c >> 3) * 2048 + r * 128 + (((c & 7) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
                }
                __syncwarp();
            }
            float sacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            {
                const int mi = lane >> 3;
                const int r = (lane & 7) + (mi >> 1) * 8;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int c = 2 * (ks & 3) + (mi & 1);
                    const uint32_t addr = sb + (ks >> 2) * 2048 + r * 128 + ((c ^ (r & 7)) << 4);
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(addr, b0, b1, b2, b3);
                    mma_16816_full(sacc[0], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
                    mma_16816_full(sacc[1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
                }
            }
            float vsc[2][2] = {{1.f, 1.f}, {1.f, 1.f}};
            if (KV == 1) {  // per-token dequantisation scales of this thread's 4 score columns
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float2 kk, vv;
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(kk.x), "=f"(kk.y)
                                 : "r"(scale0 + s * 128 + (nt * 8 + j4 * 2) * 4));
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(vv.x), "=f"(vv.y)
                                 : "r"(scale0 + s * 128 + 64 + (nt * 8 + j4 * 2) * 4));
                    sacc[nt][0] *= kk.x; sacc[nt][1] *= kk.y;
                    sacc[nt][2] *= kk.x; sacc[nt][3] *= kk.y;
                    vsc[nt][0] = vv.x; vsc[nt][1] = vv.y;
                }
            }
            // causal mask + online softmax; sacc[nt][0,1] -> row t0, sacc[nt][2,3] -> row t1
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kp = kp0 + nt * 8 + j4 * 2 + e;
                    const bool in = (nt * 8 + j4 * 2 + e) < nvalid;
                    sacc[nt][e] = (in && kp <= qpos0) ? sacc[nt][e] : -INFINITY;
                    sacc[nt][2 + e] = (in && kp <= qpos1) ? sacc[nt][2 + e] : -INFINITY;
                    mx0 = fmaxf(mx0, sacc[nt][e]);
                    mx1 = fmaxf(mx1, sacc[nt][2 + e]);
                }
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
            const float c0 = (mn0 == -INFINITY) ? 1.f : fast_exp2(m_run[0] - mn0);
            const float c1 = (mn1 == -INFINITY) ? 1.f : fast_exp2(m_run[1] - mn1);
            float p[2][4];
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    p[nt][e] = (sacc[nt][e] == -INFINITY) ? 0.f : fast_exp2(sacc[nt][e] - mn0);
                    p[nt][2 + e] = (sacc[nt][2 + e] == -INFINITY) ? 0.f : fast_exp2(sacc[nt][2 + e] - mn1);
                    ps0 += p[nt][e];
                    ps1 += p[nt][2 + e];
                }
            }
            l_run[0] = fmaf(l_run[0], c0, ps0);
            l_run[1] = fmaf(l_run[1], c1, ps1);
            m_run[0] = mn0;
            m_run[1] = mn1;
            if (__any_sync(0xffffffffu, c0 != 1.f || c1 != 1.f)) {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    o[t][0] *= c0; o[t][1] *= c0;
                    o[t][2] *= c1; o[t][3] *= c1;
                }
            }
            if (KV == 1) {  // V dequantisation folded into P (the denominator above used the unscaled weights)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {  // (masked tokens keep an exact 0 whatever their scale bytes hold)
                    p[nt][0] = p[nt][0] == 0.f ? 0.f : p[nt][0] * vsc[nt][0];
                    p[nt][1] = p[nt][1] == 0.f ? 0.f : p[nt][1] * vsc[nt][1];
                    p[nt][2] = p[nt][2] == 0.f ? 0.f : p[nt][2] * vsc[nt][0];
                    p[nt][3] = p[nt][3] == 0.f ? 0.f : p[nt][3] * vsc[nt][1];
                }
            }
            const uint32_t pa0 = pack_half2(p[0][0], p[0][1]), pa1 = pack_half2(p[0][2], p[0][3]);
            const uint32_t pa2 = pack_half2(p[1][0], p[1][1]), pa3 = pack_half2(p[1][2], p[1][3]);
            {
                const int mi = lane >> 3;
                const int r = (lane & 7) + (mi & 1) * 8;
#pragma unroll
                for (int n2 = 0; n2 < 8; ++n2) {
                    const int c = 2 * (n2 & 3) + (mi >> 1);
                    const uint32_t addr = sb + 4096 + (n2 >> 2) * 2048 + r * 128 + ((c ^ (r & 7)) << 4);
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(addr, b0, b1, b2, b3);
                    mma_16816_full(o[2 * n2], pa0, pa1, pa2, pa3, b0, b1);
                    mma_16816_full(o[2 * n2 + 1], pa0, pa1, pa2, pa3, b2, b3);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + s * 8);
        if (++s == S) {
            s = 0;
            ph ^= 1u;
        }
    }
    // normalise and store: o[nt][0,1] -> (row t0, dims nt*8 + j4*2 + {0,1}); o[nt][2,3] -> row t1
    float l0 = l_run[0], l1 = l_run[1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / (l0 + 1e-6f), i1 = 1.f / (l1 + 1e-6f);  // softmax_lut.cpp:224 epsilon (App. A D4)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
        const int d = nt * 8 + j4 * 2;
        if (t0 < a.Tq) *reinterpret_cast<float2*>(a.out + (qbase + t0) * D + d) = make_float2(o[nt][0] * i0, o[nt][1] * i0);
        if (t1 < a.Tq) *reinterpret_cast<float2*>(a.out + (qbase + t1) * D + d) = make_float2(o[nt][2] * i1, o[nt][3] * i1);
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace pa

using namespace pa;

// prefill_tc.cu: tcgen05 flash-attention prefill (fp16 pages, head_dim 128)
int pa_prefill_tc_launch(int kv, int head_dim, const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                         const float* d_k_scales, const float* d_v_scales, const int32_t* d_table, int num_beams,
                         int num_heads, int num_tiles, int total_pages, const int32_t* d_beam_ids,
                         const int32_t* d_ctx_start, int B, int Tq, int tile_size, float temperature, int token_major,
                         cudaStream_t st);

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
                         void* d_workspace, size_t workspace_bytes, pa_stream_t stream, int token_major = 0) {
    PA_CHECK_ARG(d_q && d_out && d_k_pool && d_v_pool && d_table);
    PA_CHECK_ARG(B >= 0 && Tq > 0 && num_heads > 0 && head_dim > 0 && head_dim % 4 == 0);
    PA_CHECK_ARG(num_beams > 0 && num_tiles > 0 && total_pages > 0 && tile_size > 0 && temperature != 0.f);
    if (B == 0) return PA_OK;
    if ((head_dim == 128 || head_dim == 64) && tile_size % 16 == 0 && (uintptr_t)d_k_pool % 128 == 0 && (uintptr_t)d_v_pool % 128 == 0 &&
        d_q != d_out && !(getenv("PA_PREFILL_FA") && atoi(getenv("PA_PREFILL_FA")) == 0) &&
        (kv == 0 || ((uintptr_t)d_k_scales % 16 == 0 && (uintptr_t)d_v_scales % 16 == 0))) {
        // tcgen05 kernel (prefill_tc.cu, fp16 and int8 pages) unless PA_PREFILL_TC=0 asks for the mma.sync kernel below
        if (!(getenv("PA_PREFILL_TC") && atoi(getenv("PA_PREFILL_TC")) == 0)) {
            const int stc = pa_prefill_tc_launch(kv, head_dim, d_q, d_out, d_k_pool, d_v_pool, d_k_scales, d_v_scales, d_table,
                                                 num_beams, num_heads, num_tiles, total_pages, d_beam_ids, d_ctx_start, B,
                                                 Tq, tile_size, temperature, token_major, as_stream(stream));
            if (stc != PA_ERR_UNSUPPORTED) return stc;
        }
        if (token_major) return PA_ERR_UNSUPPORTED;  // only the tcgen05 kernel reads strided rows
        // mma.sync flash-attention kernel (head_dim 128), straight on the [B, H, Tq, D] layout (no workspace)
        CUtensorMap tmK, tmV;
        const uint64_t total_tokens = (uint64_t)total_pages * tile_size;
        bool maps_ok = head_dim == 128;
        if (kv == 0 && maps_ok) maps_ok = make_pool_map(&tmK, d_k_pool, total_tokens) && make_pool_map(&tmV, d_v_pool, total_tokens);
        else memset(&tmK, 0, sizeof(tmK)), memset(&tmV, 0, sizeof(tmV));  // unused by the int8 variant
        if (maps_ok) {
            PrefillArgs pa{d_q, d_out, d_table, d_beam_ids, d_ctx_start, num_beams, num_heads, num_tiles,
                           total_pages, B, Tq, tile_size, 1.4426950408889634f / temperature,
                           static_cast<const int8_t*>(d_k_pool), static_cast<const int8_t*>(d_v_pool), d_k_scales,
                           d_v_scales};
            const int nqt = (Tq + kPfQ - 1) / kPfQ;
            const int64_t ctas = (int64_t)B * num_heads * nqt;
            PA_CHECK_ARG(ctas <= 0x7fffffff);
            const size_t smem = (size_t)kPfStages * kPfStageBytes + (kv == 1 ? kPfStages * 128 + kPfRawStages * kPfRawBytes : 0) +
                                (2 * kPfStages + 2 * kPfRawStages) * 8 + (kPfStages + kPfRawStages) * 4 + 1024;
            static bool attr_done[64][2] = {};
            int dev = 0;
            cudaGetDevice(&dev);
            auto kern = kv == 0 ? prefill_fa_kernel<0> : prefill_fa_kernel<1>;
            if (!attr_done[dev & 63][kv]) {
                cudaError_t e0 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e0 != cudaSuccess) return (int)e0;
                attr_done[dev & 63][kv] = true;
            }
            kern<<<(unsigned)ctas, kv == 0 ? 160 : 192, smem, as_stream(stream)>>>(tmK, tmV, pa);
            PA_RETURN_LAUNCH_STATUS();
        }
    }
    if (token_major) return PA_ERR_UNSUPPORTED;
    PA_CHECK_ARG(d_workspace);
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

// Token-major variants: q/out are [B, Tq, H, D] (the decoders' activation layout after the QKV projection, so no
// permute + copy either side of the attention).  Served by the tcgen05 kernel only; PA_ERR_UNSUPPORTED where
// that kernel does not apply (head_dim other than 64/128, pages that are not 16 << k tokens, misaligned pools) --
// the caller then permutes and takes pa_paged_prefill_f16/_i8.
PA_API int pa_paged_prefill_f16_tokmajor(const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                                         const int32_t* d_table, int num_beams, int num_heads, int num_tiles,
                                         int total_pages, const int32_t* d_beam_ids, const int32_t* d_ctx_start, int B,
                                         int Tq, int head_dim, int tile_size, float temperature, pa_stream_t stream) {
    return prefill_entry(0, d_q, d_out, d_k_pool, d_v_pool, nullptr, nullptr, d_table, num_beams, num_heads, num_tiles,
                         total_pages, d_beam_ids, d_ctx_start, B, Tq, head_dim, tile_size, temperature, nullptr, 0,
                         stream, 1);
}

PA_API int pa_paged_prefill_i8_tokmajor(const float* d_q, float* d_out, const int8_t* d_k_pool, const int8_t* d_v_pool,
                                        const float* d_k_scales, const float* d_v_scales, const int32_t* d_table,
                                        int num_beams, int num_heads, int num_tiles, int total_pages,
                                        const int32_t* d_beam_ids, const int32_t* d_ctx_start, int B, int Tq,
                                        int head_dim, int tile_size, float temperature, pa_stream_t stream) {
    PA_CHECK_ARG(d_k_scales && d_v_scales);
    return prefill_entry(1, d_q, d_out, d_k_pool, d_v_pool, d_k_scales, d_v_scales, d_table, num_beams, num_heads,
                         num_tiles, total_pages, d_beam_ids, d_ctx_start, B, Tq, head_dim, tile_size, temperature,
                         nullptr, 0, stream, 1);
}
