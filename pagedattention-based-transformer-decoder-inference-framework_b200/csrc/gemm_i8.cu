// gemm_i8.cu -- s8 x s8 -> s32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::i8) with
// the fused oneDNN-style epilogue of attention_cpu/dnnl_matmul_int8.cpp:7-76.
//
//   acc[b,m,n] = sum_k A[b,m,k] * B[b,k,n]                (exact int32)
//   C_s8       = sat_s8(rne(act(alpha*acc + bias[n])))     alpha = scaleA*scaleB/scaleC
//
// Shapes of interest (decoder MLP, decoder/mlp.hpp:23-41): M = decode batch (<= 256),
// [M x hidden] . [hidden x 4*hidden] and back.  At M = 256 the GEMM sits on the ridge: the
// weight matrix B must stream from HBM exactly once, so a CTA owns a 128-column slab of B over
// (a split of) K and computes ALL rows of M against it (two 128-row accumulators in TMEM).
//
// Pipeline (one CTA per SM, 192 threads):
//   warp 0   : TMA producer -- cp.async.bulk.tensor (UTMALDG) of A tiles [128 x 128 B] (K-major,
//              128B swizzle) and the B tile [128 k-rows x 128 B] (N contiguous = "MN-major",
//              128B swizzle) into a 4-stage shared-memory ring, mbarrier complete_tx.
//   warp 1   : allocates TMEM, one elected lane issues tcgen05.mma.cta_group::1.kind::i8
//              (M=128, N=128, K=32 per instruction), tcgen05.commit releases ring slots and
//              finally signals the epilogue.
//   warps 2-5: epilogue -- tcgen05.ld (32 lanes x 32 columns per warp), scale/bias/activation/
//              round-to-nearest-even/saturate, 128-bit stores of s8 (or raw s32 / split-K
//              partial tiles that a second kernel sums exactly and finishes).
// B is consumed in the reference's own [K, N] row-major layout through an MN-major UMMA
// descriptor: no transpose pass, no repacking of weights.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "pa_common.cuh"

namespace pa {
namespace gemm {

constexpr int BM = 128;      // rows per UMMA / TMEM accumulator
constexpr int BN = 128;      // columns per CTA slab
constexpr int BK = 128;      // bytes of K per pipeline stage (one 128B swizzle row)
constexpr int UK = 32;       // K per tcgen05.mma kind::i8
#ifndef PA_GEMM_STAGES
#define PA_GEMM_STAGES 4
#endif
constexpr int STAGES = PA_GEMM_STAGES;
constexpr int TILE_BYTES = BM * BK;          // 16 KiB: one A tile or one B tile
constexpr int TMEM_COLS = 2 * BN;            // two accumulators
constexpr int NTHREADS = 192;

struct Args {
    int8_t* C8;
    float* Cf;           // f32 output act(alpha*acc + bias) (dequantising epilogue; may be null)
    const float* a_qscale;  // [BATCH*M] per-row: alpha_row = alpha / a_qscale[row] (dynamic activation scales)
    int32_t* C32;        // raw accumulators out (may be null)
    int32_t* acc_ws;     // split-K partial tiles [ksplit][BATCH*M][N], null when ksplit == 1
    const float* bias;
    float alpha;
    int act;
    int M, N, K;
    int stages;          // 1-CTA kernel: ring depth (runtime: more, smaller stages when M is small)
    int a_bytes;         // 1-CTA kernel: bytes of A per stage (m_tiles * a_rows * 128)
    int ksplit;          // K splits
    int kb_per_split;    // BK blocks per split
    int64_t BATCH_rows;  // BATCH * M
    // dynamic row quantisation of the OUTPUT fused into the epilogue (DQ instances)
    unsigned int* dq_rowmax;   // [BATCH*M] bit patterns of max_n |act(alpha*acc + bias)| (non-negative floats order as uints)
    unsigned int* dq_barrier;  // [2] arrive / done counters of the grid barrier
    float* dq_scales;          // [BATCH*M] out: 127 / (rowmax + 1e-6)   (compute_minmax_scale)
    unsigned int dq_expected;  // CTAs in the grid
    // dq_rowmax and dq_barrier are zero ONCE (caller) and left zero by every launch: the last CTA to have read its
    // rows' maxima clears them (no memset node in front of the kernel).
#ifdef PA_GEMM_PROBE
    unsigned long long* probe;  // [0] producer wait cycles, [1] mma wait cycles, [2] total cycles (CTA 0)
#endif
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                               uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (InstrDescriptor): c=S32 (2) [4,6), a=INT8 (1) [7,10), b=INT8 (1) [10,13),
// a K-major (0) [15], b MN-major (1) [16], N>>3 [17,23), M>>4 [24,29).
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) |
                            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int ACT>
__device__ __forceinline__ float epilogue_f32(int acc, float alpha, float bias) {
    float v = __fadd_rn(__fmul_rn(alpha, (float)acc), bias);
    if (ACT == PA_ACT_RELU) v = fmaxf(v, 0.f);
    else if (ACT == PA_ACT_GELU) v = gelu_erf(v);
    return v;
}

template <int ACT>
__device__ __forceinline__ int epilogue_one(int acc, float alpha, float bias) {
    // alpha*acc and +bias round separately, as the reference epilogue restated in the oracle
    // (adding a 0.0f bias is exact, so "no bias" passes 0).
    float v = __fadd_rn(__fmul_rn(alpha, (float)acc), bias);
    if (ACT == PA_ACT_RELU) v = fmaxf(v, 0.f);
    else if (ACT == PA_ACT_GELU) v = gelu_erf(v);
    v = fminf(127.f, fmaxf(-128.f, rintf(v)));  // rne, saturate (NaN -> -128)
    return (int)v;
}

// Same result as epilogue_one for finite inputs, in fewer instructions: cvt.rni.sat.s8.f32 rounds to
// nearest-even and saturates to [-128, 127] in one step.
template <int ACT>
__device__ __forceinline__ uint32_t epilogue_s8(int acc, float alpha, float bias) {
    float v = __fadd_rn(__fmul_rn(alpha, (float)acc), bias);
    if (ACT == PA_ACT_RELU) v = fmaxf(v, 0.f);
    else if (ACT == PA_ACT_GELU) v = gelu_erf(v);
    uint32_t q;
    asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(v));
    return q;
}
// clamp(std::round(y)) to int8 in four instructions (int8_quant.cpp:8-9: round half AWAY from zero, saturate):
// floor(|y| + 0.5) with the sign put back.  The addition is done round-TOWARD-ZERO, which makes it exact where it
// matters: RZ(|y| + 0.5) >= n exactly when |y| + 0.5 >= n for every integer n (n is representable), so truncating it
// gives floor(|y| + 0.5) -- e.g. 0.49999997 + 0.5 stays below 1 (round-to-nearest would give 1.0 and the wrong answer;
// SURVEY App. B has that value).  cvt.rzi.sat truncates and saturates to [-128, 127] in one instruction.
__device__ __forceinline__ uint32_t quantize_half_away_s8(float y) {
    const float half = __uint_as_float((__float_as_uint(y) & 0x80000000u) | 0x3f000000u);  // copysign(0.5, y)
    uint32_t q;
    asm("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(__fadd_rz(y, half)));
    return q;
}
__device__ __forceinline__ uint32_t pack_s8x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return (a & 0xffu) | ((b & 0xffu) << 8) | ((c & 0xffu) << 16) | (d << 24);
}
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

// EPI: 0 = integer outputs only (raw s32 and/or split-K reduction), 1/2/3 = s8 output with
// activation none/relu/gelu (plus optional raw s32).
// CL: thread-block cluster size along N.  The CL CTAs of a cluster work on adjacent B slabs and
// need the same A tiles: each loads 1/CL of the A rows and TMA-multicasts them to all CL CTAs,
// which divides the L2 -> SM traffic for A by CL (the main loop is L2-read bound otherwise).
constexpr int MAXST = 12;                     // 1-CTA kernel: maximum ring depth
constexpr int RING_BYTES = 4 * 3 * TILE_BYTES;  // 192 KB of shared memory for the ring in every configuration

template <int EPI, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args g) {
    extern __shared__ uint8_t smem_raw[];
    asm volatile("griddepcontrol.launch_dependents;");  // a split-K reduction launched behind this grid may become resident
    // 1024-byte alignment for the 128B-swizzled tiles.
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + RING_BYTES;  // full[MAXST], empty[MAXST], tmem_full
    const uint32_t tmem_slot = bar0 + (2 * MAXST + 1) * 8;
    auto full_bar = [&](int s) { return bar0 + s * 8; };
    auto empty_bar = [&](int s) { return bar0 + (MAXST + s) * 8; };
    const uint32_t tmem_full_bar = bar0 + 2 * MAXST * 8;
    // Stage = [A: a_bytes][B: 16 KB].  For M <= 128 only ceil(M/8)*8 rows of A are loaded (a_bytes <
    // 16 KB) and the ring gets deeper: the UMMA still reads 128 rows, the rows past a_bytes are
    // whatever follows in the ring -- they only feed accumulator rows >= M, which are never stored.
    const int nst = g.stages;
    const uint32_t a_bytes = (uint32_t)g.a_bytes, stage_bytes = a_bytes + TILE_BYTES;

#ifdef PA_GEMM_PROBE
    const long long t_entry = clock64();
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    const int m_chunk = blockIdx.y / g.ksplit, split = blockIdx.y % g.ksplit;
    const int m0 = m_chunk * 2 * BM;
    const int batch = blockIdx.z;
    const int m_tiles = (g.M - m0 > BM) ? 2 : 1;
    const int total_kb = (g.K + BK - 1) / BK;
    const int kb0 = split * g.kb_per_split;
    const int kb1 = min(total_kb, kb0 + g.kb_per_split);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CL);  // every CTA of the cluster must have drained slot s
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB));
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();  // peers' barriers are initialised before any remote arrive lands
    tc_fence_after();
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // programmatic dependent launch: everything above ran under the tail of the kernel in front; its results are read below
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef PA_GEMM_PROBE
    long long t_epi0 = 0;
    const long long t_setup = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 5 && g.probe) g.probe[4] = t_setup - t_entry;
#endif

    if (warp == 0) {
        if (elect_one()) {  // elect.sync, not lane == 0: operands stay in uniform registers (pa_common.cuh)
#ifdef PA_GEMM_PROBE
            long long pw = 0, t_start = clock64();
#endif
            int s = 0;
            uint32_t ph = 1;  // empty barriers start "free"
            for (int i = 0; i < nkb; ++i) {
#ifdef PA_GEMM_PROBE
                long long t0 = clock64();
#endif
                mbar_wait(empty_bar(s), ph);
#ifdef PA_GEMM_PROBE
                pw += clock64() - t0;
#endif
                const uint32_t st = base + s * stage_bytes;
                // bytes that will land: the A boxes actually issued for this M chunk + the B tile
                const uint32_t a_load = (m_tiles == 2) ? 2u * TILE_BYTES : (a_bytes < (uint32_t)TILE_BYTES ? a_bytes : (uint32_t)TILE_BYTES);
                mbar_arrive_expect_tx(full_bar(s), a_load + TILE_BYTES);
                const int k0 = (kb0 + i) * BK;
                if (CL == 1) {
                    if (m_tiles == 2) {
                        tma_load_3d(st, &tmA, k0, m0, batch, full_bar(s));
                        tma_load_3d(st + TILE_BYTES, &tmA, k0, m0 + BM, batch, full_bar(s));
                    } else {
                        tma_load_3d(st, &tmA, k0, m0, batch, full_bar(s));  // box = a_bytes / 128 rows
                    }
                } else {
                    // This CTA's share of the A rows of the stage (tmA's box holds a_rows rows),
                    // delivered to the same offset in every CTA of the cluster.
                    const int a_rows = m_tiles * BM / CL;
                    const int r0 = (int)crank * a_rows;
                    tma_load_3d_mc(st + r0 * BK, &tmA, k0, m0 + r0, batch, full_bar(s), kMask);
                }
                tma_load_3d(st + a_bytes, &tmB, n0, k0, batch, full_bar(s));
                if (++s == nst) {
                    s = 0;
                    ph ^= 1u;
                }
            }
#ifdef PA_GEMM_PROBE
            if (blockIdx.x == 5 && g.probe) { g.probe[0] = pw; g.probe[2] = clock64() - t_start; }
#endif
        }
    } else if (warp == 1) {
        if (elect_one()) {
#ifdef PA_GEMM_PROBE
            long long mw = 0, t_start = clock64();
#endif
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
#ifdef PA_GEMM_PROBE
                long long t0 = clock64();
#endif
                mbar_wait(full_bar(s), ph);
#ifdef PA_GEMM_PROBE
                mw += clock64() - t0;
#endif
                tc_fence_after();
                const uint32_t st = base + s * stage_bytes;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    if (mt < m_tiles) {
#pragma unroll
                        for (int ks = 0; ks < BK / UK; ++ks) {
                            // A: K-major SW128, 8-row groups 1024 B apart; +32 B per K step inside the swizzle row.
                            const uint64_t da = make_desc(st + mt * TILE_BYTES + ks * UK, 16, 1024);
                            // B: MN-major SW128, 8 k-rows per 1024 B group; 32 k-rows = 4096 B per K step.
                            const uint64_t db = make_desc(st + a_bytes + ks * UK * BK, TILE_BYTES, 1024);
                            umma_i8(tmem_base + mt * BN, da, db, kIdesc, (i > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                }
                // slot reusable once these MMAs have read it (signalled to every CTA of the cluster)
                if (CL == 1) umma_commit(empty_bar(s));
                else umma_commit_mc(empty_bar(s), kMask);
                if (++s == nst) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            umma_commit(tmem_full_bar);
#ifdef PA_GEMM_PROBE
            if (blockIdx.x == 5 && g.probe) { g.probe[1] = mw; g.probe[3] = clock64() - t_start; }
#endif
        }
    } else {
        // Epilogue warps 2..5 own TMEM lane quarters (warp % 4).  N % 16 == 0, so a 32-column
        // chunk is either fully valid, half valid (16 columns) or out of range: all register
        // indices below are compile-time (nothing spills to local memory).
        constexpr int ACT = EPI == 2 ? PA_ACT_RELU : (EPI == 3 ? PA_ACT_GELU : PA_ACT_NONE);
        const int qtr = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
#ifdef PA_GEMM_PROBE
        t_epi0 = clock64();
#endif
        for (int mt = 0; mt < m_tiles; ++mt) {
            const int row = m0 + mt * BM + qtr * 32 + lane;
            const bool row_ok = row < g.M;
            const int64_t out_row = ((int64_t)batch * g.M + row) * g.N;
            // per-row dynamic activation scale (batch_quantize's scales[], int8_quant.cpp:15-28)
            const float alpha = (g.a_qscale && row_ok) ? __fdiv_rn(g.alpha, __ldg(g.a_qscale + (int64_t)batch * g.M + row))
                                                       : g.alpha;
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; ++cc) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(qtr * 32) << 16) + mt * BN + cc * 32, r);
                const int col0 = n0 + cc * 32;
                if (!row_ok || col0 >= g.N) continue;
                const bool hi_ok = col0 + 32 <= g.N;  // else exactly 16 valid columns
                if (g.acc_ws) {  // split-K: this split's partial tile, 128 B per thread per chunk
                    int32_t* wp = g.acc_ws + (int64_t)split * g.BATCH_rows * g.N + out_row + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (j < 16 || hi_ok)
                            *reinterpret_cast<int4*>(wp + j) = make_int4((int)r[j], (int)r[j + 1], (int)r[j + 2], (int)r[j + 3]);
                    continue;
                }
                if (g.C32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (j < 16 || hi_ok)
                            *reinterpret_cast<int4*>(g.C32 + out_row + col0 + j) =
                                make_int4((int)r[j], (int)r[j + 1], (int)r[j + 2], (int)r[j + 3]);
                }
                if (EPI != 0 && g.Cf) {
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        if (w < 4 || hi_ok) {
                            float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (g.bias) bj = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + w);
                            *reinterpret_cast<float4*>(g.Cf + out_row + col0 + 4 * w) =
                                make_float4(epilogue_f32<ACT>((int)r[4 * w + 0], alpha, bj.x),
                                            epilogue_f32<ACT>((int)r[4 * w + 1], alpha, bj.y),
                                            epilogue_f32<ACT>((int)r[4 * w + 2], alpha, bj.z),
                                            epilogue_f32<ACT>((int)r[4 * w + 3], alpha, bj.w));
                        }
                    }
                } else if (EPI != 0) {
                    uint32_t packed[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (g.bias && (w < 4 || hi_ok)) bj = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + w);
                        packed[w] = (uint32_t)(epilogue_one<ACT>((int)r[4 * w + 0], alpha, bj.x) & 0xff) |
                                    ((uint32_t)(epilogue_one<ACT>((int)r[4 * w + 1], alpha, bj.y) & 0xff) << 8) |
                                    ((uint32_t)(epilogue_one<ACT>((int)r[4 * w + 2], alpha, bj.z) & 0xff) << 16) |
                                    ((uint32_t)(epilogue_one<ACT>((int)r[4 * w + 3], alpha, bj.w) & 0xff) << 24);
                    }
                    *reinterpret_cast<uint4*>(g.C8 + out_row + col0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    if (hi_ok)
                        *reinterpret_cast<uint4*>(g.C8 + out_row + col0 + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                }
            }
        }
    }
#ifdef PA_GEMM_PROBE
    if (warp == 2 && lane == 0 && blockIdx.x == 5 && g.probe) {
        g.probe[5] = t_epi0 - t_setup;       // setup -> accumulators ready
        g.probe[6] = clock64() - t_epi0;     // epilogue
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();  // no CTA exits while peers may still signal its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------- 2-CTA kernel (M > 128)
// CTA pair (cluster of 2) = one UMMA M=256 x N=256 tile: tcgen05.mma.cta_group::2.kind::i8 issued
// by the leader reads A rows 0-127 / 128-255 and B columns 0-127 / 128-255 from the two CTAs'
// shared memories, so per 128-byte K block each SM ingests 16 KB of A + 16 KB of B for
// 128 x 256 x 128 MACs -- 1.5x less L2->SM traffic per MAC than the single-CTA 2x(128x128) form,
// which is what bounds this GEMM (measured: the chip-wide L2 output cap, ~43 B/clk/SM, not HBM or
// the tensor pipe; profiles/r01_gemm_notes.md).  6 stages x 32 KB ring per CTA; both CTAs'
// TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor.cta_group::2), the
// leader's tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs.
// Epilogue: 8 warps (two per TMEM lane quarter, 128 columns each).
constexpr int BN2 = 256;                      // columns per CTA pair
constexpr int STAGES2 = 6;
constexpr int STAGE2_BYTES = 2 * TILE_BYTES;  // A half (128 rows) + B half (128 columns)
constexpr int NTHREADS2 = 320;
constexpr uint32_t kIdesc2 = (2u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) |
                             ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_remote_arrive_expect_tx(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                 uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_i8_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// cta_group::2 load delivered to every CTA of `mask` at the same offset; each copy's bytes complete on the full
// barrier of the destination's PAIR LEADER (the barrier address names the even CTA of the issuer's pair).
__device__ __forceinline__ void tma_load_3d_2cta_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                    uint32_t bar_cluster_addr, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}

// OCC = 2 (used when the grid has more CTA pairs than one wave): two CTAs per SM with a 3-stage ring each (same
// bytes in flight per SM, 2 x 256 TMEM columns), so the epilogue, set-up and pipeline fill of one tile run under
// the main loop of the other CTA's tile.
// CLP = 2: a cluster of FOUR CTAs = two pairs on adjacent column slabs of the same rows.  Both pairs need the
// same A tile, so every CTA fetches 64 of its 128 A rows and multicasts them to its counterpart in the other
// pair: A crosses the L2 -> SM fabric once per cluster instead of once per pair (at M = 256 the whole GEMM is
// bound by that fabric, profiles/r01_gemm_notes.md).  A ring slot is then reusable only when BOTH pairs have
// consumed it (empty barrier count 2, each leader's commit multicast to all four CTAs).
// DQ = true: the s8 output is the dynamically quantised row (compute_minmax_scale + batch_quantize,
// int8_quant.cpp:59-64, 15-28) of act(alpha_row * acc + bias) -- the "fc1 -> int8_quant -> fc2" hand-off of the INT8
// decoder MLP without the f32 round trip through memory (16 MiB written and re-read per layer at M = 256).  A row's
// scale needs its maximum over ALL N columns, which are spread over every CTA pair of the grid, so the epilogue runs in
// two passes over accumulators that simply STAY IN TMEM: pass 1 computes the f32 values and folds the row maxima into
// global memory (atomicMax on the bit patterns), a grid-wide barrier follows (every CTA is resident: the host only
// uses this instance when the grid fits one wave), pass 2 re-reads TMEM, quantises with the final scale and stores
// s8.  Values, scale and rounding are bit-identical to the unfused kernels.
template <int EPI, int OCC, int CLP, bool DQ = false>
__global__ void __launch_bounds__(NTHREADS2, OCC)
gemm_i8_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args g) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int NST = OCC == 2 ? 3 : STAGES2;
    asm volatile("griddepcontrol.launch_dependents;");  // a split-K reduction launched behind this grid may become resident
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + NST * STAGE2_BYTES;  // full[S], empty[S], tmem_full
    const uint32_t tmem_slot = bar0 + (2 * NST + 1) * 8;
    auto full_bar = [&](int s) { return bar0 + s * 8; };
    auto empty_bar = [&](int s) { return bar0 + (NST + s) * 8; };
    const uint32_t tmem_full_bar = bar0 + 2 * NST * 8;

#ifdef PA_GEMM_PROBE
    const long long t_entry = clock64();
    const bool probe_cta = blockIdx.x == 10 && blockIdx.y == 0 && g.probe;
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cluster_rank = cluster_ctarank();
    const uint32_t crank = cluster_rank & 1u;   // rank inside the CTA pair (0 = leader)
    const uint32_t cpair = cluster_rank >> 1;   // pair inside the cluster (CLP = 2)
    const uint32_t lead_rank = cluster_rank & ~1u;
    const uint16_t pair_mask = (uint16_t)(3u << (2 * cpair));
    const uint16_t all_mask = (uint16_t)((1u << (2 * CLP)) - 1u);
    const int n0 = (blockIdx.x >> 1) * BN2;
    const int m_chunk = blockIdx.y / g.ksplit, split = blockIdx.y % g.ksplit;
    const int m0 = m_chunk * 2 * BM;
    const int batch = blockIdx.z;
    const int total_kb = (g.K + BK - 1) / BK;
    const int kb0 = split * g.kb_per_split;
    const int kb1 = min(total_kb, kb0 + g.kb_per_split);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full_bar(s), 2);   // one arrive.expect_tx from each CTA's producer (leader's copy is used)
            mbar_init(empty_bar(s), CLP);  // each pair leader's commit, multicast to every CTA of the cluster
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB));
    }
    if (warp == 1) tmem_alloc2(tmem_slot, BN2);
    tc_fence_before();
    __syncthreads();
    cluster_sync();  // both CTAs' barriers initialised and TMEM allocated before any cross-CTA signal
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // Launched with programmatic stream serialisation: barriers, TMEM and the cluster handshake above were set up while the
    // kernel in front of this one in the stream was still draining; its results (A, the split-K workspace) are touched below.
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef PA_GEMM_PROBE
    const long long t_setup = clock64();
    if (threadIdx.x == 0 && probe_cta) g.probe[4] = t_setup - t_entry;
    long long pw = 0, mw = 0;
#endif

    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NST;
#ifdef PA_GEMM_PROBE
                const long long tw0 = clock64();
#endif
                mbar_wait(empty_bar(s), ((i / NST) & 1) ^ 1);
#ifdef PA_GEMM_PROBE
                pw += clock64() - tw0;
#endif
                const uint32_t st = base + s * STAGE2_BYTES;
                const uint32_t lead_full = mapa_shared(full_bar(s), lead_rank);
                if (crank == 0) mbar_arrive_expect_tx(full_bar(s), STAGE2_BYTES);
                else mbar_remote_arrive_expect_tx(lead_full, STAGE2_BYTES);
                const int k0 = (kb0 + i) * BK;
                if (CLP == 1) {
                    tma_load_3d_2cta(st, &tmA, k0, m0 + (int)crank * BM, batch, lead_full);
                } else {  // tmA's box holds 64 rows: this CTA's share of the A half both pairs need
                    tma_load_3d_2cta_mc(st + cpair * (TILE_BYTES / 2), &tmA, k0, m0 + (int)crank * BM + (int)cpair * (BM / 2),
                                        batch, lead_full, (uint16_t)((1u << crank) | (1u << (2 + crank))));
                }
                tma_load_3d_2cta(st + TILE_BYTES, &tmB, n0 + (int)crank * BN, k0, batch, lead_full);
            }
#ifdef PA_GEMM_PROBE
            if (probe_cta) { g.probe[0] = pw; g.probe[2] = clock64() - t_setup; }
#endif
        }
    } else if (warp == 1) {
        if (crank == 0 && elect_one()) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NST;
#ifdef PA_GEMM_PROBE
                const long long tw0 = clock64();
#endif
                mbar_wait(full_bar(s), (i / NST) & 1);
#ifdef PA_GEMM_PROBE
                mw += clock64() - tw0;
#endif
                tc_fence_after();
                const uint32_t st = base + s * STAGE2_BYTES;
#pragma unroll
                for (int ks = 0; ks < BK / UK; ++ks) {
                    const uint64_t da = make_desc(st + ks * UK, 16, 1024);
                    const uint64_t db = make_desc(st + TILE_BYTES + ks * UK * BK, TILE_BYTES, 1024);
                    umma_i8_2cta(tmem_base, da, db, kIdesc2, (i > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit_2cta(empty_bar(s), all_mask);
            }
            umma_commit_2cta(tmem_full_bar, pair_mask);
#ifdef PA_GEMM_PROBE
            if (probe_cta) { g.probe[1] = mw; g.probe[3] = clock64() - t_setup; }
#endif
        }
    } else {
        constexpr int ACT = EPI == 2 ? PA_ACT_RELU : (EPI == 3 ? PA_ACT_GELU : PA_ACT_NONE);
        const int qtr = warp & 3;             // TMEM lane quarter this warp may read
        const int chalf = (warp - 2) >> 2;    // which 128 columns of the 256
        // While the main loop runs: stage this warp's 128 bias values in shared memory (the epilogue
        // then reads them as broadcasts instead of 8 dependent L2 round trips per chunk).
        float* bias_sm = reinterpret_cast<float*>(smem_raw + (bar0 + 128 - smem_u32(smem_raw))) + (warp - 2) * 128;  // 16 B aligned
        if (EPI != 0) {
            const int c = n0 + chalf * 128 + lane * 4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.bias && c < g.N) bv = __ldg(reinterpret_cast<const float4*>(g.bias + c));
            *reinterpret_cast<float4*>(bias_sm + lane * 4) = bv;
            __syncwarp();
        }
        const int row = m0 + (int)crank * BM + qtr * 32 + lane;
        const bool row_ok = row < g.M;
        const int64_t out_row = ((int64_t)batch * g.M + row) * g.N;
        const float alpha = (g.a_qscale && row_ok) ? __fdiv_rn(g.alpha, __ldg(g.a_qscale + (int64_t)batch * g.M + row))
                                                   : g.alpha;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
#ifdef PA_GEMM_PROBE
        const long long t_epi0 = clock64();
#endif
        if (DQ) {
            const int64_t grow = (int64_t)batch * g.M + row;
            const int cbase = n0 + chalf * 128;
            // ---- pass 1: row maxima of |act(alpha * acc + bias)| over this thread's 128 columns ----
            float vmax = 0.f;
#pragma unroll 1
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t rr[2][32];
                const uint32_t taddr = tmem_base + ((uint32_t)(qtr * 32) << 16) + chalf * 128 + c2 * 64;
                tmem_ld_32x32_nowait(taddr, rr[0]);
                tmem_ld_32x32_nowait(taddr + 32, rr[1]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int cl = c2 * 64 + hc * 32 + j;
                        const float v = epilogue_f32<ACT>((int)rr[hc][j], alpha, bias_sm[cl]);
                        if (cbase + cl < g.N) vmax = fmaxf(vmax, fabsf(v));
                    }
                }
            }
            // CTA-level maximum first (the two column halves of a row meet in shared memory; the ring is idle by now),
            // then ONE atomicMax per row and CTA
            float* rowmax_sm = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
            if (chalf == 1) rowmax_sm[qtr * 32 + lane] = vmax;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (chalf == 0 && row_ok) atomicMax(g.dq_rowmax + grow, __float_as_uint(fmaxf(vmax, rowmax_sm[qtr * 32 + lane])));
            // ---- grid barrier over the epilogue threads of every CTA ----
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == 2 && lane == 0) {
                atomicAdd(g.dq_barrier, 1u);
                unsigned int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(g.dq_barrier) : "memory");
                } while (seen < g.dq_expected);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- pass 2: quantise with the row's final scale ----
            float scale = 1.f;
            if (row_ok) {
                const float m = __uint_as_float(__ldcg(g.dq_rowmax + grow));
                scale = __fdiv_rn(127.f, __fadd_rn(m, 1e-6f));          // compute_minmax_scale
                if (n0 == 0 && chalf == 0) g.dq_scales[grow] = scale;
            }
            // every thread of this CTA holds its scale: the last CTA of the grid to get here clears the words
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == 2 && lane == 0) {
                if (atomicAdd(g.dq_barrier + 1, 1u) + 1u == g.dq_expected) {
                    for (int64_t i = 0; i < g.BATCH_rows; ++i) g.dq_rowmax[i] = 0u;
                    g.dq_barrier[0] = 0u;
                    g.dq_barrier[1] = 0u;
                }
            }
#pragma unroll 1
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t rr[2][32];
                const uint32_t taddr = tmem_base + ((uint32_t)(qtr * 32) << 16) + chalf * 128 + c2 * 64;
                tmem_ld_32x32_nowait(taddr, rr[0]);
                tmem_ld_32x32_nowait(taddr + 32, rr[1]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    const int col0 = cbase + c2 * 64 + hc * 32;
                    if (!row_ok || col0 >= g.N) continue;
                    const bool hi_ok = col0 + 32 <= g.N;
                    uint32_t packed[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        uint32_t b4[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float v = epilogue_f32<ACT>((int)rr[hc][4 * w + t], alpha, bias_sm[c2 * 64 + hc * 32 + 4 * w + t]);
                            b4[t] = quantize_half_away_s8(__fmul_rn(v, scale));   // batch_quantize
                        }
                        packed[w] = pack_s8x4(b4[0], b4[1], b4[2], b4[3]);
                    }
                    *reinterpret_cast<uint4*>(g.C8 + out_row + col0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    if (hi_ok)
                        *reinterpret_cast<uint4*>(g.C8 + out_row + col0 + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                }
            }
        } else
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
            // two 32-column chunks per TMEM round trip
            uint32_t rr[2][32];
            const uint32_t taddr = tmem_base + ((uint32_t)(qtr * 32) << 16) + chalf * 128 + c2 * 64;
            tmem_ld_32x32_nowait(taddr, rr[0]);
            tmem_ld_32x32_nowait(taddr + 32, rr[1]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (g.acc_ws) {
                // Split-K partial tile: transpose the warp's 32 rows x 64 columns through shared memory
                // (the ring is idle once the accumulator is complete) so that every store instruction
                // writes 256 contiguous bytes of 2 rows instead of 16 bytes of 32 rows.
                uint32_t* tsm = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw))) + (warp - 2) * (32 * 65);
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) tsm[lane * 65 + hc * 32 + j] = rr[hc][j];
                }
                __syncwarp();
                const int colbase = n0 + chalf * 128 + c2 * 64;
                const int rbase = m0 + (int)crank * BM + qtr * 32;
#pragma unroll 4
                for (int it = 0; it < 16; ++it) {
                    const int rl = it * 2 + (lane >> 4);
                    const int cw = (lane & 15) * 4;
                    const int grow = rbase + rl;
                    if (grow < g.M && colbase + cw < g.N) {
                        const uint32_t* sp = tsm + rl * 65 + cw;
                        int32_t* wp = g.acc_ws + (int64_t)split * g.BATCH_rows * g.N +
                                      ((int64_t)batch * g.M + grow) * g.N + colbase + cw;
                        __stcg(reinterpret_cast<int4*>(wp), make_int4((int)sp[0], (int)sp[1], (int)sp[2], (int)sp[3]));
                    }
                }
                __syncwarp();
                continue;
            }
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
                uint32_t(&r)[32] = rr[hc];
                const int cc = c2 * 2 + hc;
                const int col0 = n0 + chalf * 128 + cc * 32;
                if (!row_ok || col0 >= g.N) continue;
                const bool hi_ok = col0 + 32 <= g.N;
                if (g.acc_ws) continue;  // split-K partial tiles: coalesced path below
                if (g.C32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (j < 16 || hi_ok)
                            *reinterpret_cast<int4*>(g.C32 + out_row + col0 + j) =
                                make_int4((int)r[j], (int)r[j + 1], (int)r[j + 2], (int)r[j + 3]);
                }
                if (EPI != 0 && g.Cf) {
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        if (w < 4 || hi_ok) {
                            const float4 bj = *reinterpret_cast<const float4*>(bias_sm + cc * 32 + 4 * w);
                            *reinterpret_cast<float4*>(g.Cf + out_row + col0 + 4 * w) =
                                make_float4(epilogue_f32<ACT>((int)r[4 * w + 0], alpha, bj.x),
                                            epilogue_f32<ACT>((int)r[4 * w + 1], alpha, bj.y),
                                            epilogue_f32<ACT>((int)r[4 * w + 2], alpha, bj.z),
                                            epilogue_f32<ACT>((int)r[4 * w + 3], alpha, bj.w));
                        }
                    }
                } else if (EPI != 0) {
                    uint32_t packed[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        const float4 bj = *reinterpret_cast<const float4*>(bias_sm + cc * 32 + 4 * w);
                        packed[w] = pack_s8x4(epilogue_s8<ACT>((int)r[4 * w + 0], alpha, bj.x),
                                              epilogue_s8<ACT>((int)r[4 * w + 1], alpha, bj.y),
                                              epilogue_s8<ACT>((int)r[4 * w + 2], alpha, bj.z),
                                              epilogue_s8<ACT>((int)r[4 * w + 3], alpha, bj.w));
                    }
                    *reinterpret_cast<uint4*>(g.C8 + out_row + col0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    if (hi_ok)
                        *reinterpret_cast<uint4*>(g.C8 + out_row + col0 + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                }
            }
        }
#ifdef PA_GEMM_PROBE
        if (warp == 2 && lane == 0 && probe_cta) {
            g.probe[5] = t_epi0 - t_setup;
            g.probe[6] = clock64() - t_epi0;
        }
#endif
    }
#ifdef PA_GEMM_PROBE
    if (warp == 2 && lane == 0 && probe_cta) {
        // (epilogue warps only reach here with t_epi0 defined)
    }
#endif
    tc_fence_before();
    __syncthreads();
    cluster_sync();  // the peer may still be reading its accumulators / signalling this CTA's barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, BN2);
    }
}

// ---------------------------------------------------------------- persistent 2-CTA kernel (more tiles than one wave)
// One CTA pair per two SMs (74 pairs), each looping over 256 x 256 tiles; TMEM holds TWO accumulators (2 x 256 of
// the 512 columns) so the 8 epilogue warps drain tile j while the UMMA thread already fills tile j+1, and the TMA
// ring simply keeps running across tile boundaries.  Against one tile per CTA (two CTAs per SM, OCC = 2) this removes
// the wave quantisation: fc1 at M = 2048 is 512 tiles = 3.46 waves of 148 resident pairs (the last wave 46 % full),
// here 6.92 rounds of 74 pairs at full per-pair speed.  Tiles are numbered m-chunk fastest, so the pairs running at
// the same time share B slabs (weights cross HBM once, the re-reads hit L2).
// Barriers: full / empty per ring stage as above; tmem_full[2] (UMMA commit -> both CTAs' epilogue warps);
// tmem_empty[2] on the LEADER (16 arrivals: 8 epilogue warps of each CTA, the peer's by remote arrive).
// Every wait is bounded: a protocol error becomes a trap, not a hung GPU.
__device__ __forceinline__ void mbar_wait_trap(uint32_t bar, uint32_t parity) {
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++n > (1u << 27)) __trap();
    }
}
__device__ __forceinline__ void mbar_remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(NTHREADS2, 1)
gemm_i8_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args g,
                       int n_slabs, int m_chunks, int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int NST = STAGES2;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + NST * STAGE2_BYTES;  // full[NST], empty[NST], tmem_full[2], tmem_empty[2]
    auto full_bar = [&](int s) { return bar0 + s * 8; };
    auto empty_bar = [&](int s) { return bar0 + (NST + s) * 8; };
    auto tfull_bar = [&](int b) { return bar0 + (2 * NST + b) * 8; };
    auto tempty_bar = [&](int b) { return bar0 + (2 * NST + 2 + b) * 8; };
    const uint32_t tmem_slot = bar0 + (2 * NST + 4) * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank() & 1u;
    const int pid = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int total_kb = (g.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full_bar(s), 2);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 16);
        }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB));
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 2 * BN2);
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto tile_coords = [&](int t, int& n0, int& m0, int& batch) {
        const int per_batch = n_slabs * m_chunks;
        batch = t / per_batch;
        const int rem = t - batch * per_batch;
        n0 = (rem / m_chunks) * BN2;
        m0 = (rem % m_chunks) * 2 * BM;
    };

    if (warp == 0) {
        if (elect_one()) {
            uint32_t it = 0;
            for (int t = pid; t < total_tiles; t += npairs) {
                int n0, m0, batch;
                tile_coords(t, n0, m0, batch);
                for (int i = 0; i < total_kb; ++i, ++it) {
                    const int s = it % NST;
                    mbar_wait_trap(empty_bar(s), ((it / NST) & 1) ^ 1);
                    const uint32_t st = base + s * STAGE2_BYTES;
                    const uint32_t lead_full = mapa_shared(full_bar(s), 0);
                    if (crank == 0) mbar_arrive_expect_tx(full_bar(s), STAGE2_BYTES);
                    else mbar_remote_arrive_expect_tx(lead_full, STAGE2_BYTES);
                    const int k0 = i * BK;
                    tma_load_3d_2cta(st, &tmA, k0, m0 + (int)crank * BM, batch, lead_full);
                    tma_load_3d_2cta(st + TILE_BYTES, &tmB, n0 + (int)crank * BN, k0, batch, lead_full);
                }
            }
        }
    } else if (warp == 1) {
        if (crank == 0 && elect_one()) {
            uint32_t it = 0, j = 0;
            for (int t = pid; t < total_tiles; t += npairs, ++j) {
                const uint32_t buf = j & 1u;
                mbar_wait_trap(tempty_bar(buf), ((j >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
                tc_fence_after();
                for (int i = 0; i < total_kb; ++i, ++it) {
                    const int s = it % NST;
                    mbar_wait_trap(full_bar(s), (it / NST) & 1);
                    tc_fence_after();
                    const uint32_t st = base + s * STAGE2_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BK / UK; ++ks) {
                        const uint64_t da = make_desc(st + ks * UK, 16, 1024);
                        const uint64_t db = make_desc(st + TILE_BYTES + ks * UK * BK, TILE_BYTES, 1024);
                        umma_i8_2cta(tmem_base + buf * BN2, da, db, kIdesc2, (i > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit_2cta(empty_bar(s), 3);
                }
                umma_commit_2cta(tfull_bar(buf), 3);
            }
        }
    } else {
        constexpr int ACT = EPI == 2 ? PA_ACT_RELU : (EPI == 3 ? PA_ACT_GELU : PA_ACT_NONE);
        const int qtr = warp & 3;
        const int chalf = (warp - 2) >> 2;
        float* bias_sm = reinterpret_cast<float*>(smem_raw + (bar0 + 256 - smem_u32(smem_raw))) + (warp - 2) * 128;  // past the barriers + TMEM slot
        const uint32_t lead_tempty0 = mapa_shared(tempty_bar(0), 0);
        uint32_t j = 0;
        for (int t = pid; t < total_tiles; t += npairs, ++j) {
            int n0, m0, batch;
            tile_coords(t, n0, m0, batch);
            const uint32_t buf = j & 1u;
            if (EPI != 0) {
                const int c = n0 + chalf * 128 + lane * 4;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.bias && c < g.N) bv = __ldg(reinterpret_cast<const float4*>(g.bias + c));
                __syncwarp();
                *reinterpret_cast<float4*>(bias_sm + lane * 4) = bv;
                __syncwarp();
            }
            const int row = m0 + (int)crank * BM + qtr * 32 + lane;
            const bool row_ok = row < g.M;
            const int64_t out_row = ((int64_t)batch * g.M + row) * g.N;
            const float alpha = (g.a_qscale && row_ok) ? __fdiv_rn(g.alpha, __ldg(g.a_qscale + (int64_t)batch * g.M + row))
                                                       : g.alpha;
            mbar_wait_trap(tfull_bar(buf), (j >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t rr[2][32];
                const uint32_t taddr = tmem_base + ((uint32_t)(qtr * 32) << 16) + buf * BN2 + chalf * 128 + c2 * 64;
                tmem_ld_32x32_nowait(taddr, rr[0]);
                tmem_ld_32x32_nowait(taddr + 32, rr[1]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    uint32_t(&r)[32] = rr[hc];
                    const int cc = c2 * 2 + hc;
                    const int col0 = n0 + chalf * 128 + cc * 32;
                    if (!row_ok || col0 >= g.N) continue;
                    const bool hi_ok = col0 + 32 <= g.N;
                    if (g.C32) {
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 4)
                            if (jj < 16 || hi_ok)
                                *reinterpret_cast<int4*>(g.C32 + out_row + col0 + jj) =
                                    make_int4((int)r[jj], (int)r[jj + 1], (int)r[jj + 2], (int)r[jj + 3]);
                    }
                    if (EPI != 0 && g.Cf) {
#pragma unroll
                        for (int w = 0; w < 8; ++w) {
                            if (w < 4 || hi_ok) {
                                const float4 bj = *reinterpret_cast<const float4*>(bias_sm + cc * 32 + 4 * w);
                                *reinterpret_cast<float4*>(g.Cf + out_row + col0 + 4 * w) =
                                    make_float4(epilogue_f32<ACT>((int)r[4 * w + 0], alpha, bj.x),
                                                epilogue_f32<ACT>((int)r[4 * w + 1], alpha, bj.y),
                                                epilogue_f32<ACT>((int)r[4 * w + 2], alpha, bj.z),
                                                epilogue_f32<ACT>((int)r[4 * w + 3], alpha, bj.w));
                            }
                        }
                    } else if (EPI != 0) {
                        uint32_t packed[8];
#pragma unroll
                        for (int w = 0; w < 8; ++w) {
                            const float4 bj = *reinterpret_cast<const float4*>(bias_sm + cc * 32 + 4 * w);
                            packed[w] = pack_s8x4(epilogue_s8<ACT>((int)r[4 * w + 0], alpha, bj.x),
                                                  epilogue_s8<ACT>((int)r[4 * w + 1], alpha, bj.y),
                                                  epilogue_s8<ACT>((int)r[4 * w + 2], alpha, bj.z),
                                                  epilogue_s8<ACT>((int)r[4 * w + 3], alpha, bj.w));
                        }
                        *reinterpret_cast<uint4*>(g.C8 + out_row + col0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        if (hi_ok)
                            *reinterpret_cast<uint4*>(g.C8 + out_row + col0 + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                    }
                }
            }
            // this warp has read its part of accumulator `buf`: tell the leader's UMMA thread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_remote_arrive(lead_tempty0 + buf * 8);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 2 * BN2);
    }
}

// Split-K second pass: sum the ksplit partial tiles (exact int32), then C32 copy and/or C8 epilogue.
template <int ACT>
__global__ void gemm_i8_splitk_epilogue_kernel(const Args g, int64_t rows) {
    // launched with programmatic stream serialisation behind the GEMM: resident early, blocks here until the partial
    // tiles are complete and visible (a no-op for an ordinary launch); the GEMM behind it may set itself up meanwhile
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int64_t n4 = rows * g.N / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        int4 a = __ldcs(reinterpret_cast<const int4*>(g.acc_ws) + i);
        for (int sp = 1; sp < g.ksplit; ++sp) {
            const int4 b = __ldcs(reinterpret_cast<const int4*>(g.acc_ws) + sp * n4 + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (g.C32) reinterpret_cast<int4*>(g.C32)[i] = a;
        if (g.C8 || g.Cf) {
            const int col = (int)((i * 4) % g.N);
            const float alpha = g.a_qscale ? __fdiv_rn(g.alpha, __ldg(g.a_qscale + (i * 4) / g.N)) : g.alpha;
            float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.bias) bj = __ldg(reinterpret_cast<const float4*>(g.bias + col));
            if (g.Cf)
                reinterpret_cast<float4*>(g.Cf)[i] =
                    make_float4(epilogue_f32<ACT>(a.x, alpha, bj.x), epilogue_f32<ACT>(a.y, alpha, bj.y),
                                epilogue_f32<ACT>(a.z, alpha, bj.z), epilogue_f32<ACT>(a.w, alpha, bj.w));
            else
                reinterpret_cast<uint32_t*>(g.C8)[i] =
                    (uint32_t)(epilogue_one<ACT>(a.x, alpha, bj.x) & 0xff) |
                    ((uint32_t)(epilogue_one<ACT>(a.y, alpha, bj.y) & 0xff) << 8) |
                    ((uint32_t)(epilogue_one<ACT>(a.z, alpha, bj.z) & 0xff) << 16) |
                    ((uint32_t)(epilogue_one<ACT>(a.w, alpha, bj.w) & 0xff) << 24);
        }
    }
}

// ---- host: tensor maps --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 3-D u8 tensor [batch][rows][cols] (cols contiguous), box [1][box_rows][128 B], 128B swizzle.
static bool make_map(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint64_t batch,
                     uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {cols, rows, batch};
    cuuint64_t strides[2] = {cols, cols * rows};  // bytes, dims 1..2
    cuuint32_t box[3] = {128, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tile / split-K geometry shared by the launcher and pa_gemm_i8_workspace_bytes.
struct Plan {
    bool two_cta;
    int n_slabs, m_chunks, ksplit, kb_per_split;
};
static Plan make_plan(int BATCH, int M, int N, int K, int sm_count) {
    Plan p;
    p.two_cta = M > BM && !(getenv("PA_GEMM_2CTA") && atoi(getenv("PA_GEMM_2CTA")) == 0);
    p.n_slabs = p.two_cta ? (N + BN2 - 1) / BN2 : (N + BN - 1) / BN;
    p.m_chunks = (M + 2 * BM - 1) / (2 * BM);
    const int total_kb = (K + BK - 1) / BK;
    const int64_t ctas = (int64_t)p.n_slabs * p.m_chunks * BATCH * (p.two_cta ? 2 : 1);
    int ksplit = (int)(sm_count / ctas);
    if (ksplit > total_kb / 8) ksplit = total_kb / 8;  // >= 8 K blocks (1 KiB of K) per split
    if (ksplit < 1) ksplit = 1;
    p.kb_per_split = (total_kb + ksplit - 1) / ksplit;
    p.ksplit = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
    return p;
}

}  // namespace gemm
}  // namespace pa

using namespace pa;
using namespace pa::gemm;

static int gemm_i8_launch(const int8_t* d_A, const int8_t* d_B, int8_t* d_C_s8, float* d_C_f32, int32_t* d_C_s32,
                          int BATCH, int M, int N, int K, float alpha_host, const float* d_a_qscale,
                          const float* d_bias, int act, void* d_workspace, size_t workspace_bytes,
                          pa_stream_t stream, float* d_dq_scales = nullptr) {
    PA_CHECK_ARG(d_A && d_B && (d_C_s8 || d_C_s32 || d_C_f32));
    PA_CHECK_ARG(BATCH > 0 && M > 0 && N > 0 && K > 0);
    PA_CHECK_ARG(!d_C_f32 || (uintptr_t)d_C_f32 % 16 == 0);
    PA_CHECK_ARG(act == PA_ACT_NONE || act == PA_ACT_RELU || act == PA_ACT_GELU);
    if (K % 16 != 0 || N % 16 != 0) return PA_ERR_UNSUPPORTED;
    PA_CHECK_ARG((uintptr_t)d_A % 16 == 0 && (uintptr_t)d_B % 16 == 0);
    PA_CHECK_ARG(!d_C_s8 || (uintptr_t)d_C_s8 % 16 == 0);
    PA_CHECK_ARG(!d_C_s32 || (uintptr_t)d_C_s32 % 16 == 0);
    const DeviceInfo& di = device_info();
    if (!di.ok) return PA_ERR_NO_DEVICE;
    cudaStream_t st = as_stream(stream);

    const Plan plan = make_plan(BATCH, M, N, K, di.sm_count);
    const bool two_cta = plan.two_cta;
    const int n_slabs = plan.n_slabs, m_chunks = plan.m_chunks;
    const int m_tiles0 = M > BM ? 2 : 1;
    // Cluster multicast of A needs one M chunk (uniform m_tiles) and n_slabs divisible by CL.
    int CLs = 1;
    // Measured (profiles/r01_gemm_notes.md): equal speed for CL = 1/2/4 at the C4 shapes -- the
    // loop is bound by B bytes in flight, not by L2 reads -- so default to pairs, which cut the
    // L2 read traffic for A in half and never strand SMs.  PA_GEMM_CLUSTER=1|2|4 overrides.
    if (m_chunks == 1 && !two_cta) {
        if (n_slabs % 4 == 0) CLs = 4;
        else if (n_slabs % 2 == 0) CLs = 2;
    }
    const int cl_max = CLs;
    if (CLs > 2) CLs = 2;
    const char* cl_env = getenv("PA_GEMM_CLUSTER");
    if (cl_env && !two_cta) {
        const int want = atoi(cl_env);
        if (want == 1 || (want == 2 && cl_max >= 2) || (want == 4 && cl_max == 4)) CLs = want;
    }
    // 1-CTA kernel ring geometry.  M <= 128: load only ceil(M/8)*8 rows of A per stage and use the
    // saved shared memory for a deeper ring (more weight bytes in flight: small-M GEMMs are pure weight
    // streaming); no cluster multicast in that case.
    int a_rows_box = BM, a_bytes = m_tiles0 * TILE_BYTES, stages = STAGES;
    if (!two_cta && M <= BM) {
        CLs = 1;
        a_rows_box = ((M + 7) / 8) * 8;
        a_bytes = a_rows_box * BK;
        stages = RING_BYTES / (a_bytes + TILE_BYTES);
        if (stages > MAXST) stages = MAXST;
    } else if (!two_cta && CLs > 1) {
        a_rows_box = m_tiles0 * BM / CLs;
    }
    CUtensorMap tmA, tmB;
    const bool mc_pairs = two_cta && n_slabs % 2 == 0 && !(getenv("PA_GEMM_MC") && atoi(getenv("PA_GEMM_MC")) == 0);
    if (!make_map(&tmA, d_A, (uint64_t)K, (uint64_t)M, (uint64_t)BATCH, two_cta ? (mc_pairs ? BM / 2 : BM) : a_rows_box))
        return PA_ERR_UNSUPPORTED;
    if (!make_map(&tmB, d_B, (uint64_t)N, (uint64_t)K, (uint64_t)BATCH, BK)) return PA_ERR_UNSUPPORTED;

    Args g{};
    g.C8 = d_C_s8;
    g.Cf = d_C_f32;
    g.a_qscale = d_a_qscale;
    g.C32 = d_C_s32;
    g.bias = d_bias;
    g.alpha = alpha_host;
    g.act = act;
    g.M = M; g.N = N; g.K = K;
    g.stages = stages;
    g.a_bytes = a_bytes;
#ifdef PA_GEMM_PROBE
    extern unsigned long long* pa_gemm_probe_buf;
    g.probe = pa_gemm_probe_buf;
#endif
    // Split-K partial tiles live in the CALLER's workspace (pa_gemm_i8_workspace_bytes): the library owns no
    // growable scratch, so pointers baked into a captured CUDA graph stay valid and streams never share it.
    // Without a (large enough) workspace the GEMM runs unsplit: same result, fewer CTAs.
    const int64_t rows = (int64_t)BATCH * M;
    int ksplit = plan.ksplit;
    g.kb_per_split = plan.kb_per_split;
    if (ksplit > 1) {
        const size_t need = (size_t)ksplit * rows * N * sizeof(int32_t);
        if (d_workspace && workspace_bytes >= need && (uintptr_t)d_workspace % 16 == 0) {
            g.acc_ws = static_cast<int32_t*>(d_workspace);
        } else {
            ksplit = 1;
            g.kb_per_split = (K + BK - 1) / BK;
        }
    }
    g.ksplit = ksplit;
    const bool dq = d_dq_scales != nullptr;
    if (dq) {
        // fused output quantisation: one wave of CTA pairs, no split-K, activation none / relu (see the DQ instance)
        const int64_t ctas = (int64_t)n_slabs * m_chunks * BATCH * 2;
        if (!two_cta || plan.ksplit != 1 || ctas > di.sm_count || act == PA_ACT_GELU) return PA_ERR_UNSUPPORTED;
        const size_t need = 16 + (size_t)rows * sizeof(unsigned int);
        if (!d_workspace || workspace_bytes < need || (uintptr_t)d_workspace % 16 != 0) return PA_ERR_WORKSPACE;
        g.dq_barrier = static_cast<unsigned int*>(d_workspace);  // zero once (caller); every launch leaves it zero
        g.dq_rowmax = g.dq_barrier + 4;
        g.dq_scales = d_dq_scales;
        g.dq_expected = (unsigned int)ctas;
    }
    g.BATCH_rows = rows;
    const int epi = (ksplit > 1 || !(d_C_s8 || d_C_f32)) ? 0 : 1 + act;
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const Args);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e;
    const int64_t tiles2 = (int64_t)n_slabs * m_chunks * BATCH;
    const bool persist = two_cta && !dq && ksplit == 1 && tiles2 > di.sm_count / 2 && tiles2 < (1ll << 30) &&
                         !(getenv("PA_GEMM_PERSIST") && atoi(getenv("PA_GEMM_PERSIST")) == 0);
    if (persist) {
        // more tiles than one wave of CTA pairs: persistent pairs with double-buffered TMEM accumulators
        CUtensorMap tmA2;
        if (!make_map(&tmA2, d_A, (uint64_t)K, (uint64_t)M, (uint64_t)BATCH, BM)) return PA_ERR_UNSUPPORTED;
        using PKernelFn = void (*)(const CUtensorMap, const CUtensorMap, const Args, int, int, int);
        static const PKernelFn pk[4] = {gemm_i8_persist_kernel<0>, gemm_i8_persist_kernel<1>, gemm_i8_persist_kernel<2>,
                                        gemm_i8_persist_kernel<3>};
        PKernelFn kern = pk[epi];
        const size_t smemp = (size_t)STAGES2 * STAGE2_BYTES + 256 + 8 * 128 * sizeof(float) + 1024;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp);
        if (e != cudaSuccess) return (int)e;
        const int npairs = (int)(tiles2 < di.sm_count / 2 ? tiles2 : di.sm_count / 2);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * npairs));
        cfg.blockDim = dim3(NTHREADS2);
        cfg.dynamicSmemBytes = smemp;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tmA2, tmB, g, n_slabs, m_chunks, (int)tiles2);
        if (e != cudaSuccess) return (int)e;
        e = cudaGetLastError();
        return e == cudaSuccess ? PA_OK : (int)e;
    }
    if (two_cta) {
        // more CTA pairs than one wave (148 SMs = 74 pairs): two CTAs per SM, see the kernel's header
        const int64_t pairs = (int64_t)n_slabs * m_chunks * ksplit * BATCH;
        const int occ2 = (pairs > di.sm_count / 2 && !(getenv("PA_GEMM_OCC2") && atoi(getenv("PA_GEMM_OCC2")) == 0)) ? 1 : 0;
        const size_t smem2 = (size_t)(occ2 ? 3 : STAGES2) * STAGE2_BYTES + 128 + 8 * 128 * sizeof(float) + 1024;  // ring, barriers, bias, align
        // cluster of two pairs sharing A by multicast when the slabs pair up (PA_GEMM_MC=0 turns it off)
        const int clp2 = mc_pairs ? 1 : 0;
        static const KernelFn kernels2[2][2][4] = {
            {{gemm_i8_2cta_kernel<0, 1, 1>, gemm_i8_2cta_kernel<1, 1, 1>, gemm_i8_2cta_kernel<2, 1, 1>, gemm_i8_2cta_kernel<3, 1, 1>},
             {gemm_i8_2cta_kernel<0, 2, 1>, gemm_i8_2cta_kernel<1, 2, 1>, gemm_i8_2cta_kernel<2, 2, 1>, gemm_i8_2cta_kernel<3, 2, 1>}},
            {{gemm_i8_2cta_kernel<0, 1, 2>, gemm_i8_2cta_kernel<1, 1, 2>, gemm_i8_2cta_kernel<2, 1, 2>, gemm_i8_2cta_kernel<3, 1, 2>},
             {gemm_i8_2cta_kernel<0, 2, 2>, gemm_i8_2cta_kernel<1, 2, 2>, gemm_i8_2cta_kernel<2, 2, 2>, gemm_i8_2cta_kernel<3, 2, 2>}}};
        KernelFn kern = kernels2[clp2][occ2][epi];
        if (dq) {
            static const KernelFn kernels_dq[2][2] = {
                {gemm_i8_2cta_kernel<1, 1, 1, true>, gemm_i8_2cta_kernel<2, 1, 1, true>},
                {gemm_i8_2cta_kernel<1, 1, 2, true>, gemm_i8_2cta_kernel<2, 1, 2, true>}};
            kern = kernels_dq[clp2][act == PA_ACT_RELU ? 1 : 0];
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e != cudaSuccess) return (int)e;
        }
        static bool attr_set2[64][2][2][4] = {};
        if (!dq && !attr_set2[dev & 63][clp2][occ2][epi]) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e != cudaSuccess) return (int)e;
            attr_set2[dev & 63][clp2][occ2][epi] = true;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * n_slabs), (unsigned)(m_chunks * ksplit), (unsigned)BATCH);
        cfg.blockDim = dim3(NTHREADS2);
        cfg.dynamicSmemBytes = smem2;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = clp2 ? 4u : 2u;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (dq) {
            // the epilogue contains a grid-wide barrier: a COOPERATIVE launch makes the driver guarantee (or refuse)
            // that every CTA is resident at once instead of leaving it to the grid-size check above
            attr[1].id = cudaLaunchAttributeCooperative;
            attr[1].val.cooperative = 1;
            cfg.numAttrs = 2;
        } else if (!(getenv("PA_GEMM_PDL") && atoi(getenv("PA_GEMM_PDL")) == 0)) {
            // set-up (barriers, TMEM, cluster handshake) under the tail of the kernel in front: see griddepcontrol.wait
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.numAttrs = 2;
        }
        e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, g);
        if (e != cudaSuccess) return (int)e;
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    } else {
        const size_t smem = (size_t)RING_BYTES + (2 * MAXST + 1) * 8 + 16 + 1024;
        static const KernelFn kernels[3][4] = {
            {gemm_i8_kernel<0, 1>, gemm_i8_kernel<1, 1>, gemm_i8_kernel<2, 1>, gemm_i8_kernel<3, 1>},
            {gemm_i8_kernel<0, 2>, gemm_i8_kernel<1, 2>, gemm_i8_kernel<2, 2>, gemm_i8_kernel<3, 2>},
            {gemm_i8_kernel<0, 4>, gemm_i8_kernel<1, 4>, gemm_i8_kernel<2, 4>, gemm_i8_kernel<3, 4>}};
        const int cli = CLs == 4 ? 2 : (CLs == 2 ? 1 : 0);
        KernelFn kern = kernels[cli][epi];
        static bool attr_set[64][3][4] = {};
        if (!attr_set[dev & 63][cli][epi]) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set[dev & 63][cli][epi] = true;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)n_slabs, (unsigned)(m_chunks * ksplit), (unsigned)BATCH);
        cfg.blockDim = dim3(NTHREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CLs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (!(getenv("PA_GEMM_PDL") && atoi(getenv("PA_GEMM_PDL")) == 0)) {
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.numAttrs = 2;
        }
        e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, g);
        if (e != cudaSuccess) return (int)e;
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    if (ksplit > 1) {
        const int64_t n4 = rows * N / 4;
        int blocks = (int)((n4 + 255) / 256);
        if (blocks > di.sm_count * 8) blocks = di.sm_count * 8;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)blocks);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see the kernel: it waits for the GEMM itself
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (act == PA_ACT_RELU) e = cudaLaunchKernelEx(&cfg, gemm_i8_splitk_epilogue_kernel<PA_ACT_RELU>, g, rows);
        else if (act == PA_ACT_GELU) e = cudaLaunchKernelEx(&cfg, gemm_i8_splitk_epilogue_kernel<PA_ACT_GELU>, g, rows);
        else e = cudaLaunchKernelEx(&cfg, gemm_i8_splitk_epilogue_kernel<PA_ACT_NONE>, g, rows);
        if (e != cudaSuccess) return (int)e;
    }
    return PA_OK;
}

PA_API size_t pa_gemm_i8_workspace_bytes(int BATCH, int M, int N, int K) {
    if (BATCH <= 0 || M <= 0 || N <= 0 || K <= 0) return 0;
    const DeviceInfo& di = device_info();
    const Plan p = make_plan(BATCH, M, N, K, di.ok ? di.sm_count : 148);
    return p.ksplit > 1 ? (size_t)p.ksplit * BATCH * M * N * sizeof(int32_t) : 0;
}

PA_API int pa_gemm_i8(const int8_t* d_A, const int8_t* d_B, int8_t* d_C_s8, int32_t* d_C_s32, int BATCH, int M,
                      int N, int K, float scaleA, float scaleB, float scaleC, const float* d_bias, int act,
                      void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_C_s8 || d_C_s32);
    PA_CHECK_ARG(scaleC != 0.f);
    // dnnl_matmul_int8.cpp:40, fp32, left to right
    return gemm_i8_launch(d_A, d_B, d_C_s8, nullptr, d_C_s32, BATCH, M, N, K, scaleA * scaleB / scaleC, nullptr,
                          d_bias, act, d_workspace, workspace_bytes, stream);
}

PA_API int pa_gemm_i8_dequant(const int8_t* d_A, const int8_t* d_B, float* d_C_f32, int BATCH, int M, int N, int K,
                              const float* d_a_qscale, float b_dequant, const float* d_bias, int act,
                              void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_C_f32 && d_a_qscale);
    return gemm_i8_launch(d_A, d_B, nullptr, d_C_f32, nullptr, BATCH, M, N, K, b_dequant, d_a_qscale, d_bias, act,
                          d_workspace, workspace_bytes, stream);
}

PA_API size_t pa_gemm_i8_dynquant_workspace_bytes(int BATCH, int M, int N) {
    (void)N;
    if (BATCH <= 0 || M <= 0) return 0;
    return 16 + (size_t)BATCH * M * sizeof(unsigned int);
}

// "fc1 -> int8_quant" in one kernel (see the DQ instance of gemm_i8_2cta_kernel).  PA_ERR_UNSUPPORTED when the shape
// does not fit one wave of CTA pairs (M <= 128, more pairs than SMs / 2, split-K shapes, gelu): the caller then
// uses pa_gemm_i8_dequant + pa_row_quantize_dynamic_i8, which give the same bits.
PA_API int pa_gemm_i8_dynquant(const int8_t* d_A, const int8_t* d_B, int8_t* d_C_s8, float* d_c_qscale, int BATCH, int M,
                               int N, int K, const float* d_a_qscale, float b_dequant, const float* d_bias, int act,
                               void* d_workspace, size_t workspace_bytes, pa_stream_t stream) {
    PA_CHECK_ARG(d_C_s8 && d_c_qscale && d_a_qscale);
    return gemm_i8_launch(d_A, d_B, d_C_s8, nullptr, nullptr, BATCH, M, N, K, b_dequant, d_a_qscale, d_bias, act,
                          d_workspace, workspace_bytes, stream, d_c_qscale);
}
