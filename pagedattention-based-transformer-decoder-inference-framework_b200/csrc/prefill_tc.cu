// prefill_tc.cu -- flash-attention prefill over the paged cache on the 5th-generation tensor cores
// (tcgen05.mma kind::f16, accumulators in TMEM), fp16 and int8 pages, head_dim 64 / 128.  SURVEY 8(f) row 1.
//
// One CTA = 256 consecutive query positions of one (row, head), handled as TWO query tiles of 128 rows
// (A, B) that share every K/V tile; KV is consumed in tiles of 64 tokens (4 page units of 16 tokens, each
// located through the page table):
//   warps 0-3   : softmax group A, warps 4-7: softmax group B.  One thread per query row (= TMEM lane):
//                 tcgen05.ld its 64 scores, causal / context / unmapped-page mask, online softmax in registers,
//                 P written back to TMEM (packed fp16 over the first 32 columns of the scores it came from).
//   warp 8 (+9) : TMEM allocation (all 512 columns) + the elected thread(s) that issue the UMMAs:
//                   S_X[sb] (128 x 64, fp32, TMEM)  = Q_X (smem, fp16) . K^T        8 x (M128 N64  K16)
//                   O_X     (128 x 128, fp32, TMEM) += P_X (TMEM, fp16) . V (smem)   4 x (M128 N128 K16)
//                 fp16 pages: one issuing warp PER QUERY TILE (MW = 2; a thread's mbarrier waits queue behind its own
//                 tcgen05.commit, so a single issuer idles the pipe at every P hand-off); int8 pages: one issuer,
//                 staggered order S_A(i+1) | P.V_A(i) | S_B(i+1) | P.V_B(i).
//   next warp   : producer, one elected thread: fp16 pages = 16 TMA tensor boxes per tile (K tile as the K-major B
//                 operand of S, V tile as the MN-major B operand of P.V, 3-stage ring); int8 pages = bulk copies
//                 of raw units into a raw ring.
//   last 2 warps: (int8 pages) converters: raw units -> the same swizzled fp16 stage, 1/scale per token beside it.
// O accumulates in TMEM across all KV tiles (accumulate flag), scaled by a per-row REFERENCE maximum that
// is only raised when the tile maximum exceeds it by more than 8 (log2 units): P stays below 2^8 in fp16,
// l and O use the same reference so O / l is exact, and the row rescale (tcgen05.ld, multiply, tcgen05.st)
// happens a few times per row instead of once per tile.  S is double-buffered in TMEM so S(i+1) is computed
// while the softmax of tile i runs; P.V(i-1) is only waited for when a row rescale needs O.
// Design history, the experiments behind each choice and the rejected variants: profiles/r01_prefill_tc_notes.md.
#include <cstdlib>
#include <cstring>

#include "mma_utils.cuh"
#include "pa_common.cuh"

namespace pa {
namespace ptc {

constexpr int QT = 128;               // queries per query tile (= TMEM lanes)
constexpr int KT = 64;                // tokens per KV tile
// HD = head_dim (64 or 128) = HD / 64 blocks of 64 dims (128 B of fp16 per row, the SWIZZLE_128B span):
__host__ __device__ constexpr int q_bytes(int hd) { return (hd / 64) * QT * 128; }  // per query tile: k-blocks of [128 rows x 128 B]
__host__ __device__ constexpr int k_bytes(int hd) { return (hd / 64) * KT * 128; }  // K: k-blocks of [64 token rows x 128 B];
                                                                                    // V: n-blocks of the same shape
__host__ __device__ constexpr int stage_bytes(int hd) { return 2 * k_bytes(hd); }
// NQ = query tiles per CTA.  NQ = 2 (default): warps 0-3 softmax A, 4-7 softmax B, then UMMA issuer(s), producer, 3-stage ring,
// all 512 TMEM columns (tile X: S0 [128X, +64), S1 [128X+64, +64); O_X [256 + 128X, +128)), one CTA per SM.
// NQ = 1 (fp16 pages only): warps 0-3 softmax, 4 UMMA, 5 producer, 2-stage ring, 256 TMEM columns, TWO CTAs per
// SM -- K/V bytes are staged once per 128 queries instead of 256, but the prologue / epilogue of one CTA runs
// under the main loop of the other.
__host__ __device__ constexpr int stages_for(int nq) { return nq == 2 ? 3 : 2; }
// barriers: kv_full/kv_empty[ST]; per query tile s_full[2], p_full[2], o_full[2], q_ready
__host__ __device__ constexpr int nbar_for(int nq) { return 2 * stages_for(nq) + 7 * nq; }
// int8 pages: two converter warps (with two UMMA issuers that makes 13 warps = 128 registers per thread, a few
// spilled words; measured no faster than one issuer, so int8 defaults to MW = 1)
__host__ __device__ constexpr int conv_warps(int kv, int mw) { return kv ? 2 : 0; }
__host__ __device__ constexpr int threads_for(int kv, int nq, int mw = 1) { return (4 * nq + mw + 1 + conv_warps(kv, mw)) * 32; }
// int8 pages (KV = 1): the producer bulk-copies RAW units (16 tokens: 2 KB of K, 2 KB of V, 16 + 16 f32 scales)
// into a raw ring, the converter warps (12 warps in all still get 168 registers) rewrite them as the same swizzled
// fp16 stage the fp16 path gets from TMA (exact: PRMT to 1024 + u, HSUB2) and leave 1/scale per token next to it; the softmax threads apply
// the K scale to the score columns and fold the V scale into P (int8_quant.cpp:46-57: x = q / scale).
constexpr int RS = 3;                 // raw ring stages (one 64-token tile each)
__host__ __device__ constexpr int raw_unit(int hd) { return 2 * 16 * hd + 64 + 64; }  // K rows, V rows, k scales, v scales
constexpr int SCALE_BYTES = 2 * KT * 4;  // per fp16 stage: 1/k_scale[64], 1/v_scale[64]

struct Args {
    const float* q;
    float* out;
    const int32_t* table;
    const int32_t* beam_ids;
    const int32_t* ctx_start;
    int num_beams, H, num_tiles, total_pages, B, Tq, tile_size;
    float qscale;
    const int8_t* k8;      // int8 pools and their per-(page, token) scales (KV = 1)
    const int8_t* v8;
    const float* k_scales;
    const float* v_scales;
    int stagger;       // UMMA issue order, see the UMMA warp
    int upt_shift;     // log2(16-token units per page)
    int total_tokens;  // rows of the pool tensor map: a box at this row is all zeros (out-of-bounds fill)
    // element strides of q and out (same layout for both): batch, head, query position.  [B, H, Tq, D] is
    // (H*Tq*D, Tq*D, D); the decoders' token-major activations [B, Tq, H, D] are (Tq*H*D, D, H*D).
    int64_t sb, sh, st;
};

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in TMEM: row m of A = TMEM lane m, 32-bit column j = (A[m][2j], A[m][2j+1]) as packed fp16
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_128B (same encoding as gemm_i8.cu make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: c = F32 (1) [4,6), a = b = F16 (0), a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
// Every wait of this kernel is bounded: a protocol error becomes a trap (CUDA error at the next
// synchronisation) after ~1 s instead of a hung GPU.
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++n > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_cnt(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// -DPA_PTC_PROBE: clock64 stamps of one softmax warp per group and of the UMMA warp for 64 steady-state KV
// tiles of CTA 5, read back with pa_debug_ptc_probe (benchmarks/prefill_probe.py).  Off in the product build.
#ifdef PA_PTC_PROBE
__device__ long long g_probe[3 * 64 * 8];
#define PROBE(role, it, k)                                                                          \
    do {                                                                                            \
        if (blockIdx.x == 5 && (threadIdx.x & 31) == 0 && (it) >= 16 && (it) < 80)                  \
            g_probe[((role) * 64 + (it) - 16) * 8 + (k)] = clock64();                               \
    } while (0)
// life of CTA 5 seen from warp 0: entry | setup done | Q staged | first S seen | KV loop done | last P.V seen |
// O stored | exit
__device__ long long g_probe_cta[8];
#define PROBE_CTA(k)                                                                                \
    do {                                                                                            \
        if (blockIdx.x == 5 && threadIdx.x == 0) g_probe_cta[k] = clock64();                        \
    } while (0)
#else
#define PROBE(role, it, k) do { } while (0)
#define PROBE_CTA(k) do { } while (0)
#endif

// MW = UMMA-issuing warps.  MW = 2 (fp16 pages, two query tiles): one issuing warp per query tile.  A thread's
// mbarrier waits queue behind its own tcgen05.commit, so a single issuer sees every "P ready?" wait only after the
// MMAs it has just issued have drained (~250 cycles of idle tensor pipe per wait, clock64 probes); with one issuer
// per tile those waits run under the other tile's MMAs.
template <int KV, int NQ, int HD, int MW = 1>
__global__ void __launch_bounds__(threads_for(KV, NQ, MW), (NQ == 1 && KV == 0) ? 2 : 1) prefill_tc_kernel(const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV,
                                                                 const Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int ST = stages_for(NQ), NBAR = nbar_for(NQ), TMEM_COLS = NQ * 256;  // (HD = 64 leaves columns unused)
    constexpr int D = HD, KB = HD / 64;
    constexpr int Q_BYTES = q_bytes(HD), K_BYTES = k_bytes(HD), STAGE = stage_bytes(HD);
    constexpr int RAW_UNIT = raw_unit(HD), RAW_STAGE = 4 * RAW_UNIT;
    constexpr int W_MMA = 4 * NQ, W_PROD = 4 * NQ + MW, W_CONV = 4 * NQ + MW + 1;  // warp roles after the softmax groups
    PROBE_CTA(0);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_sm = base;
    const uint32_t kv_sm = q_sm + NQ * Q_BYTES;
    const uint32_t raw_sm = kv_sm + ST * STAGE;                       // KV = 1 only
    const uint32_t scale_sm = raw_sm + (KV ? RS * RAW_STAGE : 0);     // KV = 1 only
    const uint32_t bar0 = scale_sm + (KV ? ST * SCALE_BYTES : 0);
    auto kv_full = [&](int s) { return bar0 + s * 8; };
    auto kv_empty = [&](int s) { return bar0 + (ST + s) * 8; };
    // s_full, p_full and o_full alternate between two barriers (tile i -> barrier i & 1, phase i >> 1): a parity
    // wait is only unambiguous while the barrier is at most ONE phase ahead of the waiter, and with S computed
    // one tile ahead the softmax threads and the UMMA thread can be a tile apart in either direction.
    auto s_full = [&](int x, int sb) { return bar0 + (2 * ST + 7 * x + sb) * 8; };
    auto p_full = [&](int x, int par) { return bar0 + (2 * ST + 7 * x + 2 + par) * 8; };
    auto o_full = [&](int x, int par) { return bar0 + (2 * ST + 7 * x + 4 + par) * 8; };
    auto q_ready = [&](int x) { return bar0 + (2 * ST + 7 * x + 6) * 8; };
    auto raw_full = [&](int s) { return bar0 + (NBAR + s) * 8; };
    auto raw_empty = [&](int s) { return bar0 + (NBAR + RS + s) * 8; };
    const uint32_t tmem_slot = bar0 + (NBAR + 2 * RS) * 8;
    const uint32_t meta_sm = tmem_slot + 8;  // KV = 1: per raw stage, bit u = unit u of the tile was copied

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) {
            mbar_init(kv_full(s), KV ? conv_warps(KV, MW) : 1);  // KV = 1: one arrival per converter warp
            mbar_init(kv_empty(s), MW);
        }
        if (KV) {
            for (int s = 0; s < RS; ++s) {
                mbar_init(raw_full(s), 1);
                mbar_init(raw_empty(s), conv_warps(KV, MW));
            }
        }
        for (int x = 0; x < NQ; ++x) {
            mbar_init(s_full(x, 0), 1);
            mbar_init(s_full(x, 1), 1);
            mbar_init(p_full(x, 0), QT);
            mbar_init(p_full(x, 1), QT);
            mbar_init(o_full(x, 0), 1);
            mbar_init(o_full(x, 1), 1);
            mbar_init(q_ready(x), QT);
        }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV));
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    PROBE_CTA(1);

    constexpr int QC = NQ * QT;  // queries per CTA
    const int nqt = (a.Tq + QC - 1) / QC;
    const int bh_total = a.B * a.H;
    // CTA order: tile-major, longest tiles first (measured better than keeping the query tiles of one
    // (row, head) adjacent for L2 reuse)
    const int qt = nqt - 1 - (int)(blockIdx.x / bh_total);
    const int bh = (int)(blockIdx.x % bh_total);
    const int b = bh / a.H, h = bh - b * a.H;
    const int start = a.ctx_start ? a.ctx_start[b] : 0;
    const int cap = a.num_tiles * a.tile_size;
    // per query tile X: keys [0, kmax_X) are visible to its last query; n_X KV tiles; an absent tile has n = 0
    int kmaxs[NQ], nts[NQ];
#pragma unroll
    for (int x = 0; x < NQ; ++x) {
        const int q0 = qt * QC + x * QT;
        const int q_last = min(a.Tq, q0 + QT) - 1;
        kmaxs[x] = q0 < a.Tq ? max(0, min(cap, start + q_last + 1)) : 0;
        nts[x] = (kmaxs[x] + KT - 1) / KT;
    }
    const int kmax_c = max(kmaxs[0], kmaxs[NQ - 1]);  // kmax is monotone in x, an absent tile has 0
    const int n_tiles = max(nts[0], nts[NQ - 1]);
    const int upt_mask = (1 << a.upt_shift) - 1;  // 16-token units per page - 1 (page sizes are 16 << k)
    const int beam = a.beam_ids ? a.beam_ids[b] : b;
    const int32_t* trow = ((unsigned)beam < (unsigned)a.num_beams)
                              ? a.table + ((int64_t)beam * a.H + h) * a.num_tiles : nullptr;

    if (KV == 1 && warp == W_PROD) {
        // ------------------------------------------------------------ producer, int8 pages: raw units by bulk copy
        if (elect_one()) {
            int* meta = reinterpret_cast<int*>(smem_raw + (meta_sm - smem_u32(smem_raw)));
            int rs = 0;
            uint32_t ph = 1;
            for (int i = 0; i < n_tiles; ++i) {
                int64_t row0[4];
                uint32_t mask = 0;
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int u = i * 4 + uu;
                    int page = (trow && u * 16 < kmax_c) ? __ldg(trow + (u >> a.upt_shift)) : -1;
                    if ((unsigned)page >= (unsigned)a.total_pages) page = -1;
                    row0[uu] = (int64_t)page * a.tile_size + (u & upt_mask) * 16;  // token row in the pools
                    if (page >= 0) mask |= 1u << uu;
                }
                mbar_wait_wd(raw_empty(rs), ph);
                meta[rs] = (int)mask;
                mbar_arrive_expect_tx(raw_full(rs), (uint32_t)__popc(mask) * RAW_UNIT);
                const uint32_t dst = raw_sm + rs * RAW_STAGE;
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    if ((mask >> uu) & 1u) {
                        const uint32_t d = dst + uu * RAW_UNIT;
                        bulk_g2s_nohint(d, a.k8 + row0[uu] * D, 16 * D, raw_full(rs));
                        bulk_g2s_nohint(d + 16 * D, a.v8 + row0[uu] * D, 16 * D, raw_full(rs));
                        bulk_g2s_nohint(d + 32 * D, a.k_scales + row0[uu], 64, raw_full(rs));
                        bulk_g2s_nohint(d + 32 * D + 64, a.v_scales + row0[uu], 64, raw_full(rs));
                    }
                }
                if (++rs == RS) {
                    rs = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (KV == 1 && warp >= W_CONV) {
        // ------------------------------------------------------------ converters: raw int8 units -> fp16 stage
        // warp cw rewrites units 2 cw and 2 cw + 1 of every tile; lane = (token row, 64-dim half).  K and V have
        // the same shared-memory image ([64-dim half][token row of 128 B], SWIZZLE_128B); only the UMMA
        // descriptors read them differently.
        const int* meta = reinterpret_cast<const int*>(smem_raw + (meta_sm - smem_u32(smem_raw)));
        const int cw = warp - W_CONV;
        const int tr = lane >> 1, hf = lane & 1;
        const __half2 off = __floats2half2_rn(1152.f, 1152.f);
        int rs = 0, fs = 0;
        uint32_t rph = 0, fph = 1;
        for (int i = 0; i < n_tiles; ++i) {
            mbar_wait_wd(raw_full(rs), rph);
            mbar_wait_wd(kv_empty(fs), fph);
            const uint32_t mask = (uint32_t)meta[rs];
            const uint32_t dst = kv_sm + fs * STAGE;
#pragma unroll
            for (int k2 = 0; k2 < 4 / conv_warps(1, MW); ++k2) {
                const int uu = (4 / conv_warps(1, MW)) * cw + k2;
                const uint32_t src = raw_sm + rs * RAW_STAGE + uu * RAW_UNIT;
                const bool have = (mask >> uu) & 1u;
                const int r = uu * 16 + tr;  // token row inside the tile
#pragma unroll
                for (int kvsel = 0; kvsel < 2; ++kvsel) {
                    const uint32_t rowbase = dst + kvsel * K_BYTES + r * 128;
#pragma unroll
                    for (int c4 = 0; c4 < D / 32; ++c4) {  // 16 int8 -> two 16-byte chunks of 8 halfs
                        uint32_t hv[8];
                        if (have) {
                            const uint4 w = lds_128(src + kvsel * (16 * D) + tr * D + hf * (D / 2) + c4 * 16);
                            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint32_t xb = ww[e] ^ 0x80808080u;  // u = b + 128
                                uint32_t p01 = __byte_perm(xb, 0x64646464u, 0x4140);  // halfs (1024 + u0, 1024 + u1)
                                uint32_t p23 = __byte_perm(xb, 0x64646464u, 0x4342);
                                __half2 a01 = __hsub2(*reinterpret_cast<__half2*>(&p01), off);
                                __half2 a23 = __hsub2(*reinterpret_cast<__half2*>(&p23), off);
                                hv[2 * e] = *reinterpret_cast<uint32_t*>(&a01);
                                hv[2 * e + 1] = *reinterpret_cast<uint32_t*>(&a23);
                            }
                        } else {  // unmapped page / unit past the context: finite zeros (its scores are masked)
#pragma unroll
                            for (int e = 0; e < 8; ++e) hv[e] = 0u;
                        }
                        const int c = hf * (D / 16) + 2 * c4;  // 16-byte chunk of the fp16 row; 8 chunks per 64-dim block
                        const uint32_t boxrow = rowbase + (c >> 3) * 8192;
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(boxrow + ((((c) & 7) ^ (r & 7)) << 4)),
                                     "r"(hv[0]), "r"(hv[1]), "r"(hv[2]), "r"(hv[3]) : "memory");
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(boxrow + ((((c + 1) & 7) ^ (r & 7)) << 4)),
                                     "r"(hv[4]), "r"(hv[5]), "r"(hv[6]), "r"(hv[7]) : "memory");
                    }
                }
                // 1 / scale per token: lanes 0-15 the unit's K rows, 16-31 its V rows.  Tokens past the context
                // get 0 (their P is 0, and 0 x a garbage scale must not become NaN).
                float sc = 0.f;
                const int tok = i * KT + uu * 16 + (lane & 15);
                if (have && tok < kmax_c) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sc) : "r"(src + 32 * D + lane * 4));
                    sc = fast_rcp(sc);
                }
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(scale_sm + fs * SCALE_BYTES + (lane >> 4) * (KT * 4) +
                                                              (uu * 16 + (lane & 15)) * 4), "f"(sc) : "memory");
            }
            fence_proxy_async();  // the UMMA reads the stage through the async proxy
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(kv_full(fs));
                mbar_arrive(raw_empty(rs));
            }
            if (++rs == RS) {
                rs = 0;
                rph ^= 1u;
            }
            if (++fs == ST) {
                fs = 0;
                fph ^= 1u;
            }
        }
    } else if (warp == W_PROD) {
        // ------------------------------------------------------------ TMA producer
        // ONE elected thread issues the 16 boxes of a tile (unit uu x {K lo, K hi, V lo, V hi}).  Everything
        // under elect_one() stays in uniform registers, so the 16 UTMALDG go out back to back; page ids are
        // fetched one tile ahead.
        if (elect_one()) {
            int pgn[4];
            auto load_pages = [&](int tile) {
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int u = tile * 4 + uu;
                    int page = (trow && u * 16 < kmax_c) ? __ldg(trow + (u >> a.upt_shift)) : -1;
                    if ((unsigned)page >= (unsigned)a.total_pages) page = -1;
                    // unmapped page / unit past the context: a box outside the tensor is filled with zeros (finite
                    // operands; the softmax threads mask the scores of unmapped units themselves)
                    pgn[uu] = page >= 0 ? page * a.tile_size + (u & upt_mask) * 16 : a.total_tokens;
                }
            };
            load_pages(0);
            int s = 0;
            uint32_t ph = 1;
            for (int i = 0; i < n_tiles; ++i) {
                int row0[4];
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) row0[uu] = pgn[uu];
                if (i + 1 < n_tiles) load_pages(i + 1);
                mbar_wait_wd(kv_empty(s), ph);
                const uint32_t st = kv_sm + s * STAGE;
                mbar_arrive_expect_tx(kv_full(s), STAGE);
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) tma_load_2d(st + kb * 8192 + uu * 2048, &tmK, kb * 64, row0[uu], kv_full(s));
                }
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_2d(st + K_BYTES + kb * 8192 + uu * 2048, &tmV, kb * 64, row0[uu], kv_full(s));
                }
                if (++s == ST) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (MW == 2 && (warp == W_MMA || warp == W_MMA + 1)) {
        // ------------------------------------------------------------ UMMA issuers, one per query tile
        constexpr uint32_t kIdescS = idesc_f16(QT, KT, 0, 0);  // Q K-major, K K-major
        constexpr uint32_t kIdescO = idesc_f16(QT, D, 0, 1);   // P (TMEM), V MN-major
        const int x = warp - W_MMA;
        const int ntx = nts[x];
        auto issue_S = [&](int i, int stage) {  // S_x(i) into buffer i & 1 (see the single-issuer branch below)
            if (elect_one()) {
                const uint32_t st = kv_sm + stage * STAGE;
                const uint32_t qx = q_sm + x * Q_BYTES;
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint64_t da = make_desc(qx + (ks >> 2) * (QT * 128) + (ks & 3) * 32, 16, 1024);
                    const uint64_t db = make_desc(st + (ks >> 2) * (KT * 128) + (ks & 3) * 32, 16, 1024);
                    umma_f16(tmem_base + x * 2 * KT + (i & 1) * KT, da, db, kIdescS, ks > 0 ? 1u : 0u);
                }
                umma_commit(s_full(x, i & 1));
            }
            __syncwarp();
        };
        if (n_tiles > 0) {
            mbar_wait_wd(kv_full(0), 0);
            if (ntx > 0) {
                mbar_wait_wd(q_ready(x), 0);
                tc_fence_after();
                issue_S(0, 0);
            }
        }
        int s = 0;
        uint32_t kv_ph = 0;
        for (int i = 0; i < n_tiles; ++i) {
            const int s_next = (s + 1 == ST) ? 0 : s + 1;
            if (x == 0) PROBE(2, i, 0);  // issuer of tile A: loop top | S(i+1) issued | P(i) seen | P.V(i) issued
            if (i + 1 < n_tiles) {
                // every stage is waited for by both issuers, also the ones a finished tile no longer reads: its
                // arrival on kv_empty below must fall into the phase of THIS use of the stage
                const uint32_t ph_next = (s + 1 == ST) ? (kv_ph ^ 1u) : kv_ph;
                mbar_wait_wd(kv_full(s_next), ph_next);
                tc_fence_after();
                if (i + 1 < ntx) issue_S(i + 1, s_next);  // runs under the softmax of tile i
            }
            if (x == 0) PROBE(2, i, 1);
            if (i < ntx) {
                const uint32_t st = kv_sm + s * STAGE;
                const int nvalid_c = min(KT, kmax_c - i * KT);
                if (KV == 0 && nvalid_c < KT && (nvalid_c & 15)) {
                    // rows of the last page past the context end may hold anything (0 x NaN = NaN): zero them
                    // (both issuers may do this for the same tile: same zeros)
                    const int r0 = nvalid_c;
                    const int r1 = (nvalid_c + 15) & ~15;
                    for (int idx = lane; idx < (r1 - r0) * (D / 8); idx += 32) {
                        const int r = r0 + idx / (D / 8), c = idx % (D / 8);
                        const uint32_t addr = st + K_BYTES + (c >> 3) * 8192 + r * 128 + (((c & 7) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
                    }
                    fence_proxy_async();
                    __syncwarp();
                }
                mbar_wait_wd(p_full(x, i & 1), (uint32_t)((i >> 1) & 1));
                if (x == 0) PROBE(2, i, 2);
                tc_fence_after();
                if (elect_one()) {
                    const int nvalid = min(KT, kmaxs[x] - i * KT);
                    const int ksteps = (nvalid + 15) >> 4;  // tokens past the causal limit contribute nothing
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t ta = tmem_base + x * 2 * KT + (i & 1) * KT + ks * 8;  // P(i), packed fp16
                        const uint64_t db = make_desc(st + K_BYTES + ks * 2048, 8192, 1024);
                        umma_f16_ts(tmem_base + NQ * 2 * KT + x * D, ta, db, kIdescO, (i > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(o_full(x, i & 1));
                }
                __syncwarp();
                if (x == 0) PROBE(2, i, 3);
            }
            if (elect_one()) umma_commit(kv_empty(s));  // arrives once this issuer's reads of the stage (if any) are done
            __syncwarp();
            if (++s == ST) {
                s = 0;
                kv_ph ^= 1u;
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------ UMMA issuer
        constexpr uint32_t kIdescS = idesc_f16(QT, KT, 0, 0);  // Q K-major, K K-major
        constexpr uint32_t kIdescO = idesc_f16(QT, D, 0, 1);   // P K-major, V MN-major
        // S_X(i) into buffer i & 1 of tile X.  The buffer was last read (tcgen05.ld) by the softmax of tile
        // i - 2, and this thread has already waited for p_full_X(i - 2), which every softmax thread signals
        // after its tcgen05.wait::ld: no separate "S buffer empty" barrier is needed.
        auto issue_S = [&](int x, int i, int stage) {
            if (elect_one()) {
                const uint32_t st = kv_sm + stage * STAGE;
                const uint32_t qx = q_sm + x * Q_BYTES;
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {  // k-steps of 16 dims: k-block ks>>2, 32 B per step inside it
                    const uint64_t da = make_desc(qx + (ks >> 2) * (QT * 128) + (ks & 3) * 32, 16, 1024);
                    const uint64_t db = make_desc(st + (ks >> 2) * (KT * 128) + (ks & 3) * 32, 16, 1024);
                    umma_f16(tmem_base + x * 2 * KT + (i & 1) * KT, da, db, kIdescS, ks > 0 ? 1u : 0u);
                }
                umma_commit(s_full(x, i & 1));
            }
            __syncwarp();
        };
        // Both query tiles at once, k-steps interleaved (A0 B0 A1 B1 ...): consecutive MMAs accumulate into
        // different TMEM tiles and share the K operand descriptor.  Measured gain is small (16 MMAs: 1202 -> 1107
        // cycles, clock64 probes): these N = 64 instructions cost ~70 cycles each whatever their order.
        auto issue_S2 = [&](int i, int stage) {
            if (elect_one()) {
                const uint32_t st = kv_sm + stage * STAGE;
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint64_t db = make_desc(st + (ks >> 2) * (KT * 128) + (ks & 3) * 32, 16, 1024);
#pragma unroll
                    for (int x = 0; x < NQ; ++x) {
                        const uint64_t da = make_desc(q_sm + x * Q_BYTES + (ks >> 2) * (QT * 128) + (ks & 3) * 32, 16, 1024);
                        umma_f16(tmem_base + x * 2 * KT + (i & 1) * KT, da, db, kIdescS, ks > 0 ? 1u : 0u);
                    }
                }
                umma_commit(s_full(0, i & 1));
                umma_commit(s_full(NQ - 1, i & 1));
            }
            __syncwarp();
        };
        if (n_tiles > 0) {
            mbar_wait_wd(kv_full(0), 0);
#pragma unroll
            for (int x = 0; x < NQ; ++x) {
                if (nts[x] > 0) {
                    mbar_wait_wd(q_ready(x), 0);
                    tc_fence_after();
                    issue_S(x, 0, 0);
                }
            }
        }
        int s = 0;
        uint32_t kv_ph = 0;
        for (int i = 0; i < n_tiles; ++i) {
            // S(i+1) as soon as its K tile has landed, so it overlaps the softmax of tile i
            const int s_next = (s + 1 == ST) ? 0 : s + 1;
            PROBE(2, i, 0);
            if (i + 1 < n_tiles) {
                const uint32_t ph_next = (s + 1 == ST) ? (kv_ph ^ 1u) : kv_ph;
                mbar_wait_wd(kv_full(s_next), ph_next);
                tc_fence_after();
                if (NQ == 2 && a.stagger) {
                    if (i + 1 < nts[0]) issue_S(0, i + 1, s_next);  // S_B(i+1) follows P.V_A(i), below
                } else if (NQ == 2 && i + 1 < nts[0] && i + 1 < nts[NQ - 1]) {
                    issue_S2(i + 1, s_next);
                } else {
#pragma unroll
                    for (int x = 0; x < NQ; ++x)
                        if (i + 1 < nts[x]) issue_S(x, i + 1, s_next);
                }
            }
            PROBE(2, i, 1);
            const uint32_t st = kv_sm + s * STAGE;
            const int nvalid_c = min(KT, kmax_c - i * KT);
            if (KV == 0 && nvalid_c < KT && (nvalid_c & 15)) {
                // rows of the last page past the context end may hold anything (0 x NaN = NaN): zero them
                const int r0 = nvalid_c;
                const int r1 = (nvalid_c + 15) & ~15;
                for (int idx = lane; idx < (r1 - r0) * (D / 8); idx += 32) {
                    const int r = r0 + idx / (D / 8), c = idx % (D / 8);
                    const uint32_t addr = st + K_BYTES + (c >> 3) * 8192 + r * 128 + (((c & 7) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
            }
            // O_X += P_X(i) . V(i)   (the softmax threads finished any row rescale before arriving on p_full)
            auto issue_PV = [&](int x) {
                PROBE(2, i, 2 + 2 * x);
                tc_fence_after();
                if (elect_one()) {
                    const int nvalid = min(KT, kmaxs[x] - i * KT);
                    const int ksteps = (nvalid + 15) >> 4;  // tokens past the causal limit contribute nothing
                    for (int ks = 0; ks < ksteps; ++ks) {
                        // P(i) sits in the first 32 columns of the S buffer it was computed from
                        const uint32_t ta = tmem_base + x * 2 * KT + (i & 1) * KT + ks * 8;
                        const uint64_t db = make_desc(st + K_BYTES + ks * 2048, 8192, 1024);
                        umma_f16_ts(tmem_base + NQ * 2 * KT + x * D, ta, db, kIdescO, (i > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(o_full(x, i & 1));
                }
                __syncwarp();
                PROBE(2, i, 3 + 2 * x);
            };
            if (NQ == 2 && a.stagger) {
                // Staggered order: S_A(i+1) | P.V_A(i) | S_B(i+1) | P.V_B(i).  Group B's scores arrive half an
                // iteration after group A's, so the two groups are in their exponentials (MUFU-bound) at different
                // times instead of in lock-step, and each has a whole iteration of tensor work between "S ready"
                // and "P needed".
                if (i < nts[0]) {
                    mbar_wait_wd(p_full(0, i & 1), (uint32_t)((i >> 1) & 1));
                    issue_PV(0);
                }
                if (i + 1 < nts[1]) issue_S(1, i + 1, s_next);
                if (i < nts[1]) {
                    mbar_wait_wd(p_full(1, i & 1), (uint32_t)((i >> 1) & 1));
                    issue_PV(1);
                }
            } else {
                // in the order the groups deliver P
                uint32_t pend = 0;
#pragma unroll
                for (int x = 0; x < NQ; ++x) pend |= (i < nts[x] ? 1u : 0u) << x;
                uint32_t idle = 0;
                while (pend) {
                    if (++idle > (1u << 26)) __trap();
#pragma unroll
                    for (int x = 0; x < NQ; ++x) {
                        if (!((pend >> x) & 1u)) continue;
                        if (!__all_sync(0xffffffffu, mbar_try_wait(p_full(x, i & 1), (uint32_t)((i >> 1) & 1)))) continue;
                        pend &= ~(1u << x);
                        issue_PV(x);
                    }
                }
            }
            if (elect_one()) umma_commit(kv_empty(s));
            __syncwarp();
            if (++s == ST) {
                s = 0;
                kv_ph ^= 1u;
            }
        }
    } else {
        // ------------------------------------------------------------ softmax / accumulate: one thread per query
        const int x = warp >> 2;               // query tile of this warp's group
        const int qtr = warp & 3;              // TMEM lane quadrant
        const int row = qtr * 32 + lane;       // TMEM lane = row of the Q tile
        const int q0 = qt * QC + x * QT;
        const int t = q0 + row;                // query position within the prompt chunk
        const bool t_ok = t < a.Tq;
        const int qpos = t_ok ? start + t : -1;
        const int kmax = kmaxs[x], nt = nts[x];
        const uint32_t lane_base = (uint32_t)(qtr * 32) << 16;
        if (nt > 0) {
            // Q rows of this warp -> fp16, pre-scaled, into the swizzled K-major A tile.  One row per step, the
            // whole warp on its 512 bytes (coalesced); 16 rows in flight.
            const float* qbase = a.q + (int64_t)b * a.sb + (int64_t)h * a.sh;
            const uint32_t qx = q_sm + x * Q_BYTES;
            constexpr int LPR = D / 4;    // lanes per row (one float4 each); 128 / D rows per load instruction
            constexpr int RPI = 32 / LPR;
            const int lr = lane % LPR, rsub = lane / LPR;
            const int c = lr >> 1;  // 16-byte chunk (8 halfs) of the row this lane contributes to
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += 16 * RPI) {
                float4 v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int tr = q0 + qtr * 32 + r0 + j * RPI + rsub;
                    v[j] = tr < a.Tq ? __ldg(reinterpret_cast<const float4*>(qbase + (int64_t)tr * a.st) + lr)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int rr = qtr * 32 + r0 + j * RPI + rsub;
                    const uint32_t w0 = pack_half2(v[j].x * a.qscale, v[j].y * a.qscale);
                    const uint32_t w1 = pack_half2(v[j].z * a.qscale, v[j].w * a.qscale);
                    const uint32_t addr = qx + (c >> 3) * (QT * 128) + rr * 128 + (((c & 7) ^ (rr & 7)) << 4) + (lane & 1) * 8;
                    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(w0), "r"(w1) : "memory");
                }
            }
            fence_proxy_async();
            mbar_arrive_cnt(q_ready(x));
            PROBE_CTA(2);
        }
        float m_ref = -INFINITY, l_run = 0.f;  // reference maximum of this row (see the header), running sum
        const uint32_t s_addr = tmem_base + lane_base + x * 2 * KT;
        const uint32_t o_addr = tmem_base + lane_base + NQ * 2 * KT + x * D;
        // P.V(t') completion: tile t' -> barrier t' & 1, phase t' >> 1.  Every thread consumes the phases of each
        // barrier in order and at the latest two tiles late (P.V(t') can only be issued after this thread's own
        // arrival on p_full(t'), so the barrier is never more than one phase ahead of what is waited for).
        int oc[2] = {0, 0};
        auto ensure_pv = [&](int tp) {
            const int bsel = tp & 1, k = tp >> 1;
            if (oc[bsel] <= k) {
                mbar_wait_wd(o_full(x, bsel), (uint32_t)(k & 1));
                oc[bsel] = k + 1;
            }
        };
        int pgn[4];  // page ids of the next tile's 4 units
        auto load_pages = [&](int tile) {
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const int u = tile * 4 + uu;
                pgn[uu] = (trow && u * 16 < kmax) ? __ldg(trow + (u >> a.upt_shift)) : -1;
            }
        };
        load_pages(0);
        int sfs = 0;          // KV = 1: fp16 stage of tile i and the parity of its kv_full phase
        uint32_t sfph = 0;
#pragma unroll 1
        for (int i = 0; i < nt; ++i) {
            const int sb = i & 1;
            // which of the tile's 4 units are mapped (an unmapped page is skipped: ...fused.cu:32); the table
            // entries were fetched one tile ahead, so their latency is never waited for
            int pg[4];
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) pg[uu] = pgn[uu];
            load_pages(i + 1);
            if (qtr == 0) PROBE(x, i, 0);
            mbar_wait_wd(s_full(x, sb), (uint32_t)((i >> 1) & 1));
            tc_fence_after();
            if (qtr == 0) PROBE(x, i, 1);
            if (i == 0) PROBE_CTA(3);
            uint32_t sr[2][32];
            tmem_ld32(s_addr + sb * KT, sr[0]);
            tmem_ld32(s_addr + sb * KT + 32, sr[1]);
            tmem_wait_ld();
            if (KV == 1) {
                // int8 pages: score column k is (q . k_q[k]) / k_scale[k].  Waiting on the stage's own barrier
                // (long complete) makes the converter's scale writes visible to this thread.
                mbar_wait_wd(kv_full(sfs), sfph);
                const uint32_t ksc = scale_sm + sfs * SCALE_BYTES;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const uint4 f = lds_128(ksc + c * 16);
                    const int k0 = 4 * c;
                    sr[k0 >> 5][k0 & 31] = __float_as_uint(__uint_as_float(sr[k0 >> 5][k0 & 31]) * __uint_as_float(f.x));
                    sr[k0 >> 5][(k0 + 1) & 31] = __float_as_uint(__uint_as_float(sr[k0 >> 5][(k0 + 1) & 31]) * __uint_as_float(f.y));
                    sr[k0 >> 5][(k0 + 2) & 31] = __float_as_uint(__uint_as_float(sr[k0 >> 5][(k0 + 2) & 31]) * __uint_as_float(f.z));
                    sr[k0 >> 5][(k0 + 3) & 31] = __float_as_uint(__uint_as_float(sr[k0 >> 5][(k0 + 3) & 31]) * __uint_as_float(f.w));
                }
            }
            if (qtr == 0) PROBE(x, i, 2);
            // causal / context / unmapped-page mask (only tiles that need one) + tile maximum
            const int kp0 = i * KT;
            bool all_mapped = true;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) all_mapped &= (unsigned)pg[uu] < (unsigned)a.total_pages;
            const bool full_tile = (kp0 + KT - 1 <= qpos) && (kp0 + KT <= kmax) && all_mapped;
            if (!full_tile) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int kp = kp0 + hh * 32 + j;
                        const bool mapped = (unsigned)pg[(hh * 32 + j) >> 4] < (unsigned)a.total_pages;
                        if (!(kp <= qpos && kp < kmax && mapped)) sr[hh][j] = 0xff800000u;  // -inf
                    }
                }
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(sr[hh][j]));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // raise the reference maximum only when this tile exceeds it by more than 2^8
            const bool raise = mx > m_ref + 8.f || (m_ref == -INFINITY && mx > -INFINITY);
            const float m_new = raise ? mx : m_ref;
            const float corr = (raise && m_ref != -INFINITY) ? fast_exp2(m_ref - m_new) : 1.f;
            m_ref = m_new;
            // exponentials in registers (packed fp16) while P.V(i-1) is still running
            uint32_t w[32];
            float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
                float pv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k0 = 2 * k + e;  // -inf - finite -> 0; a fully masked row gives NaN, handled below
                    pv[e] = fast_exp2(__uint_as_float(sr[k0 >> 5][k0 & 31]) - m_ref);
                }
                ps4[(k >> 1) & 3] += (pv[0] + pv[1]) + (pv[2] + pv[3]);
                if (KV == 1) {  // the V scale of each token folded into its P column (the row sum stays unscaled)
                    const uint4 f = lds_128(scale_sm + sfs * SCALE_BYTES + KT * 4 + k * 8);
                    pv[0] *= __uint_as_float(f.x);
                    pv[1] *= __uint_as_float(f.y);
                    pv[2] *= __uint_as_float(f.z);
                    pv[3] *= __uint_as_float(f.w);
                }
                w[k] = pack_half2(pv[0], pv[1]);
                w[k + 1] = pack_half2(pv[2], pv[3]);
            }
            if (qtr == 0) PROBE(x, i, 3);
            const bool dead = m_ref == -INFINITY;  // nothing visible yet (padding rows of the last query tile)
            if (qtr == 0) PROBE(x, i, 4);
            if (i >= 2) ensure_pv(i - 2);  // complete since S(i) was written: keeps the phase bookkeeping tight
            if (__any_sync(0xffffffffu, corr != 1.f)) {  // warp-collective TMEM access; lanes that keep their
                // reference multiply by 1.  P.V(i-1) must be complete before O is rescaled.
                if (i > 0) ensure_pv(i - 1);
                tc_fence_after();
#pragma unroll
                for (int c4 = 0; c4 < D / 32; ++c4) {
                    uint32_t orr[32];
                    tmem_ld32(o_addr + c4 * 32, orr);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * corr);
                    tmem_st32(o_addr + c4 * 32, orr);
                }
                l_run *= corr;
            }
            // P(i) as packed fp16 over the first 32 columns of this tile's S buffer (every score of the row is
            // already in registers); it is the TMEM A operand of P.V(i).  The buffer's previous user, P.V(i-2),
            // completed before S(i) was written (same issuing thread, in order).
            if (dead) {
#pragma unroll
                for (int k = 0; k < 32; ++k) w[k] = 0u;
            }
            tmem_st32(s_addr + sb * KT, w);
            tmem_wait_st();
            if (!dead) l_run += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
            if (qtr == 0) PROBE(x, i, 5);
            tc_fence_before();
            mbar_arrive_cnt(p_full(x, sb));
            if (qtr == 0) PROBE(x, i, 6);
            if (++sfs == ST) {
                sfs = 0;
                sfph ^= 1u;
            }
        }
        PROBE_CTA(4);
        if (nt > 0) {
            if (nt > 1) ensure_pv(nt - 2);
            ensure_pv(nt - 1);
            PROBE_CTA(5);
            tc_fence_after();
            const float inv = 1.f / (l_run + 1e-6f);  // softmax_lut.cpp:224 epsilon (App. A D4)
            // O rows leave through shared memory so that every global store instruction writes whole 128-byte
            // row segments (a thread-per-row store scatters 16 bytes to each of 32 rows).  Staging area: this
            // warp's quarter of its tile's Q buffer, idle since the tile's last S MMA (which completed before
            // the P.V just waited for); 32 rows x 128 B per round, 16-byte chunks XOR-swizzled by row.
            const uint32_t ebuf = q_sm + x * Q_BYTES + qtr * (Q_BYTES / 4);
            float* out_tile = a.out + (int64_t)b * a.sb + (int64_t)h * a.sh + (int64_t)(q0 + qtr * 32) * a.st;
            const int rows_ok = a.Tq - (q0 + qtr * 32);  // rows of this warp inside the prompt chunk
#pragma unroll 1
            for (int c4 = 0; c4 < D / 32; ++c4) {
                uint32_t orr[32];
                tmem_ld32(o_addr + c4 * 32, orr);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t addr = ebuf + lane * 128 + ((j ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr),
                                 "f"(__uint_as_float(orr[4 * j]) * inv), "f"(__uint_as_float(orr[4 * j + 1]) * inv),
                                 "f"(__uint_as_float(orr[4 * j + 2]) * inv), "f"(__uint_as_float(orr[4 * j + 3]) * inv)
                                 : "memory");
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + (lane >> 3), j = lane & 7;
                    const uint4 v = lds_128(ebuf + r * 128 + ((j ^ (r & 7)) << 4));
                    if (r < rows_ok) *reinterpret_cast<uint4*>(out_tile + (int64_t)r * a.st + c4 * 32 + j * 4) = v;
                }
                __syncwarp();
            }
        } else if (t_ok) {
            // no visible key at all (zero-capacity table): the reference's 0 / (0 + eps) = 0
            float* orow = a.out + (int64_t)b * a.sb + (int64_t)h * a.sh + (int64_t)t * a.st;
            for (int j = 0; j < D; j += 4) *reinterpret_cast<float4*>(orow + j) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        PROBE_CTA(6);
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    PROBE_CTA(7);
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace ptc
}  // namespace pa

using namespace pa;

#ifdef PA_PTC_PROBE
extern "C" __attribute__((visibility("default"))) int pa_debug_ptc_probe(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, pa::ptc::g_probe, sizeof(long long) * 3 * 64 * 8);
}
extern "C" __attribute__((visibility("default"))) int pa_debug_ptc_probe_cta(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, pa::ptc::g_probe_cta, sizeof(long long) * 8);
}
#endif

// Launch helper used by prefill.cu (returns PA_ERR_UNSUPPORTED when the tensor maps cannot be built).
int pa_prefill_tc_launch(int kv, int head_dim, const float* d_q, float* d_out, const void* d_k_pool, const void* d_v_pool,
                         const float* d_k_scales, const float* d_v_scales, const int32_t* d_table, int num_beams,
                         int num_heads, int num_tiles, int total_pages, const int32_t* d_beam_ids,
                         const int32_t* d_ctx_start, int B, int Tq, int tile_size, float temperature, int token_major,
                         cudaStream_t st) {
    using namespace pa::ptc;
    CUtensorMap tmK, tmV;
    const uint64_t total_tokens = (uint64_t)total_pages * tile_size;
    if (total_tokens >= 0x7fffffffull) return PA_ERR_UNSUPPORTED;
    if (kv == 0) {
        if (!make_pool_map(&tmK, d_k_pool, total_tokens, head_dim) || !make_pool_map(&tmV, d_v_pool, total_tokens, head_dim))
            return PA_ERR_UNSUPPORTED;
    } else {  // the int8 variant stages raw units with plain bulk copies
        memset(&tmK, 0, sizeof(tmK));
        memset(&tmV, 0, sizeof(tmV));
    }
    const int upt = tile_size >> 4;
    if (head_dim != 64 && head_dim != 128) return PA_ERR_UNSUPPORTED;
    if (tile_size % 16 != 0 || (upt & (upt - 1)) != 0) return PA_ERR_UNSUPPORTED;  // pages of 16 << k tokens only
    int upt_shift = 0;
    while ((1 << upt_shift) < upt) ++upt_shift;
    Args a{d_q, d_out, d_table, d_beam_ids, d_ctx_start, num_beams, num_heads, num_tiles, total_pages, B, Tq, tile_size,
           1.4426950408889634f / temperature, static_cast<const int8_t*>(d_k_pool), static_cast<const int8_t*>(d_v_pool),
           d_k_scales, d_v_scales, (getenv("PA_PTC_STAGGER") && atoi(getenv("PA_PTC_STAGGER")) == 0) ? 0 : 1, upt_shift,
           (int)total_tokens,
           token_major ? (int64_t)Tq * num_heads * head_dim : (int64_t)num_heads * Tq * head_dim,
           token_major ? (int64_t)head_dim : (int64_t)Tq * head_dim,
           token_major ? (int64_t)num_heads * head_dim : (int64_t)head_dim};
    // NQ = 2 (256 queries per CTA share every staged K/V byte) unless that leaves SMs without a CTA: small
    // prompts then take NQ = 1 (twice the CTAs).  PA_PREFILL_NQ forces either.
    int dev = 0;
    cudaGetDevice(&dev);
    static int sm_count[64] = {};
    if (!sm_count[dev & 63]) cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    int nq = ((int64_t)B * num_heads * ((Tq + 2 * QT - 1) / (2 * QT)) < sm_count[dev & 63]) ? 1 : 2;
    if (getenv("PA_PREFILL_NQ")) nq = atoi(getenv("PA_PREFILL_NQ")) == 1 ? 1 : 2;
    const int nqt = (Tq + nq * QT - 1) / (nq * QT);
    const int64_t ctas = (int64_t)B * num_heads * nqt;
    if (ctas > 0x7fffffff) return PA_ERR_INVALID_ARG;
    const size_t smem = (size_t)nq * q_bytes(head_dim) + stages_for(nq) * stage_bytes(head_dim) +
                        (kv ? RS * 4 * raw_unit(head_dim) + stages_for(nq) * SCALE_BYTES : 0) +
                        (nbar_for(nq) + 2 * RS) * 8 + 8 + RS * 4 + 16 + 1024;
    // one UMMA issuer per query tile for fp16 pages (PA_PREFILL_MW=1: single issuer, =2: two issuers)
    // (int8 pages: 440 / 585 TFLOP/s with two issuers against 453 / 606 with one -- their limit is elsewhere -- so they
    //  stay on MW = 1 unless PA_PREFILL_MW=2)
    const char* mw_env = getenv("PA_PREFILL_MW");
    const bool mw2 = nq == 2 && (mw_env ? atoi(mw_env) == 2 : kv == 0);
    static bool attr_done[64][12] = {};
    const int ki = mw2 ? 8 + kv * 2 + (head_dim == 64 ? 1 : 0) : (head_dim == 64 ? 4 : 0) + kv * 2 + (nq == 1 ? 1 : 0);
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const Args);
    static const KernelFn kernels[12] = {prefill_tc_kernel<0, 2, 128>,    prefill_tc_kernel<0, 1, 128>, prefill_tc_kernel<1, 2, 128>,
                                         prefill_tc_kernel<1, 1, 128>,    prefill_tc_kernel<0, 2, 64>,  prefill_tc_kernel<0, 1, 64>,
                                         prefill_tc_kernel<1, 2, 64>,     prefill_tc_kernel<1, 1, 64>,  prefill_tc_kernel<0, 2, 128, 2>,
                                         prefill_tc_kernel<0, 2, 64, 2>,  prefill_tc_kernel<1, 2, 128, 2>, prefill_tc_kernel<1, 2, 64, 2>};
    KernelFn kern = kernels[ki];
    if (!attr_done[dev & 63][ki]) {
        cudaError_t e0 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e0 != cudaSuccess) return (int)e0;
        attr_done[dev & 63][ki] = true;
    }
    kern<<<(unsigned)ctas, threads_for(kv, nq, mw2 ? 2 : 1), smem, st>>>(tmK, tmV, a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PA_OK : (int)e;
}
