// linear_tf32x3.cu -- the float MLP / projection layer of the fp32 decoder (decoder/mlp.hpp:23-41:
// out[r,n] = act(bias[n] + sum_k x[r,k] * W[k*N+n])) on the 5th-gen tensor cores, at fp32-class accuracy.
//
// tcgen05.mma kind::tf32 reads 32-bit containers and uses sign, 8 exponent and the top 10 mantissa bits.  Each
// operand is therefore split exactly into  v = hi + lo,  hi = v with the low 13 mantissa bits cleared (what the
// tensor core sees of v itself), lo = v - hi (exact in fp32, 13 significant bits of which the core keeps 11), and
//     x . W  ~=  x_hi . W_lo + x_lo . W_hi + x_hi . W_hi                     ("3xTF32")
// accumulated in fp32 in TMEM.  Dropped: lo.lo (2^-22 relative) and the tail of each lo (2^-21); the tensor core's
// fp32 accumulation truncates, which adds ~K/8 * 2^-24 of the running sum (the lo terms have accumulators of their
// own, so they are not truncated against the big sum).  Measured <= 1.1e-6 of sum_k |x||W| (K = 11008, unsliced;
// 2.5e-7 at decode batches; the fp32 SIMT kernels: 1-3e-7); the tests state 1e-5 * sum_k |x||W| + 4 ulp.
// The fp32 SIMT kernels (PA_LINEAR_TC=0) stay for callers that need fp32 arithmetic proper.
//
// The layer is computed TRANSPOSED: out^T [N, rows] = W^T [N, K] . x^T [K, rows].  The weights are the M side of the
// MMA (128 output features = 128 TMEM lanes), the decode batch the N side, so a small batch costs a small MMA
// (tcgen05 time is max(M,128) * N / 256 cycles: 32 for 64 rows, where batch-as-M would pay 64 for anything <= 128)
// and the accumulator tile [128 features][rows] is what the epilogue wants: lane = feature, so every store
// instruction writes 32 consecutive features of one row (128 bytes).  W is consumed in the reference's [K, N]
// layout and x where it lies; the weights stream from HBM exactly once per 128 rows of x (L2 serves the other row tiles).
//
// One CTA = 128 features x NP rows (NP = 64 up to 64 rows, else 128) over one K slice; 320 threads:
//   warp 0    TMA producer, per 32-float K block: ONE W box [32 k][128 n] (512-byte rows, no swizzle: only threads
//             read it; or one 16 KB bulk copy of a packed block) and the x tile [NP rows][128 B] (128-byte swizzle, K-major B operand; rows / k out of range zero-filled);
//   warps 2-9 split: thread = (feature n, half of the block's k): reads its 16 weights W[k][n] down the column
//             (conflict-free: a warp reads 32 consecutive floats of one k-row), stores them and their lo parts into
//             TENSOR MEMORY as the A operands (lane n, 32 + 32 columns per stage) -- the weights never touch shared
//             memory again; x gets its lo tile beside it (same swizzled offsets, so the split never needs the
//             layout); fence to the async proxy, arrive.  After the main loop the same warps are the epilogue;
//   warp 1    one elected lane issues 12 tcgen05.mma (4 k-steps of 8 x 3 terms; A from TMEM, B from shared memory)
//             per stage, commit frees the stage.
// Shared-memory traffic per stage at 64 rows: 16 KB W written + 16 KB read, x 8 + 8 (lo) written, 8 read, 24 KB of
// B-operand reads = 80 KB per 16 KB of weights, against 128 KB for the batch-as-M form measured first (r02_notes.md).
// Work split: a K-sliced grid (row tile, feature tile, K slice) whose fp32 partial tiles linear_reduce4_kernel sums in
// slice order, or -- where that grid would not fit one wave -- the stream-K form (equal contiguous ranges of the
// (tile, K block) list, one CTA per SM, partial tiles summed in range order by linear_streamk_reduce_kernel).  Both sum
// kernels are chained by programmatic dependent launch.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "pa_common.cuh"

namespace pa {
namespace tf32x3 {

constexpr int BF = 128;            // output features per tile = UMMA M = TMEM lanes
constexpr int BKF = 32;            // floats of K per stage (one 128-byte swizzle row of x)
constexpr int W_TILE = BKF * BF * 4;   // 16 KiB: W block [32 k][128 n]
constexpr int NTHREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 split + epilogue (two per TMEM lane quarter)
// per row-tile width NP: stage = W (16 KB) + x + x_lo (NP * 128 B each); TMEM = NP accumulator columns + 64 per stage
template <int NP> struct Cfg {
    static constexpr int X_TILE = NP * 128;
    static constexpr int STAGE = W_TILE + 2 * X_TILE;
    // Accumulators: the tensor core's fp32 accumulation truncates, so the small terms of the split get accumulators of
    // their own where TMEM allows (summed in the epilogue, small terms first): the lo terms are then not truncated against
    // the big sum -- measured error 6.6e-7 -> 2.5e-7 of sum |x||W| at decode batches, speed-neutral.  3 at NP = 64,
    // 2 (lo terms / hi.hi) at NP = 128.
    static constexpr int NACC = NP == 64 ? 3 : 2;
    static constexpr int NST = NP == 64 ? 5 : 4;
    static constexpr int TMEM_COLS = 512;
    static_assert(NACC * NP + NST * 64 <= TMEM_COLS, "TMEM budget");
    static_assert(NST * STAGE <= 200 * 1024, "shared-memory budget");
};

struct Args {
    const float* bias;
    float* out;
    float* partial;   // [nslices][rows][N] when nslices > 1
    int rows, N, K, act, kslice, nslices;
    const float* w_packed;   // PACKED kernels: [feature tile][K block][32 k][128 n] (pa_linear_pack_f32), else null
    // STREAMK kernels: the (tile, K block) work list is cut into equal contiguous ranges, one per CTA
    int sk_per;              // K blocks per CTA
    int sk_maxseg;           // partial-tile slots per CTA
    int n_ft, n_mt;          // feature tiles, row tiles (tile index = ft * n_mt + mt: CTAs that run side by side share W)
    int64_t sk_total;        // tiles * K blocks
    unsigned long long* probe;   // bring-up aid (pa_debug_linear_probe): per-CTA cycle sums, null in production
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in TMEM: row m of A = TMEM lane m, 32-bit column j = A[m][j]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
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
// UMMA shared-memory descriptor, K-major, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14),
// LBO >> 4 [16,30) (unused: one swizzle row of K), SBO >> 4 [32,46) = 1024 (8-row groups), version 1 [46,48),
// SWIZZLE_128B = 2 [61,64).  x tile: rows of 128 B (32 floats of K).
// (Bring-up note: W as an MN-major tf32 B operand works only with layout 1, SWIZZLE_128B_BASE32B, fed by
//  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; with the plain 128-byte swizzle the tensor core returns zeros.  That form
//  was measured and replaced by this one, see profiles/r02_notes.md.)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// Instruction descriptor: c = F32 (1) [4,6), a = b = TF32 (2) [7,10) [10,13), a (TMEM) and b K-major (0) [15] [16],
// N >> 3 [17,23), M >> 4 [24,29).
__host__ __device__ constexpr uint32_t idesc_for(int np) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(np >> 3) << 17) | ((uint32_t)(BF >> 4) << 24);
}

// v - (v with the low 13 mantissa bits cleared): exact in fp32.  (A non-finite v gives lo = NaN, so an inf in x or W
// turns its whole output row / column into NaN where the scalar loop would give +-inf or NaN per element.)
__device__ __forceinline__ float lo_part(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    // a lost arrival would otherwise hang the GPU box: ~2 s of polling, then trap
    for (uint32_t i = 0; i < (1u << 26); ++i)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// Work of one CTA = a list of SEGMENTS (tile, K-block range).  Plain grid: one segment, from blockIdx (row tile,
// feature tile, K slice).  STREAMK: the list of all (tile, K block) pairs, tile-major, is cut into equal contiguous
// ranges, one per CTA (one CTA per SM, every CTA the same number of blocks -- no wave quantisation, one prologue per
// SM); a range covers the tail of one tile, whole tiles, and the head of another.  A segment that is a whole tile
// is finished in place (bias / relu -> out); the others leave a partial tile in the CTA's own slots, summed in
// range order by linear_streamk_reduce_kernel.  The shared-memory and TMEM rings run on across segment boundaries.
struct Segment {
    int m0, n0, ft, kb0, nkb, whole, slot;
};

template <int NP, bool STREAMK>
struct SegmentIter {
    const Args& g;
    int tk;
    int64_t w, end;
    int sg;
    __device__ SegmentIter(const Args& g_) : g(g_), tk((g_.K + BKF - 1) / BKF), sg(0) {
        if (STREAMK) {
            w = (int64_t)blockIdx.x * g.sk_per;
            end = w + g.sk_per < g.sk_total ? w + g.sk_per : g.sk_total;
        } else {
            w = 0;
            end = 1;
        }
    }
    __device__ bool next(Segment& sgm) {
        if (STREAMK) {
            if (w >= end) return false;
            const int tile = (int)(w / tk);
            sgm.kb0 = (int)(w - (int64_t)tile * tk);
            const int64_t left = end - w;
            sgm.nkb = left < tk - sgm.kb0 ? (int)left : tk - sgm.kb0;
            sgm.ft = tile / g.n_mt;
            sgm.n0 = sgm.ft * BF;
            sgm.m0 = (tile % g.n_mt) * NP;
            sgm.whole = sgm.kb0 == 0 && sgm.nkb == tk;
            sgm.slot = blockIdx.x * g.sk_maxseg + sg;
            w += sgm.nkb;
            ++sg;
            return true;
        }
        if (w >= end) return false;
        w = end;
        const int k_begin = blockIdx.z * g.kslice;               // kslice is a multiple of BKF
        const int k_end = min(g.K, k_begin + g.kslice);
        sgm.kb0 = k_begin / BKF;
        sgm.nkb = (k_end - k_begin + BKF - 1) / BKF;
        sgm.ft = blockIdx.y;
        sgm.n0 = blockIdx.y * BF;
        sgm.m0 = blockIdx.x * NP;
        sgm.whole = g.nslices == 1;
        sgm.slot = blockIdx.z;
        return true;
    }
};

template <int NP, bool PACKED, bool STREAMK>
__global__ void __launch_bounds__(NTHREADS, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Args g) {
    using C = Cfg<NP>;
    constexpr int NST = C::NST, STAGE = C::STAGE, X_TILE = C::X_TILE, NACC = C::NACC;
    constexpr int RING0 = NACC * NP;   // first TMEM column of the weight ring
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + NST * STAGE;
    auto full_bar = [&](int s) { return bar0 + s * 8; };
    auto split_bar = [&](int s) { return bar0 + (NST + s) * 8; };
    auto empty_bar = [&](int s) { return bar0 + (2 * NST + s) * 8; };
    const uint32_t tmem_full_bar = bar0 + 3 * NST * 8;   // MMAs of a segment complete -> epilogue
    const uint32_t acc_free_bar = tmem_full_bar + 8;      // epilogue has read the accumulators -> next segment's MMAs
    const uint32_t tmem_slot = acc_free_bar + 8;

    asm volatile("griddepcontrol.launch_dependents;");   // the partial-tile sum behind this grid may become resident
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tk = (g.K + BKF - 1) / BKF;
    const long long t_cta0 = g.probe ? clock64() : 0;
    const int cta_id = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(split_bar(s), 256);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(acc_free_bar, 256);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW));
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // stage layout: W [32 k][128 n] | x [NP][128 B] | x_lo;  TMEM: NACC accumulators of NP columns, then per stage W (32)
    // and W_lo (32)

    if (warp == 0) {
        if (elect_one()) {
            const uint64_t stream_policy = l2_policy_evict_first();
            SegmentIter<NP, STREAMK> segs(g);
            Segment sg;
            int it = 0;
            long long pr_wait = 0;
            // The kernel is launched with programmatic stream serialisation: it may start while the kernel before it
            // in the stream (the producer of x: a LayerNorm, the previous layer's sum kernel) is still running.  The
            // WEIGHTS do not depend on that kernel, so the first ring-full of W blocks is requested at once; x -- and
            // everything downstream of it: the split, the MMAs, every store -- waits for griddepcontrol.wait.
            int x_it = 0;           // blocks whose x tile has been requested
            bool dep_ok = false;    // griddepcontrol.wait executed
            auto load_x_upto = [&](int upto) {   // request the x tiles of blocks [x_it, upto) of the work list, in order
                SegmentIter<NP, STREAMK> xs(g);
                Segment xsg;
                int j = 0;
                while (xs.next(xsg) && j < upto) {
                    for (int i = 0; i < xsg.nkb && j < upto; ++i, ++j)
                        if (j >= x_it)
                            tma_load_2d(base + (j % NST) * STAGE + W_TILE, &tmX, (xsg.kb0 + i) * BKF, xsg.m0, full_bar(j % NST));
                }
                x_it = upto > x_it ? upto : x_it;
            };
            while (segs.next(sg)) {
                for (int i = 0; i < sg.nkb; ++i, ++it) {
                    const int s = it % NST;
                    if (it >= NST && !dep_ok) {   // the ring is full of weights: now x is needed
                        asm volatile("griddepcontrol.wait;" ::: "memory");
                        dep_ok = true;
                        load_x_upto(it);
                    }
                    const long long tp0 = g.probe ? clock64() : 0;
                    mbar_wait_bounded(empty_bar(s), ((it / NST) & 1) ^ 1);
                    if (g.probe) pr_wait += clock64() - tp0;
                    const uint32_t st = base + s * STAGE;
                    mbar_arrive_expect_tx(full_bar(s), W_TILE + X_TILE);
                    const int kb = sg.kb0 + i;
                    if (PACKED) {   // the block is ONE contiguous 16 KB run (whole DRAM pages), streamed past L2
                        const int64_t blk = (int64_t)sg.ft * tk + kb;
                        bulk_g2s(st, g.w_packed + blk * (W_TILE / 4), W_TILE, full_bar(s), stream_policy);
                    } else {
                        tma_load_2d(st, &tmW, sg.n0, kb * BKF, full_bar(s));
                    }
                    if (dep_ok) {
                        tma_load_2d(st + W_TILE, &tmX, kb * BKF, sg.m0, full_bar(s));
                        x_it = it + 1;
                    }
                }
            }
            if (!dep_ok) {   // fewer blocks than ring stages
                asm volatile("griddepcontrol.wait;" ::: "memory");
                load_x_upto(it);
            }
            if (g.probe) g.probe[cta_id * 8 + 0] = pr_wait;
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_for(NP);
            SegmentIter<NP, STREAMK> segs(g);
            Segment sg;
            int it = 0, nseg = 0;
            long long mm_wait = 0, mm_issue = 0;
            while (segs.next(sg)) {
                if (nseg > 0) {   // the epilogue of the previous segment has drained the accumulators
                    mbar_wait_bounded(acc_free_bar, (uint32_t)((nseg - 1) & 1));
                    tc_fence_after();
                }
                for (int i = 0; i < sg.nkb; ++i, ++it) {
                    const int s = it % NST;
                    const long long tm0 = g.probe ? clock64() : 0;
                    mbar_wait_bounded(split_bar(s), (it / NST) & 1);
                    const long long tm1 = g.probe ? clock64() : 0;
                    tc_fence_after();
                    const uint32_t st = base + s * STAGE;
#pragma unroll
                    for (int ks = 0; ks < BKF / 8; ++ks) {
                        const uint32_t wa = tmem_base + RING0 + s * 64 + ks * 8;
                        const uint32_t wl = wa + 32;
                        const uint64_t xa = make_desc(st + W_TILE + ks * 32);
                        const uint64_t xl = make_desc(st + W_TILE + X_TILE + ks * 32);
                        const uint32_t first = (i > 0 || ks > 0) ? 1u : 0u;
                        // accumulator 0: hi.hi; 1: lo terms (W_lo.x first); 2: x_lo.W_hi where there are three
                        umma_tf32_ts(tmem_base + (NACC > 1 ? NP : 0), wl, xa, idesc, first);
                        umma_tf32_ts(tmem_base + (NACC > 2 ? 2 * NP : (NACC > 1 ? NP : 0)), wa, xl, idesc, NACC > 2 ? first : 1u);
                        umma_tf32_ts(tmem_base, wa, xa, idesc, NACC > 1 ? first : 1u);
                    }
                    umma_commit(empty_bar(s));
                    if (g.probe) {
                        mm_wait += tm1 - tm0;
                        mm_issue += clock64() - tm1;
                    }
                }
                umma_commit(tmem_full_bar);
                ++nseg;
            }
            if (g.probe) {
                g.probe[cta_id * 8 + 1] = mm_wait;
                g.probe[cta_id * 8 + 2] = mm_issue;
            }
        }
    } else {
        const int t = threadIdx.x - 64;      // 0..255
        const int qtr = warp & 3;            // TMEM lane quarter this warp may touch
        const int half = (warp - 2) >> 2;    // which half of a block's k (split) / of the row tile (epilogue)
        const int f = qtr * 32 + lane;       // feature of the tile = TMEM lane
        SegmentIter<NP, STREAMK> segs(g);
        Segment sg;
        int it = 0, nseg = 0;
        long long cv_wait = 0, cv_work = 0, cv_blocks = 0, cv_epi = 0;
        while (segs.next(sg)) {
            for (int i = 0; i < sg.nkb; ++i, ++it) {
                const int s = it % NST;
                const long long tc0 = (g.probe && t == 0) ? clock64() : 0;
                mbar_wait_bounded(full_bar(s), (it / NST) & 1);
                const long long tc1 = (g.probe && t == 0) ? clock64() : 0;
                const uint32_t st = base + s * STAGE;
                // W: column f of the block, k rows [half * 16, half * 16 + 16) -> 16 columns of W and of W_lo in TMEM
                {
                    uint32_t wv[16], wlo[16];
                    const uint32_t col = st + f * 4 + half * 16 * 512;
#pragma unroll
                    for (int j = 0; j < 16; ++j) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv[j]) : "r"(col + j * 512));
#pragma unroll
                    for (int j = 0; j < 16; ++j) wlo[j] = __float_as_uint(lo_part(__uint_as_float(wv[j])));
                    const uint32_t ta = tmem_base + ((uint32_t)(qtr * 32) << 16) + RING0 + s * 64 + half * 16;
                    tmem_st16(ta, wv);
                    tmem_st16(ta + 32, wlo);
                }
                // x tile -> x_lo: NP * 8 chunks of 16 bytes
#pragma unroll
                for (int j = 0; j < NP / 32; ++j) {
                    const uint32_t off = (j * 256 + t) * 16;
                    const uint4 v = lds_128(st + W_TILE + off);
                    const float a = lo_part(__uint_as_float(v.x)), b = lo_part(__uint_as_float(v.y));
                    const float c = lo_part(__uint_as_float(v.z)), d = lo_part(__uint_as_float(v.w));
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(st + W_TILE + X_TILE + off), "f"(a), "f"(b),
                                 "f"(c), "f"(d)
                                 : "memory");
                }
                fence_proxy_async();
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(split_bar(s));
                if (g.probe && t == 0) {
                    cv_wait += tc1 - tc0;
                    cv_work += clock64() - tc1;
                    cv_blocks += 1;
                }
            }
            const long long te0 = (g.probe && t == 0) ? clock64() : 0;
            // ---- epilogue of the segment: lane = feature, columns = rows of x; this warp takes half of the row tile ----
            const int n = sg.n0 + f;
            const bool n_ok = n < g.N;
            mbar_wait_bounded(tmem_full_bar, (uint32_t)(nseg & 1));
            tc_fence_after();
            const float bias = (sg.whole && g.bias && n_ok) ? __ldg(g.bias + n) : 0.f;
            // whole tile: out [rows][N]; K slice of the plain grid: partial [slice][rows][N]; stream-K: this CTA's slot
            // [NP rows][128 features]
            float* dst;
            int64_t ld;
            if (sg.whole) {
                dst = g.out + (int64_t)sg.m0 * g.N + n;
                ld = g.N;
            } else if (STREAMK) {
                dst = g.partial + (int64_t)sg.slot * NP * BF + f;
                ld = BF;
            } else {
                dst = g.partial + ((int64_t)sg.slot * g.rows + sg.m0) * g.N + n;
                ld = g.N;
            }
            const bool store_ok = n_ok || (STREAMK && !sg.whole);   // pad features of a slot are written (zeros), never read
            constexpr int CH = NP / 64;   // 32-row chunks per warp
#pragma unroll 1
            for (int c = half * CH; c < half * CH + CH; ++c) {
                uint32_t rr[32];
                const uint32_t ta = tmem_base + ((uint32_t)(qtr * 32) << 16) + c * 32;
                if (NACC == 1) {
                    tmem_ld_32x32(ta, rr);
                } else {
                    uint32_t r1[32];
                    tmem_ld_32x32(ta + NP, r1);          // lo terms (W_lo.x)
                    if (NACC > 2) {
                        tmem_ld_32x32(ta + 2 * NP, rr);  // x_lo.W_hi
#pragma unroll
                        for (int j = 0; j < 32; ++j) r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __uint_as_float(rr[j]));
                    }
                    tmem_ld_32x32(ta, rr);               // hi.hi
#pragma unroll
                    for (int j = 0; j < 32; ++j) rr[j] = __float_as_uint(__uint_as_float(rr[j]) + __uint_as_float(r1[j]));
                }
                if (!store_ok) continue;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int rl = c * 32 + j;
                    if (sg.m0 + rl >= g.rows && !(STREAMK && !sg.whole)) break;
                    float v = __uint_as_float(rr[j]);
                    if (sg.whole) {
                        v += bias;
                        if (g.act == PA_ACT_RELU) v = fmaxf(v, 0.f);
                    }
                    dst[(int64_t)rl * ld] = v;
                }
            }
            tc_fence_before();
            mbar_arrive(acc_free_bar);
            if (g.probe && t == 0) cv_epi += clock64() - te0;
            ++nseg;
        }
        if (g.probe && t == 0) {
            g.probe[cta_id * 8 + 3] = cv_wait;
            g.probe[cta_id * 8 + 4] = cv_work;
            g.probe[cta_id * 8 + 6] = cv_epi;
            g.probe[cta_id * 8 + 7] = cv_blocks;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (g.probe && threadIdx.x == 0) g.probe[cta_id * 8 + 5] = clock64() - t_cta0;
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// Sum of the partial tiles of the stream-K form: tile t was computed by the CTAs c0 .. c1 whose ranges intersect its K
// blocks (c0 == c1: the tile was finished in place, nothing to do); they are added in range order -- the result does
// not depend on the schedule.  Launched with programmatic dependent launch behind the main kernel.
template <int NP>
__global__ void __launch_bounds__(256) linear_streamk_reduce_kernel(const float* __restrict__ partial,
                                                                    const float* __restrict__ bias, int rows, int N, int tk,
                                                                    int per, int maxseg, int n_mt, int act,
                                                                    float* __restrict__ out) {
    asm volatile("griddepcontrol.launch_dependents;");   // the next layer's kernel may start streaming its weights
    // grid: x = 16-byte column groups of a row, y = rows (strided); 32-bit index arithmetic (tiles * tk < 2^31, host-checked)
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool n_ok = n < N;
    const int ft = n / BF;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_ok && bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!n_ok) return;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        const int tile = ft * n_mt + row / NP;
        const unsigned w0 = (unsigned)tile * (unsigned)tk;
        const int c0 = (int)(w0 / (unsigned)per), c1 = (int)((w0 + (unsigned)tk - 1u) / (unsigned)per);
        if (c0 == c1) continue;   // the tile was finished in place by its one CTA
        float4 s = b4;
        for (int c = c0; c <= c1; ++c) {
            const int first_tile = (int)(((unsigned)c * (unsigned)per) / (unsigned)tk);
            const int64_t slot = (int64_t)c * maxseg + (tile - first_tile);
            const float4 p = __ldcg(reinterpret_cast<const float4*>(partial + (slot * NP + row % NP) * BF + n % BF));
            s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
        if (act == PA_ACT_RELU) {
            s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f);
        }
        *reinterpret_cast<float4*>(out + (int64_t)row * N + n) = s;
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
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
// 2-D f32 tensor [rows][cols] (cols contiguous), box [box_rows][box_cols], zero fill out of range.
static bool make_map_f32(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_cols,
                         uint32_t box_rows, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Row tile: 64 up to 64 rows, else 128.  (A 256-row tile has room for ONE accumulator only: measured slower -- prefill
// fc1 824 vs 711 us -- and 3x less accurate than two 128-row tiles with the lo terms accumulated apart.)
static int row_tile(int rows) { return rows <= 64 ? 64 : 128; }

template <int NP, bool PACKED, bool STREAMK>
static int launch(const CUtensorMap& tmX, const CUtensorMap& tmW, const Args& g, dim3 grid, cudaStream_t st) {
    using C = Cfg<NP>;
    const size_t smem = (size_t)C::NST * C::STAGE + (3 * C::NST + 2) * 8 + 16 + 1024;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_kernel<NP, PACKED, STREAMK>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, linear_tf32x3_kernel<NP, PACKED, STREAMK>, tmX, tmW, g);
    return e == cudaSuccess ? PA_OK : (int)e;
}

template <typename... A>
static int launch_pdl(void (*kern)(A...), dim3 blocks, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = blocks;
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    return e == cudaSuccess ? PA_OK : (int)e;
}

// K-slice sum behind the plain grid: 16-byte lanes (N % 4 == 0), programmatic dependent launch (its blocks are resident
// when the last partial tile lands; griddepcontrol.wait = all of the primary grid's memory is visible).  Slices are
// added in index order: the result does not depend on the schedule.
__global__ void __launch_bounds__(256) linear_reduce4_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                                             int rows, int N, int nslices, int act, float* __restrict__ out) {
    asm volatile("griddepcontrol.launch_dependents;");   // the next layer's kernel may start streaming its weights
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t total = (int64_t)rows * N;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total && bias) s = __ldg(reinterpret_cast<const float4*>(bias + i % N));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (i >= total) return;
    for (int z = 0; z < nslices; ++z) {
        const float4 p = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)z * total + i));
        s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    if (act == PA_ACT_RELU) {
        s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f);
    }
    *reinterpret_cast<float4*>(out + i) = s;
}

// Stream-K geometry: equal contiguous ranges of the (tile, K block) list, one CTA per SM (at least 8 blocks each).
struct StreamK {
    int grid, per, maxseg, n_ft, tk;
    int64_t total;
    size_t ws_bytes;
};
static StreamK streamk_plan(int rows, int K, int N, int sm_count) {
    StreamK p;
    const int np = row_tile(rows);
    p.n_ft = (N + BF - 1) / BF;
    p.tk = (K + BKF - 1) / BKF;
    p.total = (int64_t)((rows + np - 1) / np) * p.n_ft * p.tk;
    int64_t gmax = p.total / 8;
    if (gmax < 1) gmax = 1;
    const int64_t G = gmax < sm_count ? gmax : sm_count;
    p.per = (int)((p.total + G - 1) / G);
    p.grid = (int)((p.total + p.per - 1) / p.per);
    p.maxseg = (p.per + p.tk - 1) / p.tk + 1;
    p.ws_bytes = (size_t)p.grid * p.maxseg * np * BF * sizeof(float);
    return p;
}

// W [K, N] -> [feature tile][K block][32 k][128 n], zero padded: every (tile, block) the kernel streams is one
// contiguous 16 KB run.
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ W, float* __restrict__ Wp, int K, int N,
                                                           int64_t total4) {
    const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const int64_t e = i4 * 4;
    const int n = (int)(e % BF), k = (int)((e / BF) % BKF);
    const int64_t blk = e / (BF * BKF);
    const int kbs = (K + BKF - 1) / BKF;
    const int kb = (int)(blk % kbs), tile = (int)(blk / kbs);
    const int kk = kb * BKF + k, nn = tile * BF + n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kk < K) {
        const float* src = W + (int64_t)kk * N + nn;
        if (nn + 3 < N && (N & 3) == 0) {
            v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
            if (nn < N) v.x = src[0];
            if (nn + 1 < N) v.y = src[1];
            if (nn + 2 < N) v.z = src[2];
            if (nn + 3 < N) v.w = src[3];
        }
    }
    *reinterpret_cast<float4*>(Wp + e) = v;
}

}  // namespace tf32x3
}  // namespace pa

using namespace pa;

// K-slice geometry of the plain grid (the fallback when the stream-K form cannot be used: no workspace for its slots).
// One CTA per SM is resident, so the grid runs in waves of sm_count CTAs: pick the slice count that minimises
// waves x (K blocks per CTA + per-CTA set-up, ~2 blocks, + when sliced the partial tile's write and re-read, NP / 16
// blocks' worth of bytes); every slice holds >= 8 K blocks (256 k-rows).
// (Summing the slices inside the kernel -- last slice of a tile to arrive, one counter per tile -- was measured and
//  rejected: fence + counter + the last CTA's serial sum lengthen every wave, fc1 at batch 64 60 -> 88 us.)
static int tc_slices(int rows, int K, int N, int sm_count, int* kslice_out) {
    using namespace pa::tf32x3;
    const int np = row_tile(rows);
    const int64_t tiles = (int64_t)((rows + np - 1) / np) * ((N + BF - 1) / BF);
    const int total_kb = (K + BKF - 1) / BKF;
    int max_ns = total_kb / 8;
    if (max_ns > 16) max_ns = 16;
    if (max_ns < 1) max_ns = 1;
    int best_ns = 1;
    double best_cost = 1e30;
    for (int ns = 1; ns <= max_ns; ++ns) {
        const int64_t waves = (tiles * ns + sm_count - 1) / sm_count;
        const double cost = (double)waves * ((total_kb + ns - 1) / ns + 2 + (ns > 1 ? np / 16 : 0));
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best_ns = ns;
        }
    }
    int kslice = ((total_kb + best_ns - 1) / best_ns) * BKF;
    if (kslice_out) *kslice_out = kslice;
    return (K + kslice - 1) / kslice;
}

// Scratch for either form (pa_linear_workspace_bytes adds it to the SIMT kernels' need).
size_t pa_linear_tc_workspace_bytes(int rows, int K, int N, int sm_count) {
    using namespace pa::tf32x3;
    const int ns = tc_slices(rows, K, N, sm_count, nullptr);
    const size_t sliced = ns > 1 ? (size_t)ns * rows * N * sizeof(float) : 0;
    const size_t sk = streamk_plan(rows, K, N, sm_count).ws_bytes;
    return sliced > sk ? sliced : sk;
}

static unsigned long long* g_linear_probe = nullptr;
extern "C" __attribute__((visibility("default"))) void pa_debug_linear_probe(unsigned long long* d_counters) {
    g_linear_probe = d_counters;   // [CTAs][8] device counters, zeroed by the caller; null switches the probe off
}

// pa_linear_f32 / pa_linear_f32_packed on the tensor cores.  d_W: [K, N] (packed == 0) or the pa_linear_pack_f32
// layout (packed == 1).  Form: stream-K when the workspace holds its slots and the sum kernel's 16-byte lanes apply
// (N % 4 == 0, aligned out / bias; PA_LINEAR_STREAMK=0 disables), else the plain grid, K-sliced if the workspace
// allows.  PA_ERR_UNSUPPORTED when the tensor maps cannot be built.
int pa_linear_tc_run(const float* d_x, const float* d_W, const float* d_bias, int rows, int K, int N, int act,
                     float* d_out, void* d_ws, size_t ws_bytes, int packed, int sm_count, cudaStream_t st) {
    using namespace pa::tf32x3;
    const int np = row_tile(rows);
    CUtensorMap tmX, tmW;
    if (!make_map_f32(&tmX, d_x, (uint64_t)K, (uint64_t)rows, 32, (uint32_t)np, CU_TENSOR_MAP_SWIZZLE_128B))
        return PA_ERR_UNSUPPORTED;
    if (packed) memset(&tmW, 0, sizeof(tmW));
    else if (!make_map_f32(&tmW, d_W, (uint64_t)N, (uint64_t)K, BF, BKF, CU_TENSOR_MAP_SWIZZLE_NONE))
        return PA_ERR_UNSUPPORTED;
    const bool lanes16 = N % 4 == 0 && (uintptr_t)d_out % 16 == 0 && (uintptr_t)d_bias % 16 == 0 && d_ws &&
                         (uintptr_t)d_ws % 16 == 0;
    Args g{};
    g.bias = d_bias; g.out = d_out; g.partial = static_cast<float*>(d_ws);
    g.rows = rows; g.N = N; g.K = K; g.act = act; g.kslice = K; g.nslices = 1;
    g.w_packed = packed ? d_W : nullptr;
    g.probe = g_linear_probe;
    const unsigned n_mt = (unsigned)((rows + np - 1) / np), n_ft = (unsigned)((N + BF - 1) / BF);
    g.n_ft = (int)n_ft;
    g.n_mt = (int)n_mt;
    const StreamK sk = streamk_plan(rows, K, N, sm_count);
    int kslice = K;
    int nslices = tc_slices(rows, K, N, sm_count, &kslice);
    // Stream-K where the plain grid would need more than one wave (measured, L2 flushed: fc1 at batch 64, 430 CTAs in 2.9
    // waves, 58.3 -> 54.2 us; 256 rows 123.9 -> 111.6); a grid that fits ONE wave already has one prologue per SM, and the
    // accumulator hand-over at a range's tile boundary then only costs (fc2 54.1 -> 56.3, projection 27.6 -> 31.7).
    // PA_LINEAR_STREAMK=0 / 1 forces either.
    const char* env = getenv("PA_LINEAR_STREAMK");
    // ... and only for K-sliced grids of few tiles: with many whole tiles per CTA (prefill: 688 tiles) the plain grid's
    // block order keeps neighbouring CTAs on the same weights and wins (2048 rows: 714 vs 859 us).
    const bool multi_wave = (int64_t)n_mt * n_ft * nslices > sm_count && (int64_t)n_mt * n_ft <= 2 * (int64_t)sm_count;
    const bool use_sk = (env ? atoi(env) != 0 : multi_wave) && lanes16 && sk.tk >= 8 && ws_bytes >= sk.ws_bytes &&
                        sk.total + sk.tk < 0x7fffffffll;
    int rc;
    if (use_sk) {
        g.sk_per = sk.per; g.sk_maxseg = sk.maxseg; g.sk_total = sk.total;
        const dim3 grid((unsigned)sk.grid);
        if (packed) rc = np == 64 ? launch<64, true, true>(tmX, tmW, g, grid, st) : launch<128, true, true>(tmX, tmW, g, grid, st);
        else rc = np == 64 ? launch<64, false, true>(tmX, tmW, g, grid, st) : launch<128, false, true>(tmX, tmW, g, grid, st);
        if (rc != PA_OK || sk.per % sk.tk == 0) return rc;   // ranges that end on tile boundaries leave no partial tiles
        const dim3 blocks((unsigned)((N / 4 + 255) / 256), (unsigned)(rows < 32768 ? rows : 32768));
        const float* part = g.partial;
        if (np == 64)
            return launch_pdl(linear_streamk_reduce_kernel<64>, blocks, st, part, d_bias, rows, N, sk.tk, sk.per, sk.maxseg,
                              (int)n_mt, act, d_out);
        return launch_pdl(linear_streamk_reduce_kernel<128>, blocks, st, part, d_bias, rows, N, sk.tk, sk.per, sk.maxseg,
                          (int)n_mt, act, d_out);
    }
    if (nslices > 1 && !(lanes16 && ws_bytes >= (size_t)nslices * rows * N * sizeof(float))) {
        nslices = 1;
        kslice = K;
    }
    g.kslice = kslice; g.nslices = nslices;
    const dim3 grid(n_mt, n_ft, (unsigned)nslices);
    if (packed) rc = np == 64 ? launch<64, true, false>(tmX, tmW, g, grid, st) : launch<128, true, false>(tmX, tmW, g, grid, st);
    else rc = np == 64 ? launch<64, false, false>(tmX, tmW, g, grid, st) : launch<128, false, false>(tmX, tmW, g, grid, st);
    if (rc != PA_OK || nslices == 1) return rc;
    const dim3 blocks((unsigned)((((int64_t)rows * N + 3) / 4 + 255) / 256));
    const float* part = g.partial;
    return launch_pdl(linear_reduce4_kernel, blocks, st, part, d_bias, rows, N, nslices, act, d_out);
}

size_t pa_linear_tc_pack_floats(int K, int N) {
    using namespace pa::tf32x3;
    return (size_t)((N + BF - 1) / BF) * ((K + BKF - 1) / BKF) * (BF * BKF);
}

int pa_linear_tc_pack(const float* d_W, float* d_Wp, int K, int N, cudaStream_t st) {
    using namespace pa::tf32x3;
    const int64_t total4 = (int64_t)pa_linear_tc_pack_floats(K, N) / 4;
    pack_weights_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(d_W, d_Wp, K, N, total4);
    PA_RETURN_LAUNCH_STATUS();
}
