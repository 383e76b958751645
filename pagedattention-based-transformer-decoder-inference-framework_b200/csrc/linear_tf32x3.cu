// linear_tf32x3.cu -- the float MLP / projection layer of the fp32 decoder (decoder/mlp.hpp:23-41:
// out[r,n] = act(bias[n] + sum_k x[r,k] * W[k*N+n])) on the 5th-gen tensor cores, at fp32-class accuracy.
//
// tcgen05.mma kind::tf32 reads 32-bit containers and uses sign, 8 exponent and the top 10 mantissa bits.  Each
// operand is therefore split exactly into  v = hi + lo,  hi = v with the low 13 mantissa bits cleared (what the
// tensor core sees of v itself), lo = v - hi (exact in fp32, 13 significant bits of which the core keeps 11), and
//     x . W  ~=  x_lo . W_hi + x_hi . W_lo + x_hi . W_hi                     ("3xTF32")
// accumulated in fp32 in TMEM.  Dropped: lo.lo (2^-22 relative) and the tail of each lo (2^-21): per-product error
// <= ~1e-6 of |x||W|, the same order as fp32 summation-order noise of the SIMT kernels it replaces (tests state
// 2e-6 * sum_k |x||W| + fp32 eps of the result).  x and W are consumed where they lie: x [rows, K] is the K-major A
// operand, W [K, N] (the reference's layout) the MN-major B operand -- no transposed or pre-split copy of the
// weights, which stream from HBM exactly once per 128-row tile of x.
//
// One CTA = one 128 x 128 output tile over one K slice; 192 threads:
//   warp 0    TMA producer: per 32-float K block the x tile [128 rows][128 B] and four W boxes [32 k][32 n]
//             (128-byte swizzle -- 32-byte atoms for W --, out-of-range rows / columns / k zero-filled) into a 4-stage ring;
//   warps 2-5 split: thread = row of x: its 32 floats of the block go to TENSOR MEMORY as two A operands (x as it is
//             -- the core ignores the low 13 bits -- and x_lo), so the three MMAs of a k-step read x from TMEM instead
//             of three times from shared memory; the W tile gets its lo tile beside it (same swizzled offsets, so the
//             split never needs the layout); fence to the async proxy, arrive.  After the main loop the same warps
//             are the epilogue (tcgen05.ld of their TMEM lane quarter, bias / relu or K-slice partial, 16-byte stores);
//   warp 1    one elected lane issues 12 tcgen05.mma (4 k-steps of 8 x 3 terms, A from TMEM, B from shared memory)
//             per stage, commit frees the stage.
// Shared-memory traffic per stage is what bounds the loop next to the tensor pipe: 48 KB of W operand reads + 48 KB
// of split traffic (x 16, W 16 read, W_lo 16 written) per 16 KB of weights streamed from HBM.
// K slices (grid.z) write fp32 partial tiles that linear_reduce_kernel (decoder_ops.cu) sums in slice order.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "pa_common.cuh"

namespace pa {
namespace tf32x3 {

constexpr int BM = 128;            // rows per tile = UMMA M = TMEM lanes
constexpr int BN = 128;            // columns per tile = UMMA N = TMEM columns
constexpr int BKF = 32;            // floats of K per stage (one 128-byte swizzle row)
constexpr int TILE = BM * 128;     // 16 KiB: one operand tile
constexpr int STAGE = 3 * TILE;    // x, W, W_lo
constexpr int NST = 4;
constexpr int ACC_COLS = BN;                 // TMEM: accumulator, then per stage x (32 columns) and x_lo (32 columns)
constexpr int TMEM_COLS = 512;               // 128 + 4 * 64 = 384 -> next power of two
constexpr int NTHREADS = 192;
constexpr int W_BOX = BKF * 128;   // one W box: 32 k-rows x 32 columns = 4 KiB; four of them side by side per tile

struct Args {
    const float* bias;
    float* out;
    float* partial;   // [nslices][rows][N] when nslices > 1
    int rows, N, K, act, kslice, nslices;
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
// UMMA shared-memory descriptor, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14), LBO >> 4 [16,30),
// SBO >> 4 [32,46), version 1 [46,48), layout [61,64): SWIZZLE_128B = 2, SWIZZLE_128B_BASE32B = 1.
//   K-major  (x):  rows of 128 B (32 floats of K); 8-row groups 1024 B apart (SBO); LBO unused (one swizzle row of K).
//   MN-major (W):  32-bit operands transpose at 32-byte granularity, so the layout is SWIZZLE_128B_BASE32B = 1
//                  (Swizzle<2,5,2>; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), canonical form
//                  ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp:246): 128 B of N
//                  (32 floats) contiguous, the next 32 columns LBO bytes on, k-rows 128 B apart, 4-row groups SBO = 512
//                  bytes apart.  (The plain 128-byte swizzle, layout 2, is accepted for MN-major tf32 but the tensor
//                  core then returns zeros -- measured during bring-up.)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout = 2) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// Instruction descriptor: c = F32 (1) [4,6), a = b = TF32 (2) [7,10) [10,13), a K-major (0) [15], b MN-major (1) [16],
// N >> 3 [17,23), M >> 4 [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) |
                            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// v - (v with the low 13 mantissa bits cleared); 0 for inf / nan inputs' hi part is the value itself (inf - inf
// would poison finite columns of the same row with NaN through the lo terms, so non-finite values get lo = 0).
__device__ __forceinline__ float lo_part(float v) {
    const uint32_t u = __float_as_uint(v);
    const float hi = __uint_as_float(u & 0xffffe000u);
    return ((u & 0x7f800000u) == 0x7f800000u) ? 0.f : v - hi;
}

__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    // a lost arrival would otherwise hang the GPU box: ~2 s of polling, then trap
    for (uint32_t i = 0; i < (1u << 26); ++i)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

__global__ void __launch_bounds__(NTHREADS, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Args g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + NST * STAGE;
    auto full_bar = [&](int s) { return bar0 + s * 8; };
    auto split_bar = [&](int s) { return bar0 + (NST + s) * 8; };
    auto empty_bar = [&](int s) { return bar0 + (2 * NST + s) * 8; };
    const uint32_t tmem_full_bar = bar0 + 3 * NST * 8;
    const uint32_t tmem_slot = tmem_full_bar + 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, slice = blockIdx.z;
    const int k_begin = slice * g.kslice;
    const int k_end = min(g.K, k_begin + g.kslice);
    const int nkb = (k_end - k_begin + BKF - 1) / BKF;   // kslice is a multiple of BKF: blocks never straddle slices

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(split_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW));
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NST;
                mbar_wait_bounded(empty_bar(s), ((i / NST) & 1) ^ 1);
                const uint32_t st = base + s * STAGE;
                mbar_arrive_expect_tx(full_bar(s), 2 * TILE);
                const int k0 = k_begin + i * BKF;
                tma_load_2d(st, &tmX, k0, m0, full_bar(s));
#pragma unroll
                for (int j = 0; j < BN / 32; ++j) tma_load_2d(st + TILE + j * W_BOX, &tmW, n0 + j * 32, k0, full_bar(s));
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NST;
                mbar_wait_bounded(split_bar(s), (i / NST) & 1);
                tc_fence_after();
                const uint32_t st = base + s * STAGE;
#pragma unroll
                for (int ks = 0; ks < BKF / 8; ++ks) {
                    const uint32_t xa = tmem_base + ACC_COLS + s * 64 + ks * 8;
                    const uint32_t xl = xa + 32;
                    const uint64_t wa = make_desc(st + TILE + ks * 1024, W_BOX, 512, 1);
                    const uint64_t wl = make_desc(st + 2 * TILE + ks * 1024, W_BOX, 512, 1);
                    umma_tf32_ts(tmem_base, xl, wa, kIdesc, (i > 0 || ks > 0) ? 1u : 0u);   // small terms first
                    umma_tf32_ts(tmem_base, xa, wl, kIdesc, 1u);
                    umma_tf32_ts(tmem_base, xa, wa, kIdesc, 1u);
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        const int t = threadIdx.x - 64;   // 0..127
        for (int i = 0; i < nkb; ++i) {
            const int s = i % NST;
            mbar_wait_bounded(full_bar(s), (i / NST) & 1);
            const uint32_t st = base + s * STAGE;
            // x: this thread's row (TMEM lane) of the block: 8 swizzled 16-byte chunks -> 32 columns of x and of x_lo
            {
                const int r = (warp & 3) * 32 + lane;
                uint32_t xv[32], xlo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 v = lds_128(st + r * 128 + ((c ^ (r & 7)) << 4));
                    xv[4 * c] = v.x; xv[4 * c + 1] = v.y; xv[4 * c + 2] = v.z; xv[4 * c + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) xlo[j] = __float_as_uint(lo_part(__uint_as_float(xv[j])));
                const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ACC_COLS + s * 64;
                tmem_st32(ta, xv);
                tmem_st32(ta + 32, xlo);
            }
            // W tile -> W_lo: 1024 chunks of 16 bytes, 8 per thread
            {
                const uint32_t src = st + TILE, dst = src + TILE;
                uint4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = lds_128(src + (j * 128 + t) * 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float a = lo_part(__uint_as_float(v[j].x)), b = lo_part(__uint_as_float(v[j].y));
                    const float c = lo_part(__uint_as_float(v[j].z)), d = lo_part(__uint_as_float(v[j].w));
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst + (j * 128 + t) * 16), "f"(a), "f"(b),
                                 "f"(c), "f"(d)
                                 : "memory");
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(split_bar(s));
        }
        // ---- epilogue: this warp's TMEM lane quarter, 32 columns at a time ----
        const int qtr = warp & 3;
        const int row = m0 + qtr * 32 + lane;
        const bool row_ok = row < g.rows;
        if (nkb > 0) {
            mbar_wait_bounded(tmem_full_bar, 0);
            tc_fence_after();
        }
        float* dst_row = g.nslices > 1 ? g.partial + ((int64_t)slice * g.rows + row) * g.N : g.out + (int64_t)row * g.N;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            if (nkb > 0) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(qtr * 32) << 16) + c * 32, r);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (!row_ok) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int n = n0 + c * 32 + j;
                if (n >= g.N) break;   // N % 4 == 0: whole float4 groups
                float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                if (g.nslices == 1) {
                    if (g.bias) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                    }
                    if (g.act == PA_ACT_RELU) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    }
                }
                *reinterpret_cast<float4*>(dst_row + n) = v;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
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
// 2-D f32 tensor [rows][cols] (cols contiguous), box [box_rows][32 floats], 128-byte swizzle, zero fill out of range.
static bool make_map_f32(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows,
                         CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tf32x3
}  // namespace pa

using namespace pa;

// K-slice geometry.  One CTA per SM is resident (192 KB ring), so the grid runs in waves of sm_count CTAs: pick the
// slice count that minimises waves x K blocks per CTA (+ half a block per slice for the partial-tile traffic);
// every slice holds >= 8 K blocks (256 k-rows).
int pa_linear_tc_slices(int rows, int K, int N, int sm_count, int* kslice_out) {
    using namespace pa::tf32x3;
    const int64_t tiles = (int64_t)((rows + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int total_kb = (K + BKF - 1) / BKF;
    int max_ns = total_kb / 8;
    if (max_ns > 16) max_ns = 16;
    if (max_ns < 1) max_ns = 1;
    int best_ns = 1;
    double best_cost = 1e30;
    for (int ns = 1; ns <= max_ns; ++ns) {
        const int64_t waves = (tiles * ns + sm_count - 1) / sm_count;
        const double cost = (double)waves * ((total_kb + ns - 1) / ns + 2) + 0.5 * (ns - 1);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best_ns = ns;
        }
    }
    int kslice = ((total_kb + best_ns - 1) / best_ns) * BKF;
    if (kslice_out) *kslice_out = kslice;
    return (K + kslice - 1) / kslice;
}

// Launch helper for pa_linear_f32 (decoder_ops.cu).  PA_ERR_UNSUPPORTED when the tensor maps cannot be built.
int pa_linear_tc_launch(const float* d_x, const float* d_W, const float* d_bias, int rows, int K, int N, int act,
                        float* d_out, float* d_partial, int nslices, int kslice, cudaStream_t st) {
    using namespace pa::tf32x3;
    CUtensorMap tmX, tmW;
    if (!make_map_f32(&tmX, d_x, (uint64_t)K, (uint64_t)rows, BM, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map_f32(&tmW, d_W, (uint64_t)N, (uint64_t)K, BKF, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return PA_ERR_UNSUPPORTED;
    const size_t smem = (size_t)NST * STAGE + (3 * NST + 1) * 8 + 16 + 1024;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev & 63] = true;
    }
    Args g{d_bias, d_out, d_partial, rows, N, K, act, kslice, nslices};
    dim3 grid((unsigned)((rows + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)nslices);
    linear_tf32x3_kernel<<<grid, NTHREADS, smem, st>>>(tmX, tmW, g);
    PA_RETURN_LAUNCH_STATUS();
}
