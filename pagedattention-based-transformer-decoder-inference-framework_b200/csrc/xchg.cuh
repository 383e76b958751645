// xchg.cuh -- inter-GPU exchange of split-KV row partials over NVLink peer memory, flag-in-data packets.
//
// A row partial is (O[D], m, l) = D + 2 floats.  It travels as PK = D/2 + 1 packets of 16 bytes,
//     { f0, epoch, f1, epoch }                         (the layout of NCCL's LL protocol)
// written with ONE 16-byte store each.  Every 8-byte half carries its own copy of the epoch, and
// 8-byte-aligned halves of a store are delivered atomically, so a receiver that sees both flags equal to
// the epoch it waits for has the data: no __threadfence_system() between data and flag, no separate flag
// store, no second NVLink round trip.  SENDING NEVER BLOCKS; only the receive side polls.  That keeps any
// schedule deadlock-free: a warp that still has chunks to stream is never parked behind a peer.
//
// Exchange buffer of one rank (identical on all ranks, CUDA-IPC mapped into every peer):
//     uint4 [2 parity][world src][rows][PK]
// Parity = epoch & 1: a rank can run at most one step ahead of its slowest peer (it needs every peer's
// step-e data to finish step e, and a peer sends step e+1 only after it has finished reading step e), so two
// buffers suffice.  Epochs live in device memory (one counter per row) so launches replay inside CUDA graphs.
#pragma once
#include "pa_common.cuh"

namespace pa {
namespace xchg {

__host__ __device__ inline int packets_per_row(int D) { return D / 2 + 1; }
__host__ __device__ inline size_t buffer_bytes(int world, int64_t rows, int D) {
    return (size_t)2 * world * rows * packets_per_row(D) * sizeof(uint4);
}
__device__ __forceinline__ size_t slot_index(int par, int world, int src, int64_t rows, int64_t row, int pk) {
    return ((size_t)(par * world + src) * rows + row) * pk;
}

__device__ __forceinline__ void st_packet(uint4* p, float f0, float f1, uint32_t epoch) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(p), "r"(__float_as_uint(f0)), "r"(epoch),
                 "r"(__float_as_uint(f1))
                 : "memory");
}
__device__ __forceinline__ uint4 ld_packet(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// Poll one packet until both halves carry `epoch`; false on timeout (~2 s: a missing peer must not hang the GPU).
__device__ __forceinline__ bool wait_packet(const uint4* p, uint32_t epoch, float& f0, float& f1, long long t0) {
    for (;;) {
        const uint4 v = ld_packet(p);
        if (v.y == epoch && v.w == epoch) {
            f0 = __uint_as_float(v.x);
            f1 = __uint_as_float(v.z);
            return true;
        }
        if (clock64() - t0 > 4000000000ll) return false;
    }
}

// One WARP sends the merged partial of `row` to every rank (itself included).  Lane l holds O[l*VEC .. l*VEC+VEC)
// (VEC = D/32); M in natural-log units.  Returns the epoch used.
template <int D>
__device__ __forceinline__ uint32_t send_row(uint8_t* const* peers, const uint32_t* epochs, int rank, int world,
                                             int64_t rows, int64_t row, const float (&O)[D / 32], float m_nat, float L,
                                             int lane) {
    constexpr int VEC = D / 32, PK = D / 2 + 1;
    const uint32_t epoch = epochs[row] + 1u;
    const size_t slot = slot_index((int)(epoch & 1u), world, rank, rows, row, PK);
    for (int p = 0; p < world; ++p) {
        uint4* dst = reinterpret_cast<uint4*>(peers[p]) + slot;
#pragma unroll
        for (int e = 0; e < VEC; e += 2) st_packet(dst + (lane * VEC + e) / 2, O[e], O[e + 1], epoch);
        if (lane == 0) st_packet(dst + D / 2, m_nat, L, epoch);
    }
    return epoch;
}

// One WARP gathers `nsrc` (<= 64) partial rows stored as packets at base + s * stride (s = source), waits for their
// `epoch` tags and LSE-combines them WITHOUT normalising:
//   M = max m_s;  w_s = exp(m_s - M);  O = sum w_s O_s;  L = sum w_s l_s          (m in natural-log units)
// Lane l ends up with O for dims l*VEC .. l*VEC+VEC.  Sources are fetched in groups of 8 with ALL packet loads of a
// group issued before the first tag is looked at, so a group costs one memory round trip when the data is there
// (one dependent round trip per packet otherwise: ~12 us for 8 peers); packets that have not landed yet are
// re-polled one by one.  Returns false on timeout.
template <int D>
__device__ __forceinline__ bool gather_combine(const uint4* base, int nsrc, size_t stride, uint32_t epoch, int lane,
                                               float& Mg, float& Lg, float (&Og)[D / 32]) {
    constexpr int VEC = D / 32, NP = VEC / 2, SG = 8;
    const long long t0 = clock64();
    bool ok = true;
    float ms[2] = {-INFINITY, -INFINITY}, ls[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int s = c * 32 + lane;
        if (s < nsrc) ok = wait_packet(base + s * stride + D / 2, epoch, ms[c], ls[c], t0) && ok;
    }
    Mg = warp_max(fmaxf(ms[0], ms[1]));
    float wl[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) wl[c] = (ms[c] == -INFINITY) ? 0.f : __expf(ms[c] - Mg);
    Lg = warp_sum(ls[0] * wl[0] + ls[1] * wl[1]);
#pragma unroll
    for (int e = 0; e < VEC; ++e) Og[e] = 0.f;
    for (int s0 = 0; s0 < nsrc; s0 += SG) {
        uint4 pk[SG][NP];
#pragma unroll
        for (int i = 0; i < SG; ++i) {
            if (s0 + i < nsrc) {
                const uint4* src = base + (s0 + i) * stride + (lane * VEC) / 2;
#pragma unroll
                for (int e = 0; e < NP; ++e) pk[i][e] = ld_packet(src + e);
            }
        }
#pragma unroll
        for (int i = 0; i < SG; ++i) {
            if (s0 + i < nsrc) {
                const int s = s0 + i;
                const float w = __shfl_sync(0xffffffffu, s < 32 ? wl[0] : wl[1], s & 31);
                const uint4* src = base + s * stride + (lane * VEC) / 2;
#pragma unroll
                for (int e = 0; e < NP; ++e) {
                    float f0 = __uint_as_float(pk[i][e].x), f1 = __uint_as_float(pk[i][e].z);
                    if (pk[i][e].y != epoch || pk[i][e].w != epoch) ok = wait_packet(src + e, epoch, f0, f1, t0) && ok;
                    Og[2 * e] = fmaf(f0, w, Og[2 * e]);
                    Og[2 * e + 1] = fmaf(f1, w, Og[2 * e + 1]);
                }
            }
        }
    }
    return __all_sync(0xffffffffu, ok);
}

// One WARP receives the `world` partials of `row` from its own buffer, LSE-combines them and writes the row:
//   out = sum w_s O_s / (sum w_s l_s + 1e-6)      (orc_lse_combine)
// On timeout the row is written as NaN, *status is set and the epoch is NOT advanced.
template <int D>
__device__ __forceinline__ void recv_row(uint8_t* const* peers, uint32_t* epochs, int rank, int world, int64_t rows,
                                         int64_t row, float* __restrict__ out, float* __restrict__ lse_out,
                                         int* __restrict__ status, int lane) {
    constexpr int VEC = D / 32, PK = D / 2 + 1;
    const uint32_t epoch = epochs[row] + 1u;
    const uint4* mine = reinterpret_cast<const uint4*>(peers[rank]);
    float Mg, Lg, Og[VEC];
    const bool ok = gather_combine<D>(mine + slot_index((int)(epoch & 1u), world, 0, rows, row, PK), world,
                                      (size_t)rows * PK, epoch, lane, Mg, Lg, Og);
    const float inv = ok ? 1.f / (Lg + 1e-6f) : __int_as_float(0x7fc00000);
#pragma unroll
    for (int e = 0; e < VEC; ++e) out[row * D + lane * VEC + e] = Og[e] * inv;
    if (lane == 0) {
        if (lse_out) lse_out[row] = !ok ? __int_as_float(0x7fc00000) : ((Lg > 0.f) ? Mg + logf(Lg) : -INFINITY);
        if (ok) epochs[row] = epoch;
        else if (status) atomicExch(status, 1);
    }
}

}  // namespace xchg
}  // namespace pa
