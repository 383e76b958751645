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

// One WARP receives the `world` partials of `row` from its own buffer, LSE-combines them and writes the row:
//   M = max m_s;  w_s = exp(m_s - M);  out = sum w_s O_s / (sum w_s l_s + 1e-6)      (orc_lse_combine)
// On timeout the row is written as NaN, *status is set and the epoch is NOT advanced.
template <int D>
__device__ __forceinline__ void recv_row(uint8_t* const* peers, uint32_t* epochs, int rank, int world, int64_t rows,
                                         int64_t row, float* __restrict__ out, float* __restrict__ lse_out,
                                         int* __restrict__ status, int lane) {
    constexpr int VEC = D / 32, PK = D / 2 + 1;
    const uint32_t epoch = epochs[row] + 1u;
    const uint4* mine = reinterpret_cast<const uint4*>(peers[rank]);
    const int par = (int)(epoch & 1u);
    const long long t0 = clock64();
    bool ok = true;
    float ms = -INFINITY, ls = 0.f;
    for (int s = lane; s < world; s += 32)  // world <= 32: one source per lane
        ok = wait_packet(mine + slot_index(par, world, s, rows, row, PK) + D / 2, epoch, ms, ls, t0);
    const float Mg = warp_max(ms);
    const float wl = (ms == -INFINITY) ? 0.f : __expf(ms - Mg);
    const float Lg = warp_sum(ls * wl);
    float Og[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) Og[e] = 0.f;
    for (int s = 0; s < world; ++s) {
        const float w = __shfl_sync(0xffffffffu, wl, s);
        const uint4* src = mine + slot_index(par, world, s, rows, row, PK);
#pragma unroll
        for (int e = 0; e < VEC; e += 2) {
            float f0 = 0.f, f1 = 0.f;
            ok = wait_packet(src + (lane * VEC + e) / 2, epoch, f0, f1, t0) && ok;
            Og[e] = fmaf(f0, w, Og[e]);
            Og[e + 1] = fmaf(f1, w, Og[e + 1]);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
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
