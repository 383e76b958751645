// ref_attention_shim.hpp -- TEST INFRASTRUCTURE ONLY.
//
// The two helper types attention_cpu/cpu_attention_kernel.cpp:36-129 is written against but
// the reference never defines: a D-element float vector with 2-argument
// load/store ("will decode if int8_t", cpu_attention_kernel.cpp:52-53, 80, 120) and a tile
// store whose `get` takes a fourth 'k'/'v' argument (cpu_attention_kernel.cpp:72, 107).
// They carry no arithmetic of their own: load/store are element-wise casts, the store
// adaptor forwards to the reference's OWN KVTileCacheCPU<T>::get (kv_tile_cache_cpu.cpp:70-80)
// with the K/V choice folded into the tile id (K -> 2*tile, V -> 2*tile + 1).
// Force-included (-include) in front of the patched copy of the reference TU that
// oracle/build_ref_attention.py generates at build time.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "kv_cache/kv_tile_cache_cpu.hpp"

namespace refshim {

template <typename T>
struct Vec {
    std::vector<T> v;
    template <typename S>
    void load(const S* p, int n) {
        v.resize((size_t)n);
        for (int i = 0; i < n; ++i) v[(size_t)i] = static_cast<T>(p[i]);
    }
    template <typename S>
    void store(S* p, int n) const {
        for (int i = 0; i < n; ++i) p[i] = (size_t)i < v.size() ? static_cast<S>(v[(size_t)i]) : S(0);
    }
    void clear() { std::fill(v.begin(), v.end(), T(0)); }
    // `Vec<float> out_vec; out_vec.clear(); out_vec[d] += ...` (cpu_attention_kernel.cpp:99-116): a zero
    // vector that is as long as the indices used on it.
    T& operator[](int i) {
        if ((size_t)i >= v.size()) v.resize((size_t)i + 1, T(0));
        return v[(size_t)i];
    }
};

template <typename T>
struct KVTileStore4 {
    KVTileCacheCPU<T>* impl;
    const T* get(int beam_id, int head_id, int tile_id, char which) {
        return impl->get(beam_id, head_id, 2 * tile_id + (which == 'v' ? 1 : 0));
    }
};

}  // namespace refshim
