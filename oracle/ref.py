"""TEST INFRASTRUCTURE ONLY -- wrappers over oracle/_ref/libref_cpu.so.

That library is the reference's OWN int8_quant.cpp, softmax_lut.cpp and
kv_tile_cache_cpu.cpp, compiled unmodified from /root/reference by oracle/Makefile
(plus oracle/ref_shim.cpp, which only adapts std::vector signatures).  It is git-ignored
but travels to the GPU box with the gpurun snapshot.  `available()` is False when it has
not been built (no /root/reference and no prebuilt .so).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i8p = C.POINTER(C.c_int8)


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_compute_absmax.restype = C.c_float
        _lib.ref_compute_minmax_scale.restype = C.c_float
        for n in ("ref_kvcpu_f32_new", "ref_kvcpu_i8_new"):
            getattr(_lib, n).restype = C.c_void_p
        _lib.ref_kvcpu_f32_get.restype = _f32p
        _lib.ref_kvcpu_i8_get.restype = _i8p
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def quantize_to_int8(x, scale):
    x = _f32(x)
    out = np.empty(x.shape, dtype=np.int8)
    lib().ref_quantize_to_int8(_p(x, _f32p), C.c_int64(x.size), C.c_float(scale), _p(out, _i8p))
    return out


def batch_quantize(x, scales, dim):
    x, scales = _f32(x), _f32(scales)
    out = np.empty(x.shape, dtype=np.int8)
    lib().ref_batch_quantize(_p(x, _f32p), _p(scales, _f32p), scales.size, dim, _p(out, _i8p))
    return out


def compute_absmax(x):
    x = _f32(x)
    return float(lib().ref_compute_absmax(_p(x, _f32p), C.c_int64(x.size)))


def dequantize_from_int8(q, scale):
    q = np.ascontiguousarray(q, dtype=np.int8)
    out = np.empty(q.shape, dtype=np.float32)
    lib().ref_dequantize_from_int8(_p(q, _i8p), C.c_int64(q.size), C.c_float(scale), _p(out, _f32p))
    return out


def batch_dequantize(q, scales, dim):
    q = np.ascontiguousarray(q, dtype=np.int8)
    scales = _f32(scales)
    out = np.empty(q.shape, dtype=np.float32)
    lib().ref_batch_dequantize(_p(q, _i8p), _p(scales, _f32p), scales.size, dim, _p(out, _f32p))
    return out


def compute_minmax_scale(x):
    x = _f32(x)
    return float(lib().ref_compute_minmax_scale(_p(x, _f32p), C.c_int64(x.size)))


def build_exp_lut(resolution=1024, max_x=10.0):
    lut = np.empty(resolution, dtype=np.float32)
    lib().ref_build_exp_lut(resolution, C.c_float(max_x), _p(lut, _f32p))
    return lut


def _lut_call(fn, logits, scale, lut):
    logits = np.ascontiguousarray(logits, dtype=np.int32)
    assert logits.size % 8 == 0, "reference softmax_lut overruns unless len % 8 == 0"
    out = np.empty(logits.shape, dtype=np.float32)
    fn(_p(logits, _i32p), C.c_int64(logits.size), C.c_float(scale), _p(lut, _f32p), lut.size,
       _p(out, _f32p))
    return out


def softmax_lut(logits, scale, lut):
    return _lut_call(lib().ref_softmax_lut, logits, scale, lut)


def fused_softmax_lut_inplace(logits, scale, lut):
    return _lut_call(lib().ref_fused_softmax_lut_inplace, logits, scale, lut)


def softmax_batch_parallel(logits, scale, lut):
    logits = np.ascontiguousarray(logits, dtype=np.int32)
    rows, n = logits.shape
    out = np.empty(logits.shape, dtype=np.float32)
    lib().ref_softmax_batch_parallel(_p(logits, _i32p), rows, C.c_int64(n), C.c_float(scale),
                                     _p(lut, _f32p), lut.size, _p(out, _f32p))
    return out


def softmax_lut_vec(scores, temperature=1.0):
    scores = _f32(scores)
    assert scores.size % 8 == 0
    out = np.empty(scores.shape, dtype=np.float32)
    lib().ref_softmax_lut_vec(_p(scores, _f32p), scores.size, C.c_float(temperature), _p(out, _f32p))
    return out


def softmax_lut_tile(scores, temperature=1.0):
    scores = _f32(scores)
    out = np.empty(scores.shape, dtype=np.float32)
    lib().ref_softmax_lut_tile(_p(scores, _f32p), scores.size, C.c_float(temperature), _p(out, _f32p))
    return out


def apply_topk_topp_filter(probs, top_k, top_p, eos_token_id=-1, eos_thresh=0.0):
    probs = _f32(probs).copy()
    lib().ref_apply_topk_topp_filter(_p(probs, _f32p), probs.size, top_k, C.c_float(top_p),
                                     eos_token_id, C.c_float(eos_thresh))
    return probs


class KVTileCacheCPU:
    """kv_cache/kv_tile_cache_cpu.hpp:11-49 (float or int8 instantiation)."""

    def __init__(self, max_size, tile_size, dtype=np.float32):
        self._sfx = "f32" if np.dtype(dtype) == np.float32 else "i8"
        self.dtype = np.dtype(dtype)
        self.tile_size = tile_size
        self._h = C.c_void_p(getattr(lib(), f"ref_kvcpu_{self._sfx}_new")(max_size, tile_size))

    def __del__(self):
        if getattr(self, "_h", None):
            getattr(lib(), f"ref_kvcpu_{self._sfx}_free")(self._h)
            self._h = None

    def put(self, batch_id, head_id, tile_id, data):
        data = np.ascontiguousarray(data, dtype=self.dtype)
        assert data.size == self.tile_size
        getattr(lib(), f"ref_kvcpu_{self._sfx}_put")(self._h, batch_id, head_id, tile_id,
                                                    data.ctypes.data_as(C.c_void_p))

    def get(self, batch_id, head_id, tile_id):
        ptr = getattr(lib(), f"ref_kvcpu_{self._sfx}_get")(self._h, batch_id, head_id, tile_id)
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(self.tile_size,)).copy()

    def save(self, path):
        assert self._sfx == "f32"
        return lib().ref_kvcpu_f32_save(self._h, path.encode())

    def load(self, path):
        assert self._sfx == "f32"
        return lib().ref_kvcpu_f32_load(self._h, path.encode())


# ---- oracle/_ref/libref_attn.so: the reference's OWN cpu_paged_attention_forward<float> and the
# decoder's header-only LayerNorm / MLP / TokenEmbedding (oracle/build_ref_attention.py) -------------
_ATTN_PATH = os.path.join(_HERE, "_ref", "libref_attn.so")
_attn = None


def attn_available():
    return os.path.exists(_ATTN_PATH)


def attn_lib():
    global _attn
    if _attn is None:
        _attn = C.CDLL(_ATTN_PATH)
    return _attn


def cpu_paged_attention_forward(q, k_pool, v_pool, table, *, tile_size, T, beam_ids=None, temperature=1.0,
                                rope=None, top_k=0, top_p=1.0, causal=False, return_probs=False,
                                return_logits=False):
    """attention_cpu/cpu_attention_kernel.cpp:36-129, float instantiation, run on the reference's own tile
    store.  q [B,H,D]; pools [total_pages, tile_size, D] f32; table int32 [num_beams, H, num_tiles]."""
    q, k_pool, v_pool = _f32(q), _f32(k_pool), _f32(v_pool)
    table = np.ascontiguousarray(table, dtype=np.int32)
    B, H, D = q.shape
    num_beams, _, num_tiles = table.shape
    out = np.zeros((B, H, D), dtype=np.float32)
    probs = np.zeros((B, H, T), dtype=np.float32) if return_probs else None
    logits = np.zeros((B, H, T), dtype=np.float32) if return_logits else None
    beam_ids = None if beam_ids is None else np.ascontiguousarray(beam_ids, dtype=np.int32)
    rope = None if rope is None else _f32(rope)
    st = attn_lib().ref_cpu_paged_attention_f32(
        _p(q, _f32p), _p(out, _f32p), _p(k_pool, _f32p), _p(v_pool, _f32p), _p(table, _i32p), num_beams, num_tiles,
        k_pool.shape[0], _p(beam_ids, _i32p), B, H, T, D, tile_size, C.c_float(temperature), _p(rope, _f32p),
        top_k, C.c_float(top_p), 1 if causal else 0, _p(probs, _f32p), _p(logits, _f32p))
    assert st == 0, "reference cpu_paged_attention_forward threw"
    res = (out,)
    if return_probs:
        res += (probs,)
    if return_logits:
        res += (logits,)
    return res[0] if len(res) == 1 else res


def _tmp_weights(arrays):
    import tempfile
    f = tempfile.NamedTemporaryFile(suffix=".bin", delete=False)
    for a in arrays:
        f.write(_f32(a).tobytes())
    f.close()
    return f.name


def layer_norm(x, gamma, beta, eps=1e-5):
    """decoder/layer_norm.hpp:20-37 (weights through its own load_weights, :13-18)."""
    x = _f32(x)
    rows, hidden = x.shape
    out = np.empty_like(x)
    path = _tmp_weights([gamma, beta])
    try:
        st = attn_lib().ref_layer_norm_f32(path.encode(), hidden, C.c_float(eps), _p(x, _f32p), _p(out, _f32p), rows)
    finally:
        os.unlink(path)
    assert st == 0
    return out


def mlp_f32(x, fc1_w, fc1_b, fc2_w, fc2_b):
    """decoder/mlp.hpp:23-41 (weights through its own load_weights, :14-21)."""
    x = _f32(x)
    rows, hidden = x.shape
    inter = np.asarray(fc1_b).size
    out = np.empty_like(x)
    path = _tmp_weights([fc1_w, fc1_b, fc2_w, fc2_b])
    try:
        st = attn_lib().ref_mlp_f32(path.encode(), hidden, inter, _p(x, _f32p), _p(out, _f32p), rows)
    finally:
        os.unlink(path)
    assert st == 0
    return out


def token_embedding(table, ids):
    """decoder/token_embedding.hpp:19-26."""
    table = _f32(table)
    vocab, hidden = table.shape
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    out = np.empty((ids.size, hidden), dtype=np.float32)
    path = _tmp_weights([table])
    try:
        st = attn_lib().ref_token_embedding_f32(path.encode(), vocab, hidden, _p(ids, _i32p), ids.size, _p(out, _f32p))
    finally:
        os.unlink(path)
    assert st == 0
    return out
