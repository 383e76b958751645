"""TEST INFRASTRUCTURE ONLY -- numpy wrappers over oracle/liboracle_cpu.so (oracle_cpu.c).

Parity status: int8_quant / softmax_lut / filter functions are pinned bit-exact against
the reference's own objects; `paged_attention` (attention_cpu/cpu_attention_kernel.cpp:36-129) is
pinned bit-exact -- output, attention weights and logits -- against the reference's own
cpu_paged_attention_forward<float>, built by oracle/build_ref_attention.py from a temporary copy
with identifier-level fixes only (the TU does not compile as shipped); `layer_norm` / `mlp_f32`
are pinned bit-exact against decoder/layer_norm.hpp / mlp.hpp compiled unmodified
(tests/test_oracle_pinning.py, tests/golden/ref_attention_vectors.npz).  The oneDNN epilogue
(`dnnl_matmul_int8`) is PARITY UNPINNED beyond its int32 accumulators (oneDNN < 3 is absent).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_cpu.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile oracle_cpu.c (and, when /root/reference exists, oracle/_ref)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    ref_so = os.path.join(_HERE, "_ref", "libref_cpu.so")
    if os.path.isdir("/root/reference") and (force or not os.path.exists(ref_so)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference"):
        # the reference's own cpu_paged_attention_forward<float> + decoder headers (oracle/_ref/libref_attn.so)
        from . import build_ref_attention
        build_ref_attention.build(force=force)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_compute_absmax.restype = C.c_float
        _lib.orc_compute_minmax_scale.restype = C.c_float
        _lib.orc_kv_page_offset.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_quantize_weights_file.restype = C.c_float
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return int(lib().orc_num_threads())


def set_threads(n):
    """omp_set_num_threads for the oracle's loops (explicit, so OMP_NUM_THREADS=1 from torchrun does not apply)."""
    lib().orc_set_threads(int(n))
    return num_threads()


# ---- page table / pool addressing -------------------------------------------------
def pt_index(beam, head, tile, num_heads, num_tiles):
    return int(lib().orc_pt_index(int(beam), int(head), int(tile), int(num_heads), int(num_tiles)))


def pt_lookup(table, beam, head, tile, num_heads, num_tiles):
    table = _i32(table)
    return int(lib().orc_pt_lookup(_p(table, _i32p), table.size, int(beam), int(head), int(tile),
                                   int(num_heads), int(num_tiles)))


def kv_page_offset(table, beam, head, tile, num_heads, num_tiles, total_pages, tile_size, head_dim):
    table = _i32(table)
    return int(lib().orc_kv_page_offset(_p(table, _i32p), table.size, int(beam), int(head), int(tile),
                                        int(num_heads), int(num_tiles), int(total_pages),
                                        int(tile_size), int(head_dim)))


def gather_pages(pool, table, num_beams, num_heads, num_tiles, tile_size, head_dim, beam_ids=None,
                 rows=None, fill=0):
    """pool: [total_pages, tile_size, head_dim] of any dtype -> dense [R,H,num_tiles*tile_size,D]."""
    pool = np.ascontiguousarray(pool)
    table = _i32(table)
    beam_ids = _i32(beam_ids)
    R = rows if rows is not None else (len(beam_ids) if beam_ids is not None else num_beams)
    dense = np.empty((R, num_heads, num_tiles * tile_size, head_dim), dtype=pool.dtype)
    lib().orc_gather_pages(pool.ctypes.data_as(_u8p), dense.ctypes.data_as(_u8p), _p(table, _i32p),
                           num_beams, num_heads, num_tiles, pool.shape[0], tile_size, head_dim,
                           pool.dtype.itemsize, _p(beam_ids, _i32p), R, C.c_uint8(fill))
    return dense


def kv_append(k_pool, v_pool, table, num_beams, num_heads, num_tiles, tile_size, head_dim, new_k,
              new_v, positions, beam_ids=None):
    """In-place append of new_k/new_v [R,H,D] rows at `positions` [R]."""
    assert k_pool.flags.c_contiguous and v_pool.flags.c_contiguous
    table = _i32(table)
    beam_ids = _i32(beam_ids)
    positions = _i32(positions)
    new_k = np.ascontiguousarray(new_k, dtype=k_pool.dtype)
    new_v = np.ascontiguousarray(new_v, dtype=v_pool.dtype)
    lib().orc_kv_append(k_pool.ctypes.data_as(_u8p), v_pool.ctypes.data_as(_u8p), _p(table, _i32p),
                        num_beams, num_heads, num_tiles, k_pool.shape[0], tile_size, head_dim,
                        k_pool.dtype.itemsize, new_k.ctypes.data_as(_u8p), new_v.ctypes.data_as(_u8p),
                        _p(beam_ids, _i32p), _p(positions, _i32p), len(positions))


# ---- int8_quant --------------------------------------------------------------------
def quantize_to_int8(x, scale):
    x = _f32(x)
    out = np.empty(x.shape, dtype=np.int8)
    lib().orc_quantize_to_int8(_p(x, _f32p), C.c_int64(x.size), C.c_float(scale), _p(out, _i8p))
    return out


def batch_quantize(x, scales, dim):
    x = _f32(x)
    scales = _f32(scales)
    out = np.empty(x.shape, dtype=np.int8)
    lib().orc_batch_quantize(_p(x, _f32p), _p(scales, _f32p), scales.size, dim, _p(out, _i8p))
    return out


def compute_absmax(x):
    x = _f32(x)
    return float(lib().orc_compute_absmax(_p(x, _f32p), C.c_int64(x.size)))


def dequantize_from_int8(q, scale):
    q = np.ascontiguousarray(q, dtype=np.int8)
    out = np.empty(q.shape, dtype=np.float32)
    lib().orc_dequantize_from_int8(_p(q, _i8p), C.c_int64(q.size), C.c_float(scale), _p(out, _f32p))
    return out


def batch_dequantize(q, scales, dim):
    q = np.ascontiguousarray(q, dtype=np.int8)
    scales = _f32(scales)
    out = np.empty(q.shape, dtype=np.float32)
    lib().orc_batch_dequantize(_p(q, _i8p), _p(scales, _f32p), scales.size, dim, _p(out, _f32p))
    return out


def compute_minmax_scale(x):
    x = _f32(x)
    return float(lib().orc_compute_minmax_scale(_p(x, _f32p), C.c_int64(x.size)))


def batch_minmax_scale(x, dim):
    x = _f32(x)
    rows = x.size // dim
    out = np.empty(rows, dtype=np.float32)
    lib().orc_batch_minmax_scale(_p(x, _f32p), rows, dim, _p(out, _f32p))
    return out


# ---- softmax_lut family ---------------------------------------------------------------
def build_exp_lut(resolution=1024, max_x=10.0):
    lut = np.empty(resolution, dtype=np.float32)
    lib().orc_build_exp_lut(resolution, C.c_float(max_x), _p(lut, _f32p))
    return lut


def softmax_lut(logits, scale, lut):
    logits = _i32(logits)
    out = np.empty(logits.shape, dtype=np.float32)
    lib().orc_softmax_lut(_p(logits, _i32p), C.c_int64(logits.size), C.c_float(scale), _p(lut, _f32p),
                          lut.size, _p(out, _f32p))
    return out


def fused_softmax_lut(logits, scale, lut):
    logits = _i32(logits)
    out = np.empty(logits.shape, dtype=np.float32)
    lib().orc_fused_softmax_lut(_p(logits, _i32p), C.c_int64(logits.size), C.c_float(scale),
                                _p(lut, _f32p), lut.size, _p(out, _f32p))
    return out


def softmax_batch_parallel(logits, scale, lut):
    logits = _i32(logits)
    rows, n = logits.shape
    out = np.empty(logits.shape, dtype=np.float32)
    lib().orc_softmax_batch_parallel(_p(logits, _i32p), rows, C.c_int64(n), C.c_float(scale),
                                     _p(lut, _f32p), lut.size, _p(out, _f32p))
    return out


def softmax_lut_vec(scores, temperature=1.0):
    scores = _f32(scores)
    out = np.empty(scores.shape, dtype=np.float32)
    lib().orc_softmax_lut_vec(_p(scores, _f32p), scores.size, C.c_float(temperature), _p(out, _f32p))
    return out


def softmax_tile(scores, temperature=1.0):
    scores = _f32(scores)
    out = np.empty(scores.shape, dtype=np.float32)
    lib().orc_softmax_tile(_p(scores, _f32p), scores.size, C.c_float(temperature), _p(out, _f32p))
    return out


def apply_topk_topp_filter(probs, top_k, top_p, eos_token_id=-1, eos_thresh=0.0):
    probs = _f32(probs).copy()
    lib().orc_apply_topk_topp_filter(_p(probs, _f32p), probs.size, top_k, C.c_float(top_p),
                                     eos_token_id, C.c_float(eos_thresh))
    return probs


# ---- paged decode attention ----------------------------------------------------------------
def paged_attention(q, k_pool, v_pool, table, *, num_beams, num_tiles, tile_size, T=None,
                    ctx_lens=None, beam_ids=None, temperature=1.0, double_temperature=False,
                    rope=None, top_k=0, top_p=1.0, k_scales=None, v_scales=None,
                    int8_raw=False, return_probs=False, return_logits=False):
    """q [B,H,D] f32; pools [total_pages, tile_size, D] (f32, or int8 with scales
    [total_pages, tile_size] f32); table int32 [num_beams, H, num_tiles]."""
    q = _f32(q)
    B, H, D = q.shape
    table = _i32(table)
    beam_ids = _i32(beam_ids)
    ctx_lens = _i32(ctx_lens)
    if T is None:
        T = int(ctx_lens.max()) if ctx_lens is not None else num_tiles * tile_size
    if k_pool.dtype == np.int8:
        kv_kind = 2 if int8_raw else 1
        k_pool = np.ascontiguousarray(k_pool)
        v_pool = np.ascontiguousarray(v_pool)
        if kv_kind == 1:
            k_scales = _f32(k_scales)
            v_scales = _f32(v_scales)
    else:
        kv_kind = 0
        k_pool = _f32(k_pool)
        v_pool = _f32(v_pool)
    out = np.zeros((B, H, D), dtype=np.float32)
    probs = np.zeros((B, H, T), dtype=np.float32) if return_probs else None
    logits = np.zeros((B, H, T), dtype=np.float32) if return_logits else None
    rope = None if rope is None else _f32(rope)
    lib().orc_paged_attention(
        _p(q, _f32p), _p(out, _f32p), k_pool.ctypes.data_as(C.c_void_p),
        v_pool.ctypes.data_as(C.c_void_p), _p(k_scales, _f32p), _p(v_scales, _f32p), kv_kind,
        _p(table, _i32p), num_beams, H, num_tiles, k_pool.shape[0], _p(beam_ids, _i32p),
        _p(ctx_lens, _i32p), B, T, D, tile_size, C.c_float(temperature),
        1 if double_temperature else 0, _p(rope, _f32p), top_k, C.c_float(top_p),
        _p(probs, _f32p), _p(logits, _f32p))
    res = (out,)
    if return_probs:
        res += (probs,)
    if return_logits:
        res += (logits,)
    return res[0] if len(res) == 1 else res


def lse_combine(part_m, part_l, part_o):
    """part_m/l [n_parts, rows]; part_o [n_parts, rows, D] -> out [rows, D]."""
    part_m, part_l, part_o = _f32(part_m), _f32(part_l), _f32(part_o)
    n_parts, rows, D = part_o.shape
    out = np.empty((rows, D), dtype=np.float32)
    lib().orc_lse_combine(_p(part_m, _f32p), _p(part_l, _f32p), _p(part_o, _f32p), n_parts, rows, D,
                          _p(out, _f32p))
    return out


# ---- int8 matmul ----------------------------------------------------------------------------
_ACT = {"": 0, None: 0, "none": 0, "relu": 1, "gelu": 2}


def gemm_s8s8s32(A, B):
    """A [BATCH,M,K] s8, B [BATCH,K,N] s8 -> C [BATCH,M,N] s32 (exact)."""
    A = np.ascontiguousarray(A, dtype=np.int8)
    B = np.ascontiguousarray(B, dtype=np.int8)
    if A.ndim == 2:
        A, B = A[None], B[None]
    BATCH, M, K = A.shape
    N = B.shape[2]
    out = np.empty((BATCH, M, N), dtype=np.int32)
    lib().orc_gemm_s8s8s32(_p(A, _i8p), _p(B, _i8p), _p(out, _i32p), BATCH, M, N, K)
    return out


def matmul_int8_epilogue(acc, scaleA, scaleB, scaleC=1.0, bias=None, activation=""):
    acc = _i32(acc)
    N = acc.shape[-1]
    out = np.empty(acc.shape, dtype=np.int8)
    bias = None if bias is None else _f32(bias)
    lib().orc_matmul_int8_epilogue(_p(acc, _i32p), _p(out, _i8p), C.c_int64(acc.size // N), N,
                                   C.c_float(scaleA), C.c_float(scaleB), C.c_float(scaleC),
                                   _p(bias, _f32p), _ACT[activation])
    return out


def dnnl_matmul_int8(A, B, scaleA, scaleB, scaleC=1.0, bias=None, activation=""):
    return matmul_int8_epilogue(gemm_s8s8s32(A, B), scaleA, scaleB, scaleC, bias, activation)


# ---- decoder glue ---------------------------------------------------------------------------
def layer_norm(x, gamma, beta, eps=1e-5):
    x = _f32(x)
    rows, hidden = x.shape
    out = np.empty_like(x)
    gamma, beta = _f32(gamma), _f32(beta)
    lib().orc_layer_norm(_p(x, _f32p), _p(out, _f32p), _p(gamma, _f32p), _p(beta, _f32p), rows,
                         hidden, C.c_float(eps))
    return out


def mlp_f32(x, fc1_w, fc1_b, fc2_w, fc2_b):
    x = _f32(x)
    rows, hidden = x.shape
    inter = fc1_b.size
    fc1_w, fc1_b, fc2_w, fc2_b = _f32(fc1_w), _f32(fc1_b), _f32(fc2_w), _f32(fc2_b)
    out = np.empty_like(x)
    lib().orc_mlp_f32(_p(x, _f32p), _p(out, _f32p), _p(fc1_w, _f32p), _p(fc1_b, _f32p),
                      _p(fc2_w, _f32p), _p(fc2_b, _f32p), rows, hidden, inter)
    return out


def quantize_weights_file(w):
    """decoder/int8_decoder.cpp:52-56 -> (int8 array, scale)."""
    w = _f32(w).reshape(-1)
    out = np.empty(w.size, dtype=np.int8)
    scale = lib().orc_quantize_weights_file(_p(w, _f32p), C.c_int64(w.size), _p(out, _i8p))
    return out, float(scale)


def apply_rotary_embedding(q, k, rotary_emb, positions):
    """attention_kernel_utils.cuh:20-35 over rows: q, k [rows, H, D] (copies returned), rotary_emb [T, D]."""
    q, k = _f32(q).copy(), _f32(k).copy()
    rope = _f32(rotary_emb)
    rows, H, D = q.shape
    for r in range(rows):
        for h in range(H):
            lib().orc_apply_rotary_embedding(_p(q[r, h], _f32p), _p(k[r, h], _f32p), _p(rope, _f32p), D,
                                             int(positions[r]), 1)
    return q, k
