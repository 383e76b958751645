"""Generate tests/golden/ref_vectors.npz by RUNNING the reference's own objects
(oracle/_ref/libref_cpu.so = unmodified /root/reference int8_quant.cpp, softmax_lut.cpp,
kv_tile_cache_cpu.cpp).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The .npz is committed so CPU tests can pin oracle_cpu.c even where _ref is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

oracle.cpu.build()
ref = oracle.ref
rng = np.random.default_rng(20261018)
g = {}

# int8_quant.cpp
x = (rng.standard_normal(4096) * 3).astype(np.float32)
x[:13] = [0.5, -0.5, 1.5, 2.5, -2.5, 0.49999997, 126.5, 127.5, -128.5, -129, 1000, -1000, 0]
g["q_x"] = x
g["q_scale1"] = ref.quantize_to_int8(x, 1.0)
s = ref.compute_minmax_scale(x[13:])
g["q_minmax_scale"] = np.float32(s)
g["q_absmax"] = np.float32(ref.compute_absmax(x[13:]))
g["q_scaled"] = ref.quantize_to_int8(x[13:], s)
g["dq_scaled"] = ref.dequantize_from_int8(g["q_scaled"], s)
rows, dim = 64, 128
xb = rng.standard_normal((rows, dim)).astype(np.float32)
sc = np.array([ref.compute_minmax_scale(r) for r in xb], dtype=np.float32)
g["bq_x"], g["bq_scales"] = xb, sc
g["bq_q"] = ref.batch_quantize(xb, sc, dim)
g["bq_dq"] = ref.batch_dequantize(g["bq_q"], sc, dim)

# softmax_lut.cpp
lut = ref.build_exp_lut(1024, 10.0)
g["lut"] = lut
for i, (scale, n) in enumerate([(0.001, 512), (0.01, 512), (0.05, 4096)]):
    lg = rng.integers(-4000, 4000, size=n, dtype=np.int32)
    g[f"sl_logits{i}"] = lg
    g[f"sl_scale{i}"] = np.float32(scale)
    g[f"sl_out{i}"] = ref.softmax_lut(lg, scale, lut)
    g[f"sl_fused{i}"] = ref.fused_softmax_lut_inplace(lg, scale, lut)
sv = (rng.standard_normal(512) * 4).astype(np.float32)
g["slv_scores"] = sv
g["slv_t1"] = ref.softmax_lut_vec(sv, 1.0)
g["slv_t2"] = ref.softmax_lut_vec(sv, 2.0)
g["slt_t2"] = ref.softmax_lut_tile(sv[:64], 2.0)
pf = g["slv_t1"][:64].copy()
pf = pf / pf.sum()
g["flt_in"] = pf.astype(np.float32)
g["flt_k5"] = ref.apply_topk_topp_filter(pf, 5, 1.0)
g["flt_p6"] = ref.apply_topk_topp_filter(pf, 0, 0.6)
g["flt_k5_p3_eos"] = ref.apply_topk_topp_filter(pf, 5, 0.3, int(np.argmax(pf)), 0.01)

np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_vectors.npz"), **g)
print("wrote", len(g), "arrays")
