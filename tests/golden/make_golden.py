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

# ---- cpu_paged_attention_forward<float> (oracle/_ref/libref_attn.so = the reference's own
# cpu_attention_kernel.cpp:36-129 compiled by oracle/build_ref_attention.py) and the decoder's
# header-only LayerNorm / MLP / TokenEmbedding.  Stored in a second file so the first stays
# byte-identical to the r01 fixture.
a = {}
arng = np.random.default_rng(20261019)
cases = [  # (B, H, D, T, tile, temperature, unmapped, beam_ids, rope, top_k, top_p)
    (2, 2, 64, 104, 16, 1.0, True, False, False, 0, 1.0),
    (3, 1, 128, 48, 16, 11.3137085, False, True, False, 0, 1.0),
    (1, 3, 64, 40, 8, 2.0, True, False, True, 0, 1.0),
    (2, 2, 32, 64, 16, 1.0, False, False, False, 5, 1.0),
    (2, 2, 32, 64, 16, 4.0, False, True, True, 0, 0.7),
]
for i, (B, H, D, T, tile, temp, unmapped, beams, rope, top_k, top_p) in enumerate(cases):
    nt = (T + tile - 1) // tile
    P = B * H * nt + 3
    # K/V values are fp16-representable and stored as fp16 (exact; halves the fixture)
    k = arng.standard_normal((P, tile, D)).astype(np.float16).astype(np.float32)
    v = arng.standard_normal((P, tile, D)).astype(np.float16).astype(np.float32)
    q = arng.standard_normal((B, H, D)).astype(np.float32)
    table = arng.permutation(P)[:B * H * nt].astype(np.int32).reshape(B, H, nt)
    if unmapped:
        table[0, H - 1, nt // 2] = -1
        table[B - 1, 0, 0] = P + 7  # out of range: never stored -> skipped
    beam_ids = arng.permutation(B).astype(np.int32) if beams else None
    rp = arng.standard_normal(D).astype(np.float32) if rope else None
    o, p, lg = ref.cpu_paged_attention_forward(q, k, v, table, tile_size=tile, T=T, beam_ids=beam_ids,
                                               temperature=temp, rope=rp, top_k=top_k, top_p=top_p,
                                               return_probs=True, return_logits=True)
    a[f"at{i}_cfg"] = np.array([B, H, D, T, tile, top_k], np.int32)
    a[f"at{i}_f"] = np.array([temp, top_p], np.float32)
    a[f"at{i}_q"], a[f"at{i}_k"], a[f"at{i}_v"], a[f"at{i}_table"] = q, k.astype(np.float16), v.astype(np.float16), table
    if beam_ids is not None:
        a[f"at{i}_beam_ids"] = beam_ids
    if rp is not None:
        a[f"at{i}_rope"] = rp
    a[f"at{i}_out"], a[f"at{i}_probs"], a[f"at{i}_logits"] = o, p, lg
hid, inter, rows, vocab = 48, 192, 5, 37
x = (arng.standard_normal((rows, hid)) * 3).astype(np.float32)
gam, bet = arng.standard_normal(hid).astype(np.float32), arng.standard_normal(hid).astype(np.float32)
w1 = (arng.standard_normal((hid, inter)) * 0.2).astype(np.float32)
b1 = arng.standard_normal(inter).astype(np.float32)
w2 = (arng.standard_normal((inter, hid)) * 0.2).astype(np.float32)
b2 = arng.standard_normal(hid).astype(np.float32)
emb = arng.standard_normal((vocab, hid)).astype(np.float32)
ids = arng.integers(0, vocab, 9).astype(np.int32)
a.update(ln_x=x, ln_g=gam, ln_b=bet, ln_out=ref.layer_norm(x, gam, bet, 1e-5),
         mlp_w1=w1, mlp_b1=b1, mlp_w2=w2, mlp_b2=b2, mlp_out=ref.mlp_f32(x, w1, b1, w2, b2),
         emb_table=emb, emb_ids=ids, emb_out=ref.token_embedding(emb, ids))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_attention_vectors.npz"), **a)
print("wrote", len(a), "arrays (attention / decoder headers)")
