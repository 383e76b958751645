"""CPU: the attention restatement against an independent float64 numpy model, and the
split-KV / LSE-combine decomposition used by the kernels."""
import numpy as np
import pytest

from synth import make_case, oracle_attention


def numpy_attention(case):
    B, H, D = case["q"].shape
    ts, nt = case["tile_size"], case["num_tiles"]
    out = np.zeros((B, H, D))
    kp = case["k_pool"].astype(np.float64)
    vp = case["v_pool"].astype(np.float64)
    if case["kv"] == "i8":
        kp = kp / case["k_scales"][..., None]
        vp = vp / case["v_scales"][..., None]
    for b in range(B):
        ctx = case["T"] if case["ctx_lens"] is None else int(case["ctx_lens"][b])
        beam = b if case["beam_ids"] is None else int(case["beam_ids"][b])
        for h in range(H):
            ks, vs = [], []
            for t in range((ctx + ts - 1) // ts):
                pg = case["table"][beam, h, t]
                if pg < 0 or pg >= case["total_pages"]:
                    continue
                n = min(ts, ctx - t * ts)
                ks.append(kp[pg, :n]); vs.append(vp[pg, :n])
            if not ks:
                continue
            K, V = np.concatenate(ks), np.concatenate(vs)
            s = K @ case["q"][b, h].astype(np.float64) / case["temperature"]
            p = np.exp(s - s.max()); p /= p.sum()
            out[b, h] = p @ V
    return out


@pytest.mark.parametrize("kv", ["f16", "i8"])
@pytest.mark.parametrize("opts", [dict(), dict(unmapped_frac=0.05), dict(ragged=True),
                                  dict(beam_width=4, shared_prefix=64)])
def test_oracle_matches_float64_model(oracle, kv, opts):
    case = make_case(B=4, H=3, D=64, T=160, tile_size=16, seed=7, kv=kv, **opts)
    got = oracle_attention(case)
    np.testing.assert_allclose(got, numpy_attention(case), rtol=2e-4, atol=2e-5)


def test_oracle_double_temperature_switch(oracle):
    case = make_case(B=2, H=2, D=64, T=64, seed=3, temperature=2.0)
    a = oracle_attention(case)
    b = oracle_attention(case, double_temperature=True)
    case4 = dict(case, temperature=4.0)
    np.testing.assert_allclose(b, oracle_attention(case4), rtol=1e-5, atol=1e-6)
    assert np.abs(a - b).max() > 1e-3


def test_split_kv_combine_equals_global(oracle):
    """m/l/O partials over disjoint token ranges + LSE combine == global softmax."""
    c = oracle.cpu
    case = make_case(B=2, H=2, D=64, T=256, seed=11)
    full, probs, logits = oracle_attention(case, return_probs=True, return_logits=True)
    B, H, D = case["q"].shape
    dense_v = c.gather_pages(case["v_pool"].astype(np.float32), case["table"], case["num_beams"], H,
                             case["num_tiles"], case["tile_size"], D)
    parts = 4
    T = case["T"]
    pm = np.zeros((parts, B * H), np.float32); pl = np.zeros_like(pm)
    po = np.zeros((parts, B * H, D), np.float32)
    for i in range(parts):
        sl = slice(i * T // parts, (i + 1) * T // parts)
        s = logits[:, :, sl].reshape(B * H, -1)
        m = s.max(axis=1)
        e = np.exp(s - m[:, None])
        pm[i], pl[i] = m, e.sum(axis=1)
        po[i] = np.einsum("rt,rtd->rd", e, dense_v[:, :, sl].reshape(B * H, -1, D))
    out = c.lse_combine(pm, pl, po).reshape(B, H, D)
    np.testing.assert_allclose(out, full, rtol=1e-5, atol=1e-6)


def test_page_addressing(oracle):
    c = oracle.cpu
    case = make_case(B=3, H=2, D=64, T=64, seed=5, unmapped_frac=0.2)
    tb = case["table"]
    H, nt = case["H"], case["num_tiles"]
    assert c.pt_index(2, 1, 3, H, nt) == 2 * H * nt + nt + 3
    assert c.pt_lookup(tb, 5, 0, 0, H, nt) == -1 and c.pt_lookup(tb, -1, 0, 0, H, nt) == -1
    for (b, h, t) in [(0, 0, 0), (1, 1, 2), (2, 0, 3)]:
        pg = int(tb[b, h, t])
        off = c.kv_page_offset(tb, b, h, t, H, nt, case["total_pages"], 16, 64)
        assert off == (-1 if (pg < 0 or pg >= case["total_pages"]) else pg * 16 * 64)


def test_gemm_oracle(oracle):
    c = oracle.cpu
    rng = np.random.default_rng(0)
    A = rng.integers(-127, 128, size=(2, 5, 48), dtype=np.int8)
    B = rng.integers(-127, 128, size=(2, 48, 32), dtype=np.int8)
    acc = c.gemm_s8s8s32(A, B)
    np.testing.assert_array_equal(acc, np.einsum("bmk,bkn->bmn", A.astype(np.int64), B.astype(np.int64)))
    bias = rng.standard_normal(32).astype(np.float32)
    out = c.matmul_int8_epilogue(acc, 1 / 16, 1 / 16, 8.0, bias, "relu")
    ref = np.clip(np.rint(np.maximum(acc * np.float32(1 / 16 * 1 / 16 / 8.0) + bias, 0)), -128, 127)
    assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).max() <= 1
