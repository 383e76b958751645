"""Decoder surface: CUDADecoder / INT8Decoder constructor, weight files,
generate() -- checked against the CPU restatement oracle/decoder_ref.py with teacher forcing
(token-level equality is only asserted where the oracle's own top-2 margin exceeds the numeric
noise), plus the small decoder kernels one by one."""
import os

import numpy as np
import pytest
import torch


def _quantize_ref_numpy():
    from llm_decoder.decoders import quantize_file_reference
    return quantize_file_reference


# ------------------------------------------------------------------ CPU: file-format arithmetic
def test_quantize_weights_arithmetic_matches_restatement(oracle):
    """int8_decoder.cpp:52-56: signed-max scale, truncation toward zero, no clamp (wraps)."""
    q = _quantize_ref_numpy()
    rng = np.random.default_rng(7)
    for w in (np.array([0.5, -1.0, 0.25, -0.26, 0.499], np.float32),
              rng.standard_normal(4097).astype(np.float32),
              (rng.standard_normal(1000) * 1e-3).astype(np.float32)):
        got, sc = q(w)
        exp, esc = oracle.cpu.quantize_weights_file(w)
        assert sc == esc == float(w.max())
        np.testing.assert_array_equal(got, exp)
    got, _ = q(np.array([0.5, -1.0], np.float32))
    np.testing.assert_array_equal(got, np.array([127, 2], np.int8))  # -254 wraps: no clamp in the reference


# ------------------------------------------------------------------ helpers
def make_weights(rng, L, hid, V, scale=1.0, attn=False):
    inter = 4 * hid
    w = {"embedding": (rng.standard_normal((V, hid)) * scale).astype(np.float32), "layers": []}
    for _ in range(L):
        extra = {}
        if attn:  # optional attention projections (weights/README.md:31-34)
            extra = {n: (rng.standard_normal((hid, hid)) / np.sqrt(hid)).astype(np.float32) for n in ("wq", "wk", "wv", "wo")}
        w["layers"].append(dict(
            **extra,
            ln1_g=(1 + 0.1 * rng.standard_normal(hid)).astype(np.float32),
            ln1_b=(0.1 * rng.standard_normal(hid)).astype(np.float32),
            ln2_g=(1 + 0.1 * rng.standard_normal(hid)).astype(np.float32),
            ln2_b=(0.1 * rng.standard_normal(hid)).astype(np.float32),
            fc1_w=(rng.standard_normal((hid, inter)) / np.sqrt(hid)).astype(np.float32),
            fc1_b=(0.1 * rng.standard_normal(inter)).astype(np.float32),
            fc2_w=(rng.standard_normal((inter, hid)) / np.sqrt(inter)).astype(np.float32),
            fc2_b=(0.1 * rng.standard_normal(hid)).astype(np.float32)))
    return w


def write_fp32_tree(w, path, packed_mlp):
    os.makedirs(path, exist_ok=True)
    w["embedding"].tofile(os.path.join(path, "embedding.bin"))
    for i, L in enumerate(w["layers"]):
        lp = os.path.join(path, f"layer_{i}")
        os.makedirs(lp, exist_ok=True)
        np.concatenate([L["ln1_g"], L["ln1_b"]]).tofile(os.path.join(lp, "ln1.bin"))
        np.concatenate([L["ln2_g"], L["ln2_b"]]).tofile(os.path.join(lp, "ln2.bin"))
        for n in ("wq", "wk", "wv", "wo"):
            if n in L:
                L[n].tofile(os.path.join(lp, f"attn_{n}.bin"))
        if packed_mlp:   # the order MLP::load_weights reads (mlp.hpp:17-20)
            np.concatenate([L["fc1_w"].ravel(), L["fc1_b"], L["fc2_w"].ravel(), L["fc2_b"]]).tofile(
                os.path.join(lp, "mlp.bin"))
        else:            # weights/README.md:31-39 / int8_decoder.cpp:66-70
            L["fc1_w"].tofile(os.path.join(lp, "mlp_fc1.bin"))
            L["fc2_w"].tofile(os.path.join(lp, "mlp_fc2.bin"))
            np.concatenate([L["fc1_b"], L["fc2_b"]]).tofile(os.path.join(lp, "mlp_biases.bin"))


def check_teacher_forced(seq, n_prompt, ref, temperature, divide, dec_logits_last, tol):
    """Feed `seq` to the CPU oracle; every generated token must be the oracle's argmax, or within
    `tol` of the oracle's maximum (numeric near-tie)."""
    from oracle.decoder_ref import sample
    logits = None
    for t, tok in enumerate(seq[:-1]):
        logits = ref.step(tok)
        if t + 1 >= n_prompt:
            v = logits / np.float32(temperature) if divide else logits * np.float32(temperature)
            got = seq[t + 1]
            if got != sample(logits, temperature, divide):
                assert v.max() - v[got] <= tol * max(1.0, np.abs(v).max()), (t, got, int(np.argmax(v)))
    return logits


# ------------------------------------------------------------------ GPU: kernels one by one
@pytest.mark.gpu
def test_decoder_kernels_match_oracle(oracle):
    import llm_decoder as ld
    from llm_decoder import _cabi
    lib, s = _cabi.lib(), None
    rng = np.random.default_rng(11)
    rows, hid, V = 5, 192, 301
    x = rng.standard_normal((rows, hid)).astype(np.float32) * 3 + 0.5
    g, b = rng.standard_normal(hid).astype(np.float32), rng.standard_normal(hid).astype(np.float32)
    dx, dg, db = (torch.from_numpy(a).cuda() for a in (x, g, b))
    out = torch.empty_like(dx)
    _cabi.check(lib.pa_layer_norm_f32(dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, hid, 1e-5, out.data_ptr(), s))
    np.testing.assert_allclose(out.cpu().numpy(), oracle.cpu.layer_norm(x, g, b), rtol=2e-5, atol=2e-5)
    # linear (mlp.hpp:23-41), rows > 8 exercises the row-chunk loop, N not a multiple of 32
    rows2, K, N = 11, 192, 100
    x2 = rng.standard_normal((rows2, K)).astype(np.float32)
    W = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    dx2, dW, dbias = (torch.from_numpy(a).cuda() for a in (x2, W, bias))
    for act in (0, 1):
        o = torch.empty((rows2, N), device="cuda")
        _cabi.check(lib.pa_linear_f32(dx2.data_ptr(), dW.data_ptr(), dbias.data_ptr(), rows2, K, N, act, o.data_ptr(),
                                      None, 0, s))
        exp = x2.astype(np.float64) @ W.astype(np.float64) + bias
        if act:
            exp = np.maximum(exp, 0)
        np.testing.assert_allclose(o.cpu().numpy(), exp, rtol=1e-4, atol=1e-4)
    # embedding f32 / i8 (+ out-of-range id -> zeros)
    E = rng.standard_normal((V, hid)).astype(np.float32)
    ids = np.array([0, 5, V - 1, V, -1], np.int32)
    dE, dids = torch.from_numpy(E).cuda(), torch.from_numpy(ids).cuda()
    o = torch.empty((5, hid), device="cuda")
    _cabi.check(lib.pa_embedding_f32(dE.data_ptr(), dids.data_ptr(), 5, hid, V, o.data_ptr(), s))
    exp = np.zeros((5, hid), np.float32)
    exp[:3] = E[ids[:3]]
    np.testing.assert_array_equal(o.cpu().numpy(), exp)
    Eq = rng.integers(-127, 128, (V, hid), dtype=np.int8)
    dEq = torch.from_numpy(Eq).cuda()
    _cabi.check(lib.pa_embedding_i8(dEq.data_ptr(), 42.5, dids.data_ptr(), 5, hid, V, o.data_ptr(), s))
    exp[:3] = oracle.cpu.dequantize_from_int8(Eq[ids[:3]].reshape(-1), 42.5).reshape(3, hid)
    np.testing.assert_array_equal(o.cpu().numpy(), exp)
    # logits + argmax (first maximum; divide / multiply forms)
    lg = torch.empty((rows, V), device="cuda")
    _cabi.check(lib.pa_logits_f32(dx.data_ptr(), dE.data_ptr(), rows, hid, V, lg.data_ptr(), s))
    np.testing.assert_allclose(lg.cpu().numpy(), x.astype(np.float64) @ E.T.astype(np.float64), rtol=1e-4, atol=1e-3)
    _cabi.check(lib.pa_logits_i8(dx.data_ptr(), dEq.data_ptr(), 42.5, rows, hid, V, lg.data_ptr(), s))
    np.testing.assert_allclose(lg.cpu().numpy(), (x.astype(np.float64) @ Eq.T.astype(np.float64)) / 42.5, rtol=1e-4, atol=1e-3)
    # fused logits + greedy sampler == logits then argmax (first maximum), f32 and int8 tables
    best = torch.zeros(rows, dtype=torch.int64, device="cuda")
    ids_f = torch.empty(rows, dtype=torch.int32, device="cuda")
    for (tab, eb, qs) in ((dE, 4, 1.0), (dEq, 1, 42.5)):
        for t, divide in ((0.7, 1), (1.3, 0)):
            _cabi.check(lib.pa_logits_argmax(dx.data_ptr(), tab.data_ptr(), eb, qs, rows, hid, V, t, divide,
                                             lg.data_ptr(), best.data_ptr(), ids_f.data_ptr(), s))
            lv = lg.cpu().numpy()
            v = lv / np.float32(t) if divide else lv * np.float32(t)
            np.testing.assert_array_equal(ids_f.cpu().numpy(), np.argmax(v, axis=1))
            assert int(best.abs().sum()) == 0     # scratch reset for the next call
    # fused per-row dynamic quantise == compute_minmax_scale + batch_quantize (bit-exact)
    xs = torch.empty(rows, device="cuda")
    xq = torch.empty((rows, hid), dtype=torch.int8, device="cuda")
    _cabi.check(lib.pa_row_quantize_dynamic_i8(dx.data_ptr(), rows, hid, xs.data_ptr(), xq.data_ptr(), s))
    sc = oracle.cpu.batch_minmax_scale(x, hid)
    np.testing.assert_array_equal(xs.cpu().numpy(), sc)
    np.testing.assert_array_equal(xq.cpu().numpy(), oracle.cpu.batch_quantize(x, sc, hid).reshape(rows, hid))
    # split-K path of the fp32 linear layer (few column strips, long K)
    x3 = rng.standard_normal((3, 4096)).astype(np.float32)
    W3 = rng.standard_normal((4096, 64)).astype(np.float32)
    o3 = torch.empty((3, 64), device="cuda")
    dx3, dW3 = torch.from_numpy(x3).cuda(), torch.from_numpy(W3).cuda()
    # K-sliced across CTAs (caller-owned scratch) and unsliced (no scratch): same result up to summation order
    need = lib.pa_linear_workspace_bytes(3, 4096, 64)
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    o3b = torch.empty((3, 64), device="cuda")
    _cabi.check(lib.pa_linear_f32(dx3.data_ptr(), dW3.data_ptr(), None, 3, 4096, 64, 1, o3b.data_ptr(), None, 0, s))
    _cabi.check(lib.pa_linear_f32(dx3.data_ptr(), dW3.data_ptr(), None, 3, 4096, 64, 1, o3.data_ptr(), ws.data_ptr(),
                                  need, s))
    np.testing.assert_allclose(o3b.cpu().numpy(), o3.cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(o3.cpu().numpy(), np.maximum(x3.astype(np.float64) @ W3.astype(np.float64), 0),
                               rtol=1e-4, atol=1e-3)
    l2 = rng.standard_normal((rows, V)).astype(np.float32)
    l2[0, 7] = l2[0, 200] = 9.0     # tie -> first index
    l2[1, :] = -np.inf
    dl2 = torch.from_numpy(l2).cuda()
    am = torch.empty(rows, dtype=torch.int32, device="cuda")
    for t, divide in ((0.7, 1), (2.0, 0), (-1.0, 0)):
        _cabi.check(lib.pa_argmax_f32(dl2.data_ptr(), rows, V, t, divide, am.data_ptr(), s))
        v = l2 / np.float32(t) if divide else l2 * np.float32(t)
        np.testing.assert_array_equal(am.cpu().numpy(), np.argmax(v, axis=1))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 3, 64, 256), (2, 70, 144, 272), (1, 16, 128, 8192)],
                         ids=lambda s: "x".join(map(str, s)))
def test_gemm_i8_dequant_matches_oracle(oracle, shape):
    """pa_gemm_i8_dequant: per-row dynamic scales, f32 output (direct and split-K epilogues)."""
    from llm_decoder import _cabi
    BATCH, M, N, K = shape
    rng = np.random.default_rng(sum(shape))
    A = rng.integers(-127, 128, (BATCH, M, K), dtype=np.int8)
    B = rng.integers(-127, 128, (BATCH, K, N), dtype=np.int8)
    bias = rng.standard_normal(N).astype(np.float32)
    qs = rng.uniform(5, 60, BATCH * M).astype(np.float32)
    acc = oracle.cpu.gemm_s8s8s32(A, B).astype(np.float32)
    dA, dB, dbias, dqs = (torch.from_numpy(a).cuda() for a in (A, B, bias, qs))
    need = _cabi.lib().pa_gemm_i8_workspace_bytes(BATCH, M, N, K)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    for act in ("", "relu", "gelu", "relu-nows"):
        C = torch.full((BATCH, M, N), float("nan"), device="cuda")
        wp, wb = (None, 0) if act == "relu-nows" else (ws.data_ptr(), need)  # without scratch: unsplit, same bits
        act = act.split("-")[0]
        _cabi.check(_cabi.lib().pa_gemm_i8_dequant(dA.data_ptr(), dB.data_ptr(), C.data_ptr(), BATCH, M, N, K,
                                                   dqs.data_ptr(), 0.013, dbias.data_ptr(), _cabi.ACT[act], wp, wb, None))
        alpha = (np.float32(0.013) / qs).reshape(BATCH, M, 1)
        exp = (alpha * acc).astype(np.float32) + bias
        if act == "relu":
            exp = np.maximum(exp, 0)
            np.testing.assert_array_equal(C.cpu().numpy(), exp)   # same fp32 op order: bit-exact
        elif act == "gelu":
            from math import erf
            exp = 0.5 * exp * (1 + np.vectorize(erf)(exp.astype(np.float64) * 0.7071067811865476))
            np.testing.assert_allclose(C.cpu().numpy(), exp, rtol=1e-5, atol=1e-5)
        else:
            np.testing.assert_array_equal(C.cpu().numpy(), exp)


# ------------------------------------------------------------------ GPU: CUDADecoder
@pytest.mark.gpu
@pytest.mark.parametrize("packed_mlp", [True, False])
def test_cuda_decoder_generate_vs_oracle(oracle, tmp_path, packed_mlp):
    import llm_decoder as ld
    from oracle.decoder_ref import RefDecoder
    rng = np.random.default_rng(31)
    L, H, D, V, S = 2, 2, 64, 211, 48
    hid = H * D
    w = make_weights(rng, L, hid, V)
    write_fp32_tree(w, str(tmp_path / "w"), packed_mlp)
    dec = ld.CUDADecoder(L, H, D, hid, V, S)            # bindings.cpp:5-6: six ints
    dec.load_weights(str(tmp_path / "w"))
    prompt = [3, 17, 101, 5]
    out = dec.generate(prompt, 12, 0.8)                  # bindings.cpp:8-15
    assert out[:4] == prompt and len(out) == 16 and all(0 <= t < V for t in out)
    ref = RefDecoder(w, H, D)
    check_teacher_forced(out, len(prompt), ref, 0.8, 1, None, 1e-4)
    # CUDA-graph replay == eager; batched prefill of the prompt == feeding it token by token
    dec2 = ld.CUDADecoder(L, H, D, hid, V, S, use_cuda_graph=False, use_overlap=False)
    dec2.load_weights(str(tmp_path / "w"))
    assert dec2.generate(prompt, 12, 0.8) == out
    dec3 = ld.CUDADecoder(L, H, D, hid, V, S, use_prefill=False)
    dec3.load_weights(str(tmp_path / "w"))
    out3 = dec3.generate(prompt, 12, 0.8)
    check_teacher_forced(out3, len(prompt), RefDecoder(w, H, D), 0.8, 1, None, 1e-4)
    assert out3 == out
    # tolerated caller variants (api/router.py:23, cli/chat_cli.py:24)
    lst = []
    assert dec.generate(prompt, lst, 5, 1.0) is lst and lst[:4] == prompt and len(lst) == 9
    assert dec.generate(prompt, max_gen_len=3) == out[:7] or len(dec.generate(prompt, max_gen_len=3)) == 7
    # logits of one eager step against the oracle
    dec.reset()
    ref2 = RefDecoder(w, H, D)
    for tok in prompt:
        lg = dec.forward_tokens([tok]).cpu().numpy()[0]
        np.testing.assert_allclose(lg, ref2.step(tok), rtol=2e-3, atol=2e-3)
    # batch of independent sequences == one at a time
    outs = dec.generate_batch([prompt, [9, 8, 7, 6]], 6, 0.8)
    assert outs[0] == out[:10]
    assert outs[1] == dec.generate([9, 8, 7, 6], 6, 0.8)


@pytest.mark.gpu
def test_cuda_decoder_head_dim_128_prefill_kernel(oracle, tmp_path):
    """head_dim 128 + fp16 pages: the prompt goes through the tensor-core prefill kernel; tokens must still be
    the oracle's (teacher forced), and equal the token-by-token path."""
    import llm_decoder as ld
    from oracle.decoder_ref import RefDecoder
    rng = np.random.default_rng(35)
    L, H, D, V, S = 2, 1, 128, 131, 96
    hid = H * D
    w = make_weights(rng, L, hid, V)
    write_fp32_tree(w, str(tmp_path / "w"), True)
    dec = ld.CUDADecoder(L, H, D, hid, V, S)
    dec.load_weights(str(tmp_path / "w"))
    prompt = [int(t) for t in rng.integers(0, V, 70)]          # two 64-query tiles, ragged second tile
    out = dec.generate(prompt, 8, 0.9)
    check_teacher_forced(out, len(prompt), RefDecoder(w, H, D), 0.9, 1, None, 2e-3)
    dec2 = ld.CUDADecoder(L, H, D, hid, V, S, use_prefill=False)
    dec2.load_weights(str(tmp_path / "w"))
    assert dec2.generate(prompt, 8, 0.9) == out


@pytest.mark.gpu
def test_int8_decoder_head_dim_128_prefill_kernel(oracle):
    """INT8Decoder with head_dim 128: the prompt goes through the int8 variant of the prefill kernel; the
    generated tokens equal the token-by-token path up to logit near-ties (checked against its own logits)."""
    import llm_decoder as ld
    L, H, D, V, S = 2, 1, 128, 127, 96
    hid = H * D
    g = torch.Generator(device="cuda").manual_seed(36)
    outs = []
    for use_prefill in (True, False):
        dec = ld.INT8Decoder(L, H, D, hid, V, S, use_prefill=use_prefill)
        g.manual_seed(36)
        dec.embedding.copy_(torch.randint(-127, 128, dec.embedding.shape, generator=g, device="cuda", dtype=torch.int8))
        for Ly in dec.layers:
            Ly.fc1_w.copy_(torch.randint(-127, 128, Ly.fc1_w.shape, generator=g, device="cuda", dtype=torch.int8))
            Ly.fc2_w.copy_(torch.randint(-127, 128, Ly.fc2_w.shape, generator=g, device="cuda", dtype=torch.int8))
            Ly.fc1_deq = Ly.fc2_deq = 0.05 / 127
        prompt = [int(t) for t in np.random.default_rng(36).integers(0, V, 70)]
        seq = dec.generate(prompt, 6, 1.0)
        outs.append((seq, dec.logits.clone()))
    (a, la), (b_, lb) = outs
    if a != b_:   # a near-tie may flip a token; then the two paths' last logits still agree closely up to that point
        first = next(i for i, (x, y) in enumerate(zip(a, b_)) if x != y)
        assert first >= 70
    else:
        np.testing.assert_allclose(la.cpu().numpy(), lb.cpu().numpy(), rtol=2e-2, atol=2e-2 * float(lb.abs().max()))


@pytest.mark.gpu
def test_decoder_sampling_options(tmp_path):
    """generate(..., top_k / top_p / seed): device sampling instead of greedy; top_k=1 == greedy; same seed
    reproduces; every sampled token lies in the top-k set of the teacher-forced oracle logits."""
    import llm_decoder as ld
    from oracle.decoder_ref import RefDecoder
    rng = np.random.default_rng(33)
    L, H, D, V, S = 1, 2, 64, 157, 40
    hid = H * D
    w = make_weights(rng, L, hid, V)
    write_fp32_tree(w, str(tmp_path / "w"), True)
    dec = ld.CUDADecoder(L, H, D, hid, V, S)
    dec.load_weights(str(tmp_path / "w"))
    prompt = [5, 9, 2]
    greedy = dec.generate(prompt, 8, 0.9)
    assert dec.generate(prompt, 8, 0.9, top_k=1) == greedy
    a = dec.generate(prompt, 8, 0.9, top_k=5, top_p=0.95, seed=7)
    assert a == dec.generate(prompt, 8, 0.9, top_k=5, top_p=0.95, seed=7)
    ref = RefDecoder(w, H, D)
    for t, tok in enumerate(a[:-1]):
        lg = ref.step(tok)
        if t + 1 >= len(prompt):
            top5 = set(np.argsort(-lg, kind="stable")[:6].tolist())   # 6: tolerate a numeric swap at rank 5/6
            assert a[t + 1] in top5


@pytest.mark.gpu
def test_decoder_errors():
    import llm_decoder as ld
    with pytest.raises(ValueError):
        ld.CUDADecoder(1, 2, 64, 100, 50, 32)            # hidden != H*D
    dec = ld.CUDADecoder(1, 2, 64, 128, 50, 32)
    with pytest.raises(RuntimeError):
        dec.load_weights("/nonexistent/path")            # "Cannot open file" -> RuntimeError via pybind
    with pytest.raises(ValueError):
        dec.generate([1] * 30, 10, 1.0)
    # zero-initialised weights (std::vector<T>(n)): every logit 0 -> argmax = token 0
    assert dec.generate([4, 5], 3, 1.0) == [4, 5, 0, 0, 0]


# ------------------------------------------------------------------ GPU: INT8Decoder
@pytest.mark.gpu
def test_int8_decoder_quantize_load_generate(oracle, tmp_path):
    import json
    import llm_decoder as ld
    from oracle.decoder_ref import RefDecoder
    rng = np.random.default_rng(32)
    L, H, D, V, S = 2, 2, 64, 203, 40
    hid, inter = H * D, 4 * H * D
    w = make_weights(rng, L, hid, V)
    fp32, int8 = str(tmp_path / "fp32"), str(tmp_path / "int8")
    write_fp32_tree(w, fp32, packed_mlp=False)
    dec = ld.INT8Decoder(L, H, D, hid, V, S)             # bindings.cpp:18-19
    dec.quantize_weights(fp32, int8)                     # bindings.cpp:21
    # files are byte-identical to the reference's arithmetic (restated in the oracle)
    for rel in ["embedding.bin"] + [f"layer_{i}/{f}" for i in range(L)
                                    for f in ("ln1.bin", "ln2.bin", "mlp_fc1.bin", "mlp_fc2.bin", "mlp_biases.bin")]:
        exp, _ = oracle.cpu.quantize_weights_file(np.fromfile(os.path.join(fp32, rel), np.float32))
        assert np.fromfile(os.path.join(int8, rel), np.int8).tobytes() == exp.tobytes(), rel
    dec.load_quantized_weights(int8)                     # bindings.cpp:20
    out = dec.generate([1, 2, 3], 10, 1.0)
    assert len(out) == 13 and out[:3] == [1, 2, 3]
    # oracle weights = what the decoder loaded (int8 payload + per-file scale)
    scales = json.load(open(os.path.join(int8, "quant_scales.json")))
    deq = lambda rel: np.float32(scales[rel] / 127.0)   # noqa: E731
    rd = lambda rel: np.fromfile(os.path.join(int8, rel), np.int8)  # noqa: E731
    wq = {"embedding": rd("embedding.bin").reshape(V, hid), "emb_qscale": 1.0 / (scales["embedding.bin"] / 127.0),
          "layers": []}
    for i in range(L):
        r = lambda f: f"layer_{i}/{f}"  # noqa: E731
        ln1 = rd(r("ln1.bin")).astype(np.float32) * deq(r("ln1.bin"))
        ln2 = rd(r("ln2.bin")).astype(np.float32) * deq(r("ln2.bin"))
        bb = rd(r("mlp_biases.bin")).astype(np.float32) * deq(r("mlp_biases.bin"))
        wq["layers"].append(dict(ln1_g=ln1[:hid], ln1_b=ln1[hid:], ln2_g=ln2[:hid], ln2_b=ln2[hid:],
                                 fc1_w=rd(r("mlp_fc1.bin")).reshape(hid, inter), fc1_deq=float(deq(r("mlp_fc1.bin"))),
                                 fc2_w=rd(r("mlp_fc2.bin")).reshape(inter, hid), fc2_deq=float(deq(r("mlp_fc2.bin"))),
                                 fc1_b=bb[:inter], fc2_b=bb[inter:]))
    ref = RefDecoder(wq, H, D, int8=True)
    check_teacher_forced(out, 3, ref, 1.0, 0, None, 2e-3)
    dec.reset()
    ref2 = RefDecoder(wq, H, D, int8=True)
    for tok in out[:6]:
        lg = dec.forward_tokens([tok]).cpu().numpy()[0]
        exp = ref2.step(tok)
        np.testing.assert_allclose(lg, exp, rtol=5e-3, atol=5e-3 * np.abs(exp).max())
    with pytest.raises(RuntimeError):
        dec.load_quantized_weights(str(tmp_path / "missing"))
    # batches of more than 8 rows take the tensor-core logits path (activations quantised like the MLP inputs):
    # every row is still the oracle's argmax up to the quantisation noise of its logits
    prompts = [[int(t) for t in rng.integers(0, V, 3)] for _ in range(11)]
    outs = dec.generate_batch(prompts, 6, 1.0)
    assert all(len(o) == 9 and o[:3] == p for o, p in zip(outs, prompts))
    for o in outs[:4]:
        check_teacher_forced(o, 3, RefDecoder(wq, H, D, int8=True), 1.0, 0, None, 3e-2)


# ------------------------------------------------------------------ GPU: fused quantisation around the INT8 MLP
@pytest.mark.gpu
def test_layer_norm_quantize_is_bit_identical_to_the_two_kernels(oracle):
    from llm_decoder import _cabi
    lib, s = _cabi.lib(), None
    rng = np.random.default_rng(51)
    for rows, hid in ((7, 192), (256, 4096), (3, 1000)):
        x = torch.from_numpy((rng.standard_normal((rows, hid)) * 3 + 0.3).astype(np.float32)).cuda()
        g = torch.from_numpy((1 + 0.1 * rng.standard_normal(hid)).astype(np.float32)).cuda()
        b = torch.from_numpy((0.1 * rng.standard_normal(hid)).astype(np.float32)).cuda()
        n = torch.empty_like(x)
        q0 = torch.empty((rows, hid), dtype=torch.int8, device="cuda")
        s0 = torch.empty(rows, device="cuda")
        _cabi.check(lib.pa_layer_norm_f32(x.data_ptr(), g.data_ptr(), b.data_ptr(), rows, hid, 1e-5, n.data_ptr(), s))
        _cabi.check(lib.pa_row_quantize_dynamic_i8(n.data_ptr(), rows, hid, s0.data_ptr(), q0.data_ptr(), s))
        n1 = torch.empty_like(x)
        q1, s1 = torch.empty_like(q0), torch.empty_like(s0)
        _cabi.check(lib.pa_layer_norm_quantize_i8(x.data_ptr(), g.data_ptr(), b.data_ptr(), rows, hid, 1e-5, n1.data_ptr(),
                                                  s1.data_ptr(), q1.data_ptr(), s))
        assert torch.equal(q0, q1) and torch.equal(s0, s1) and torch.equal(n, n1)
        q2, s2 = torch.empty_like(q0), torch.empty_like(s0)
        _cabi.check(lib.pa_layer_norm_quantize_i8(x.data_ptr(), g.data_ptr(), b.data_ptr(), rows, hid, 1e-5, None,
                                                  s2.data_ptr(), q2.data_ptr(), s))            # no f32 output
        assert torch.equal(q0, q2) and torch.equal(s0, s2)
        # and the pair is the oracle's: layer_norm (pinned to decoder/layer_norm.hpp) -> minmax scale -> batch_quantize
        ln = oracle.cpu.layer_norm(x.cpu().numpy(), g.cpu().numpy(), b.cpu().numpy(), 1e-5)
        np.testing.assert_allclose(n.cpu().numpy(), ln, rtol=2e-5, atol=2e-5)
        sc = oracle.cpu.batch_minmax_scale(n.cpu().numpy(), hid)
        np.testing.assert_array_equal(s0.cpu().numpy(), sc)
        np.testing.assert_array_equal(q0.cpu().numpy(), oracle.cpu.batch_quantize(n.cpu().numpy(), sc, hid))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(256, 512, 256), (160, 1024, 128), (256, 16384, 4096), (200, 272, 64)])
@pytest.mark.parametrize("act", ["relu", ""])
def test_gemm_dynquant_is_bit_identical_to_dequant_then_quantize(oracle, shape, act):
    """pa_gemm_i8_dynquant (row maxima cross the grid while the accumulators wait in TMEM) == pa_gemm_i8_dequant +
    pa_row_quantize_dynamic_i8, bit for bit: s8 values and per-row scales; and both follow the oracle's chain
    exact int32 accumulators -> alpha_row*acc + bias -> act -> compute_minmax_scale -> batch_quantize."""
    from llm_decoder import _cabi
    lib = _cabi.lib()
    M, N, K = shape
    rng = np.random.default_rng(52)
    A = rng.integers(-127, 128, (1, M, K), dtype=np.int8)
    B = rng.integers(-127, 128, (1, K, N), dtype=np.int8)
    bias = rng.standard_normal(N).astype(np.float32)
    qs = rng.uniform(5, 60, M).astype(np.float32)
    dA, dB, dbias, dqs = (torch.from_numpy(a).cuda() for a in (A, B, bias, qs))
    need = max(lib.pa_gemm_i8_workspace_bytes(1, M, N, K), 16)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    dq_ws = torch.zeros(lib.pa_gemm_i8_dynquant_workspace_bytes(1, M, N), dtype=torch.uint8, device="cuda")  # zeroed ONCE
    Cf = torch.empty((1, M, N), device="cuda")
    _cabi.check(lib.pa_gemm_i8_dequant(dA.data_ptr(), dB.data_ptr(), Cf.data_ptr(), 1, M, N, K, dqs.data_ptr(), 0.011,
                                       dbias.data_ptr(), _cabi.ACT[act], ws.data_ptr(), need, None))
    q0 = torch.empty((M, N), dtype=torch.int8, device="cuda")
    s0 = torch.empty(M, device="cuda")
    _cabi.check(lib.pa_row_quantize_dynamic_i8(Cf.data_ptr(), M, N, s0.data_ptr(), q0.data_ptr(), None))
    for _ in range(3):   # three times on the same workspace: the barrier counters reset themselves
        q1 = torch.full((1, M, N), 99, dtype=torch.int8, device="cuda")
        s1 = torch.zeros(M, device="cuda")
        _cabi.check(lib.pa_gemm_i8_dynquant(dA.data_ptr(), dB.data_ptr(), q1.data_ptr(), s1.data_ptr(), 1, M, N, K,
                                            dqs.data_ptr(), 0.011, dbias.data_ptr(), _cabi.ACT[act], dq_ws.data_ptr(),
                                            dq_ws.numel(), None))
        torch.cuda.synchronize()
        assert torch.equal(s0, s1)
        assert torch.equal(q0, q1[0])
    acc = oracle.cpu.gemm_s8s8s32(A, B)[0].astype(np.float32)
    v = ((np.float32(0.011) / qs)[:, None] * acc).astype(np.float32) + bias
    if act == "relu":
        v = np.maximum(v, 0)
    sc = oracle.cpu.batch_minmax_scale(v, N)
    np.testing.assert_array_equal(s0.cpu().numpy(), sc)
    np.testing.assert_array_equal(q0.cpu().numpy(), oracle.cpu.batch_quantize(v, sc, N))


@pytest.mark.gpu
def test_gemm_dynquant_rejects_shapes_that_do_not_fit_one_wave():
    from llm_decoder import _cabi
    lib = _cabi.lib()
    ws = torch.zeros(1 << 22, dtype=torch.uint8, device="cuda")
    for (M, N, K) in ((64, 512, 256),          # M <= 128: single-CTA kernel
                      (2048, 16384, 256)):      # 8 x 64 tiles: more CTA pairs than one wave
        A = torch.zeros((1, M, K), dtype=torch.int8, device="cuda")
        B = torch.zeros((1, K, N), dtype=torch.int8, device="cuda")
        q = torch.empty((1, M, N), dtype=torch.int8, device="cuda")
        sc = torch.empty(M, device="cuda")
        qs = torch.ones(M, device="cuda")
        st = lib.pa_gemm_i8_dynquant(A.data_ptr(), B.data_ptr(), q.data_ptr(), sc.data_ptr(), 1, M, N, K, qs.data_ptr(), 1.0,
                                     None, 0, ws.data_ptr(), ws.numel(), None)
        assert st == -2, (M, N, K, st)


@pytest.mark.gpu
def test_int8_decoder_fused_quantisation_same_tokens(monkeypatch):
    """INT8Decoder with > 128 rows: LN2 + quantise and fc1 + quantise fused into their producers give the SAME tokens
    and logits as the unfused kernels (the arithmetic is bit-identical)."""
    import llm_decoder as ld
    L, H, D, V, S = 2, 2, 64, 131, 24
    hid = H * D
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("PA_MLP_FUSED_QUANT", fused)
        dec = ld.INT8Decoder(L, H, D, hid, V, S)
        g = torch.Generator(device="cuda").manual_seed(53)
        dec.embedding.copy_(torch.randint(-127, 128, dec.embedding.shape, generator=g, device="cuda", dtype=torch.int8))
        for Ly in dec.layers:
            Ly.fc1_w.copy_(torch.randint(-127, 128, Ly.fc1_w.shape, generator=g, device="cuda", dtype=torch.int8))
            Ly.fc2_w.copy_(torch.randint(-127, 128, Ly.fc2_w.shape, generator=g, device="cuda", dtype=torch.int8))
            Ly.fc1_deq = Ly.fc2_deq = 0.05 / 127
        prompts = [[(7 * i + j) % V for j in range(3)] for i in range(160)]
        seqs = dec.generate_batch(prompts, 5, 1.0)
        outs.append((seqs, dec.logits.clone(), getattr(dec.bufs, "dq_ok", None)))
    assert outs[0][2] is True and outs[1][2] is False
    assert outs[0][0] == outs[1][0]
    assert torch.equal(outs[0][1], outs[1][1])


# ------------------------------------------------------------------ GPU: optional Q/K/V/O projections (8f row 1)
@pytest.mark.gpu
def test_cuda_decoder_with_attention_projections(oracle, tmp_path):
    """attn_wq / wk / wv / wo present in the layer directories (weights/README.md:31-34): q, k, v are projected before
    the cache append / attention and the attention output goes through Wo -- tokens and logits against the oracle."""
    import llm_decoder as ld
    from oracle.decoder_ref import RefDecoder
    rng = np.random.default_rng(37)
    L, H, D, V, S = 2, 2, 64, 151, 40
    hid = H * D
    w = make_weights(rng, L, hid, V, attn=True)
    write_fp32_tree(w, str(tmp_path / "w"), True)
    dec = ld.CUDADecoder(L, H, D, hid, V, S)
    dec.load_weights(str(tmp_path / "w"))
    assert dec.layers[0].wq is not None
    prompt = [3, 17, 101, 5, 9]
    out = dec.generate(prompt, 10, 0.8)
    check_teacher_forced(out, len(prompt), RefDecoder(w, H, D), 0.8, 1, None, 1e-4)
    dec2 = ld.CUDADecoder(L, H, D, hid, V, S, use_prefill=False, use_cuda_graph=False)
    dec2.load_weights(str(tmp_path / "w"))
    assert dec2.generate(prompt, 10, 0.8) == out
    dec.reset()
    ref = RefDecoder(w, H, D)
    for tok in prompt:
        np.testing.assert_allclose(dec.forward_tokens([tok]).cpu().numpy()[0], ref.step(tok), rtol=2e-3, atol=2e-3)
    # a tree WITHOUT the files keeps the reference block (q = k = v = LN1 output)
    w0 = {"embedding": w["embedding"], "layers": [{k: v for k, v in Ly.items() if not k.startswith("w")} for Ly in w["layers"]]}
    write_fp32_tree(w0, str(tmp_path / "w0"), True)
    dec3 = ld.CUDADecoder(L, H, D, hid, V, S)
    dec3.load_weights(str(tmp_path / "w0"))
    assert dec3.layers[0].wq is None
    check_teacher_forced(dec3.generate(prompt, 6, 0.8), len(prompt), RefDecoder(w0, H, D), 0.8, 1, None, 1e-4)


@pytest.mark.gpu
def test_int8_decoder_with_attention_projections(oracle, tmp_path):
    """The same through quantize_weights / load_quantized_weights: the four projection files are quantised like every
    other file and run on the tcgen05 int8 GEMM (int8_quant -> GEMM -> dequant)."""
    import json
    import llm_decoder as ld
    from oracle.decoder_ref import RefDecoder
    rng = np.random.default_rng(38)
    L, H, D, V, S = 2, 2, 64, 149, 40
    hid, inter = H * D, 4 * H * D
    w = make_weights(rng, L, hid, V, attn=True)
    fp32, int8 = str(tmp_path / "fp32"), str(tmp_path / "int8")
    write_fp32_tree(w, fp32, packed_mlp=False)
    dec = ld.INT8Decoder(L, H, D, hid, V, S)
    dec.quantize_weights(fp32, int8)
    for i in range(L):
        for n in ("wq", "wk", "wv", "wo"):
            exp, _ = oracle.cpu.quantize_weights_file(np.fromfile(os.path.join(fp32, f"layer_{i}/attn_{n}.bin"), np.float32))
            assert np.fromfile(os.path.join(int8, f"layer_{i}/attn_{n}.bin"), np.int8).tobytes() == exp.tobytes()
    dec.load_quantized_weights(int8)
    assert dec.layers[0].wq is not None and dec.layers[0].wq.dtype == torch.int8
    out = dec.generate([1, 2, 3], 8, 1.0)
    scales = json.load(open(os.path.join(int8, "quant_scales.json")))
    deq = lambda rel: np.float32(scales[rel] / 127.0)   # noqa: E731
    rd = lambda rel: np.fromfile(os.path.join(int8, rel), np.int8)  # noqa: E731
    wq = {"embedding": rd("embedding.bin").reshape(V, hid), "emb_qscale": 1.0 / (scales["embedding.bin"] / 127.0), "layers": []}
    for i in range(L):
        r = lambda f: f"layer_{i}/{f}"  # noqa: E731
        ln1 = rd(r("ln1.bin")).astype(np.float32) * deq(r("ln1.bin"))
        ln2 = rd(r("ln2.bin")).astype(np.float32) * deq(r("ln2.bin"))
        bb = rd(r("mlp_biases.bin")).astype(np.float32) * deq(r("mlp_biases.bin"))
        Ly = dict(ln1_g=ln1[:hid], ln1_b=ln1[hid:], ln2_g=ln2[:hid], ln2_b=ln2[hid:],
                  fc1_w=rd(r("mlp_fc1.bin")).reshape(hid, inter), fc1_deq=float(deq(r("mlp_fc1.bin"))),
                  fc2_w=rd(r("mlp_fc2.bin")).reshape(inter, hid), fc2_deq=float(deq(r("mlp_fc2.bin"))),
                  fc1_b=bb[:inter], fc2_b=bb[inter:])
        for n in ("wq", "wk", "wv", "wo"):
            Ly[n] = rd(r(f"attn_{n}.bin")).reshape(hid, hid)
            Ly[n + "_deq"] = float(deq(r(f"attn_{n}.bin")))
        wq["layers"].append(Ly)
    check_teacher_forced(out, 3, RefDecoder(wq, H, D, int8=True), 1.0, 0, None, 5e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 4096, 1024), (100, 192, 100), (16, 256, 128), (300, 520, 388), (65, 36, 12)])
def test_linear_register_tiled_gemm_matches_fp64(shape, monkeypatch):
    """pa_linear_f32 with >= 16 rows and 16-byte aligned rows runs the register-tiled fp32 GEMM (cp.async tiles, 8 x 4
    accumulators per thread): against a float64 product, K-sliced and unsliced, with bias / relu, ragged M / N / K tiles;
    and it equals the strip kernel (PA_LINEAR_GEMM=0) to fp32 summation-order noise."""
    from llm_decoder import _cabi
    lib = _cabi.lib()
    rows, K, N = shape
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal((rows, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dx, dW, db = (torch.from_numpy(a).cuda() for a in (x, W, b))
    exp = x.astype(np.float64) @ W.astype(np.float64) + b
    need = lib.pa_linear_workspace_bytes(rows, K, N)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    outs = {}
    monkeypatch.setenv("PA_LINEAR_TC", "0")   # the SIMT kernels are the subject here
    for mode in ("gemm", "gemm-nows", "strip"):
        monkeypatch.setenv("PA_LINEAR_GEMM", "0" if mode == "strip" else "1")
        for act in (0, 1):
            o = torch.full((rows, N), float("nan"), device="cuda")
            wp, wb = (None, 0) if mode == "gemm-nows" else (ws.data_ptr(), need)
            _cabi.check(lib.pa_linear_f32(dx.data_ptr(), dW.data_ptr(), db.data_ptr(), rows, K, N, act, o.data_ptr(), wp, wb, None))
            e = np.maximum(exp, 0) if act else exp
            np.testing.assert_allclose(o.cpu().numpy(), e, rtol=1e-4, atol=2e-4)
            outs[(mode, act)] = o.cpu().numpy()
    np.testing.assert_allclose(outs[("gemm", 0)], outs[("strip", 0)], rtol=1e-5, atol=1e-5)
    o = torch.empty((rows, N), device="cuda")
    _cabi.check(lib.pa_linear_f32(dx.data_ptr(), dW.data_ptr(), None, rows, K, N, 0, o.data_ptr(), ws.data_ptr(), need, None))
    np.testing.assert_allclose(o.cpu().numpy(), exp - b, rtol=1e-4, atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 4096, 1024), (100, 192, 100), (16, 256, 128), (300, 520, 388), (65, 36, 12),
                                   (128, 64, 260), (257, 2048, 128), (64, 2048, 1536), (200, 1024, 2048)])
@pytest.mark.parametrize("streamk", ["0", "1"])
def test_linear_tf32x3_tensor_core_kernel_matches_fp64(shape, streamk, monkeypatch):
    """pa_linear_f32 on tcgen05 kind::tf32 with the 3-term operand split (linear_tf32x3.cu, PA_LINEAR_TC=1 forces it
    for small shapes): against a float64 product with the stated bound 1e-5 * sum_k |x||W| + 4 fp32 ulps of the
    result (measured <= 3e-6: operand split 2^-21, plus the tensor core's truncating fp32 accumulation over K) -- K-sliced and unsliced, bias / relu, ragged M / N / K tiles (TMA zero fill), rows past the tile untouched;
    and within the same bound of the fp32 SIMT kernel (PA_LINEAR_TC=0)."""
    from llm_decoder import _cabi
    lib = _cabi.lib()
    rows, K, N = shape
    # both work splits: K-sliced grid, and stream-K (equal contiguous ranges of the (tile, K block) list per CTA: ranges
    # that start / end inside a tile, whole tiles finished in place, ragged last range)
    monkeypatch.setenv("PA_LINEAR_STREAMK", streamk)
    rng = np.random.default_rng(sum(shape) + 1)
    x = (rng.standard_normal((rows, K)) * np.exp(rng.uniform(-3, 3, (rows, 1)))).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dx, dW, db = (torch.from_numpy(a).cuda() for a in (x, W, b))
    exp = x.astype(np.float64) @ W.astype(np.float64) + b
    bound = 1e-5 * (np.abs(x).astype(np.float64) @ np.abs(W).astype(np.float64) + np.abs(b)) + 4 * np.spacing(np.abs(exp).astype(np.float32))
    need = lib.pa_linear_workspace_bytes(rows, K, N)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    outs = {}
    for mode in ("tc", "tc-nows", "simt"):
        monkeypatch.setenv("PA_LINEAR_TC", "0" if mode == "simt" else "1")
        for act in (0, 1):
            guard = 512
            buf = torch.full((rows * N + 2 * guard,), float("nan"), device="cuda")
            o = buf[guard:guard + rows * N].view(rows, N)
            wp, wb = (None, 0) if mode == "tc-nows" else (ws.data_ptr(), need)
            _cabi.check(lib.pa_linear_f32(dx.data_ptr(), dW.data_ptr(), db.data_ptr(), rows, K, N, act, o.data_ptr(), wp, wb, None))
            torch.cuda.synchronize()
            got = o.cpu().numpy().astype(np.float64)
            e = np.maximum(exp, 0) if act else exp
            assert np.all(np.abs(got - e) <= bound), float(np.max(np.abs(got - e) / bound))
            assert bool(torch.isnan(buf[:guard]).all()) and bool(torch.isnan(buf[-guard:]).all())
            outs[(mode, act)] = got
    assert np.all(np.abs(outs[("tc", 0)] - outs[("simt", 0)]) <= 2 * bound)
    # no bias
    o = torch.empty((rows, N), device="cuda")
    monkeypatch.setenv("PA_LINEAR_TC", "1")
    _cabi.check(lib.pa_linear_f32(dx.data_ptr(), dW.data_ptr(), None, rows, K, N, 0, o.data_ptr(), ws.data_ptr(), need, None))
    assert np.all(np.abs(o.cpu().numpy() - (exp - b)) <= bound)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 4096, 1024), (100, 192, 100), (1, 256, 128), (300, 520, 388), (65, 36, 13),
                                   (7, 64, 50257 // 64), (257, 2048, 128)])
def test_linear_packed_weights_same_bits_as_kn_layout(shape, monkeypatch):
    """pa_linear_f32_packed (weights repacked once into [feature tile][K block][32 k][128 n], every streamed block one
    contiguous 16 KB run) against pa_linear_f32 on the tensor-core kernel: the same arithmetic in the same order, so the
    bits must match wherever both apply; and against float64 with the kernel's stated bound for the shapes only the
    packed entry serves (fewer than 16 rows, N not a multiple of 4).  Pad columns / k of the packed copy are zero."""
    from llm_decoder import _cabi
    lib = _cabi.lib()
    rows, K, N = shape
    rng = np.random.default_rng(sum(shape) + 2)
    x = rng.standard_normal((rows, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N + 3).astype(np.float32)
    dx, dW = torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda()
    db = torch.from_numpy(b).cuda()[:N] if N % 4 == 0 else torch.from_numpy(b).cuda()[1:N + 1]   # misaligned bias too
    bb = db.cpu().numpy()
    Wp = torch.full((lib.pa_linear_pack_bytes(K, N) // 4,), float("nan"), device="cuda")
    _cabi.check(lib.pa_linear_pack_f32(dW.data_ptr(), Wp.data_ptr(), K, N, None))
    torch.cuda.synchronize()
    pk = Wp.view(-1, (K + 31) // 32, 32, 128).cpu().numpy()
    full = pk.transpose(1, 2, 0, 3).reshape(pk.shape[1] * 32, -1)      # [K padded, N padded]
    assert np.array_equal(full[:K, :N], W) and not full[K:].any() and not full[:, N:].any()
    need = lib.pa_linear_workspace_bytes(rows, K, N)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    exp = x.astype(np.float64) @ W.astype(np.float64) + bb
    bound = 1e-5 * (np.abs(x).astype(np.float64) @ np.abs(W).astype(np.float64) + np.abs(bb)) + 4 * np.spacing(np.abs(exp).astype(np.float32))
    for act in (0, 1):
        guard = 512
        buf = torch.full((rows * N + 2 * guard,), float("nan"), device="cuda")
        o = buf[guard:guard + rows * N].view(rows, N)
        _cabi.check(lib.pa_linear_f32_packed(dx.data_ptr(), Wp.data_ptr(), db.data_ptr(), rows, K, N, act, o.data_ptr(),
                                             ws.data_ptr(), need, None))
        torch.cuda.synchronize()
        e = np.maximum(exp, 0) if act else exp
        assert np.all(np.abs(o.cpu().numpy() - e) <= bound)
        assert bool(torch.isnan(buf[:guard]).all()) and bool(torch.isnan(buf[-guard:]).all())
        if rows >= 16 and N % 4 == 0:
            monkeypatch.setenv("PA_LINEAR_TC", "1")
            o2 = torch.empty((rows, N), device="cuda")
            _cabi.check(lib.pa_linear_f32(dx.data_ptr(), dW.data_ptr(), db.data_ptr(), rows, K, N, act, o2.data_ptr(),
                                          ws.data_ptr(), need, None))
            assert torch.equal(o, o2)


@pytest.mark.gpu
def test_cuda_decoder_packed_weights_same_logits_and_repacked_after_inplace_write(oracle, tmp_path, monkeypatch):
    """CUDADecoder packs its fp32 matrices once into the tensor-core kernel's order: a 24-row decode step must give the
    SAME logits (bitwise) as with PA_LINEAR_PACKED=0 (the kernel reads [K, N] rows), the packed copies must exist, and
    an in-place write to a weight tensor must be picked up (the copy is re-made: torch's version counter)."""
    import llm_decoder as ld
    rng = np.random.default_rng(41)
    L, H, D, V, S = 2, 2, 64, 132, 32
    hid = H * D
    w = make_weights(rng, L, hid, V, attn=True)
    write_fp32_tree(w, str(tmp_path / "w"), True)
    toks = [int(t) for t in rng.integers(0, V, 24)]
    monkeypatch.setenv("PA_LINEAR_TC", "1")           # these layers are below the size rule: force the tensor-core kernel

    def run(packed, mutate=False):
        monkeypatch.setenv("PA_LINEAR_PACKED", "1" if packed else "0")
        dec = ld.CUDADecoder(L, H, D, hid, V, S)
        dec.load_weights(str(tmp_path / "w"))
        dec.reset()
        a = dec.forward_tokens(toks, 0.7).clone()
        if mutate:
            dec.layers[0].fc1_w.mul_(0.5)
        b = dec.forward_tokens(toks, 0.7).clone()
        return a, b, dec

    a1, b1, dec1 = run(True, mutate=True)
    a0, b0, _ = run(False, mutate=True)
    assert len(dec1.__dict__.get("_pk", {})) >= 2 * L       # fc1, fc2 (+ projections) of every layer
    assert torch.equal(a1, a0)
    assert torch.equal(b1, b0) and not torch.equal(a1, b1)


@pytest.mark.gpu
@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("streamk", ["0", "1"])
def test_linear_chain_in_a_graph_is_deterministic(packed, streamk, monkeypatch):
    """fc1 -> fc2 -> fc1 ... chains of the tensor-core linear kernel (launched with programmatic stream serialisation: the
    next layer's weights start streaming under the previous kernel, x waits for griddepcontrol.wait) captured in ONE CUDA
    graph and replayed: every replay must reproduce, bit for bit, what the same calls give one at a time with a device
    synchronisation after each -- an ordering bug between a layer's sum kernel and the next layer's loads or partial
    tiles would show up as a difference."""
    from llm_decoder import _cabi
    lib = _cabi.lib()
    monkeypatch.setenv("PA_LINEAR_STREAMK", streamk)   # both work splits (K-sliced grid + sum kernel, stream-K + its sum kernel)
    M, HID, INTER, NL = 48, 1024, 2816, 6
    g = torch.Generator(device="cuda").manual_seed(5)
    W1 = [torch.randn((HID, INTER), device="cuda", generator=g) / HID ** 0.5 for _ in range(NL)]
    W2 = [torch.randn((INTER, HID), device="cuda", generator=g) / INTER ** 0.5 for _ in range(NL)]
    b1 = torch.randn(INTER, device="cuda", generator=g)
    b2 = torch.randn(HID, device="cuda", generator=g)
    if packed:
        def pack(W, K, N):
            Wp = torch.empty(lib.pa_linear_pack_bytes(K, N) // 4, device="cuda")
            _cabi.check(lib.pa_linear_pack_f32(W.data_ptr(), Wp.data_ptr(), K, N, None))
            return Wp
        W1 = [pack(w, HID, INTER) for w in W1]
        W2 = [pack(w, INTER, HID) for w in W2]
    x0 = torch.randn((M, HID), device="cuda", generator=g)
    xs = [torch.empty((M, HID), device="cuda") for _ in range(NL + 1)]
    hs = [torch.empty((M, INTER), device="cuda") for _ in range(NL)]
    need = max(lib.pa_linear_workspace_bytes(M, HID, INTER), lib.pa_linear_workspace_bytes(M, INTER, HID), 16)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    fn = lib.pa_linear_f32_packed if packed else lib.pa_linear_f32
    os.environ["PA_LINEAR_TC"] = "1"

    def chain(sync):
        for i in range(NL):
            _cabi.check(fn(xs[i].data_ptr(), W1[i].data_ptr(), b1.data_ptr(), M, HID, INTER, 1, hs[i].data_ptr(),
                           ws.data_ptr(), need, _cabi.stream()))
            if sync:
                torch.cuda.synchronize()
            _cabi.check(fn(hs[i].data_ptr(), W2[i].data_ptr(), b2.data_ptr(), M, INTER, HID, 0, xs[i + 1].data_ptr(),
                           ws.data_ptr(), need, _cabi.stream()))
            if sync:
                torch.cuda.synchronize()

    try:
        xs[0].copy_(x0)
        chain(True)
        ref = [t.clone() for t in xs[1:]] + [t.clone() for t in hs]
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            chain(False)   # warm (attribute set-up) outside capture
            side.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                chain(False)
        for rep in range(20):
            for t in xs[1:] + hs:
                t.fill_(float("nan"))
            gr.replay()
            torch.cuda.synchronize()
            got = xs[1:] + hs
            assert all(torch.equal(a, b) for a, b in zip(got, ref)), f"replay {rep}"
    finally:
        os.environ.pop("PA_LINEAR_TC", None)
