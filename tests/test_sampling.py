"""Device sampling kernels against the oracle restatement of
attention_cpu/softmax_lut.cpp:203-256, which is itself pinned bit-exactly against the reference's
own object (tests/test_oracle_pinning.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ld(oracle):
    import llm_decoder
    return llm_decoder


def test_softmax_lut_bit_exact(ld, oracle):
    """pa_softmax_lut_i32 against the reference's OWN compiled softmax_batch_parallel / softmax_lut objects
    (oracle._ref) and the committed golden vectors: bit for bit, including the table built on the host."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
    lut = ld.sampling.build_exp_lut(1024, 10.0)
    np.testing.assert_array_equal(lut.numpy(), gold["lut"])
    dl = lut.cuda()
    for i in range(3):
        lg, sc = gold[f"sl_logits{i}"], float(gold[f"sl_scale{i}"])
        got = ld.sampling.softmax_lut(torch.from_numpy(lg[None]).cuda(), sc, dl).cpu().numpy()[0]
        np.testing.assert_array_equal(got, gold[f"sl_fused{i}"])
    rng = np.random.default_rng(12)
    for n, sc in ((8, 0.01), (512, 0.001), (4096, 0.05), (1000, 0.003)):
        x = rng.integers(-4000, 4000, size=(5, n), dtype=np.int32)
        got = ld.sampling.softmax_lut(torch.from_numpy(x).cuda(), sc, dl).cpu().numpy()
        for r in range(5):
            np.testing.assert_array_equal(got[r], oracle.cpu.fused_softmax_lut(x[r], sc, gold["lut"]))
        if oracle.ref.available() and n % 8 == 0:
            exp = oracle.ref.softmax_batch_parallel(x, sc, gold["lut"])
            np.testing.assert_array_equal(got, np.asarray(exp))


@pytest.mark.parametrize("V", [8, 1000, 50257])
def test_softmax_temperature_matches_oracle(ld, oracle, V):
    rng = np.random.default_rng(V)
    x = (rng.standard_normal((3, V)) * 4).astype(np.float32)
    for T in (1.0, 0.7, 2.5):
        got = ld.sampling.softmax_temperature(torch.from_numpy(x).cuda(), T).cpu().numpy()
        Vp = (V + 7) // 8 * 8                           # the reference's 8-wide Vec needs len % 8 == 0
        for r in range(3):
            xr = np.full(Vp, -1e9, np.float32)
            xr[:V] = x[r]
            exp = oracle.cpu.softmax_lut_vec(xr, T)[:V]
            np.testing.assert_allclose(got[r], exp, rtol=2e-5, atol=1e-9)


def _filter_case(ld, oracle, probs, top_k, top_p, eos=-1, thr=0.0):
    got = ld.sampling.apply_topk_topp_filter(torch.from_numpy(probs.copy()).cuda(), top_k, top_p, eos, thr).cpu().numpy()
    for r in range(probs.shape[0]):
        exp = oracle.cpu.apply_topk_topp_filter(probs[r], top_k, top_p, eos, thr)
        if np.array_equal(got[r], exp):
            continue
        # only entries sitting on the top-p boundary may differ (the cumulative sum is formed in a different
        # order on the device): their higher-ranked mass must be within rounding of top_p
        order = sorted(range(probs.shape[1]), key=lambda i: (probs[r, i], i), reverse=True)
        cum = np.concatenate([[0.0], np.cumsum(probs[r, order].astype(np.float64))[:-1]])
        mass = dict(zip(order, cum))
        assert top_p < 1.0
        for i in np.nonzero(got[r] != exp)[0]:
            assert abs(mass[int(i)] - top_p) < 1e-3, (r, int(i), mass[int(i)], top_p)  # the reference adds 50K f32 terms sequentially


def test_topk_topp_filter_matches_oracle(ld, oracle):
    rng = np.random.default_rng(5)
    for V in (16, 777, 50257):
        p = rng.random((4, V)).astype(np.float32) ** 6
        p[0, : V // 2] = p[0, V // 2: 2 * (V // 2)]      # exact ties: larger index ranks first
        p[1, :] = 1.0 / V                                # all equal
        p /= p.sum(axis=1, keepdims=True)
        p = p.astype(np.float32)
        for top_k, top_p in [(0, 1.0), (1, 1.0), (5, 1.0), (V, 1.0), (V + 3, 1.0), (0, 0.9), (0, 0.3), (0, 1e-9),
                             (7, 0.5), (3, 0.999), (0, 0.0)]:
            _filter_case(ld, oracle, p, top_k, top_p)
        eos = V // 3
        _filter_case(ld, oracle, p, 0, 1.0, eos, 0.0)            # EOS rule fires (prob > 0)
        _filter_case(ld, oracle, p, 0, 1.0, eos, 0.9)            # does not fire
        _filter_case(ld, oracle, p, 2, 1.0, int(np.argmax(p[2])), 1e-6)
    # SURVEY App. B known answers
    kv = np.array([[0.1, 0.4, 0.2, 0.3]], np.float32)
    for tk, tp in ((2, 1.0), (0, 0.6)):
        got = ld.sampling.apply_topk_topp_filter(torch.from_numpy(kv.copy()).cuda(), tk, tp).cpu().numpy()
        np.testing.assert_array_equal(got, np.array([[0, 0.4, 0, 0.3]], np.float32))


def test_sample_from_probs_inverse_cdf(ld):
    rng = np.random.default_rng(6)
    V = 50257
    p = rng.random((64, V)).astype(np.float32) ** 8
    p[:, rng.random(V) < 0.7] = 0.0                      # sparse, as after a top-k/top-p filter
    p[5] = 0.0
    p[5, 1234] = 0.25                                    # single survivor
    u = rng.random(64).astype(np.float32)
    u[0], u[1] = 0.0, np.float32(1.0 - 1e-7)
    ids = ld.sampling.sample_from_probs(torch.from_numpy(p).cuda(), torch.from_numpy(u).cuda()).cpu().numpy()
    c = np.cumsum(p.astype(np.float64), axis=1)
    for r in range(64):
        i = int(ids[r])
        assert p[r, i] > 0
        target = float(u[r]) * c[r, -1]
        before = c[r, i] - p[r, i]
        tol = 1e-5 * c[r, -1]
        assert before <= target + tol and c[r, i] >= target - tol, (r, i, before, target, c[r, i])
    assert ids[5] == 1234
    # end to end: logits -> ids, top_k = 1 must equal argmax whatever the draw
    x = torch.randn((8, 4096), device="cuda")
    s = ld.sampling.sample(x, temperature=0.8, top_k=1)
    assert torch.equal(s.long(), x.argmax(dim=1))
