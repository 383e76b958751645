"""SURVEY 8(f) row 4 / 8(b) "caller-side variants": the reference's Python callers (api/, cli/, web/) against
the new `llm_decoder`, unmodified.

* tests/golden/ref_callsites.json holds the SHAPE of every decoder call site in those files (extracted with
  `ast` by tests/golden/extract_callsites.py: callee, positional kinds, keyword names, literal values).
* CPU: the fixture is current (when /root/reference is present), every call shape binds to the real signatures,
  the argument normalisation of `generate` maps each form to (prompt, out-list, max_len, temperature), the
  import-path shim resolves `from decoder.cuda_decoder import CUDADecoder`, and api/router.py itself is executed
  UNMODIFIED with only its out-of-scope collaborators stubbed (tokenizer, reranker) and the device loop of the
  decoder replaced -- the real `generate` front end handles the calls the routes make.
* GPU: every call shape is replayed on the real classes (the constructor with the callers' literal GPT-2-small
  dimensions; load_weights / generate on a small model).
"""
import importlib
import inspect
import json
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")
REF = "/root/reference"
SITES = json.load(open(os.path.join(HERE, "golden", "ref_callsites.json")))


def _value(spec, sentinels):
    if "const" in spec:
        return spec["const"]
    return sentinels[spec["kind"]]


def _generate_args(site, prompt, out_list, max_tokens, temperature):
    """Positional arguments of a `generate` call site: literals as written, variables by slot
    (prompt, out-parameter list, max tokens, temperature -- the C++ order, cuda_decoder.hpp:13)."""
    by_slot = [prompt, out_list, max_tokens, temperature]
    return [a["const"] if "const" in a else by_slot[i] for i, a in enumerate(site["args"])]


def test_callsite_fixture_is_current():
    if not os.path.isdir(REF):
        pytest.skip("no /root/reference")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    try:
        import extract_callsites
        out = {"imports": {}, "sites": []}
        for sub in ("api", "cli", "web"):
            for fn in sorted(os.listdir(os.path.join(REF, sub))):
                if fn.endswith(".py"):
                    imps, sites = extract_callsites.scan(os.path.join(REF, sub, fn), f"{sub}/{fn}")
                    if imps:
                        out["imports"][f"{sub}/{fn}"] = imps
                    out["sites"] += sites
    finally:
        sys.path.pop(0)
    assert json.loads(json.dumps(out, sort_keys=True)) == SITES


def test_import_path_shim_resolves_reference_imports():
    """`from decoder.cuda_decoder import CUDADecoder` (api/router.py:4) with <package>/compat on sys.path."""
    import llm_decoder as ld
    sys.path.insert(0, os.path.join(PKG, "compat"))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "decoder" or k.startswith("decoder.")}
    try:
        for rel, imps in SITES["imports"].items():
            for imp in imps:
                mod, name = imp.split(":")
                assert getattr(importlib.import_module(mod), name) is getattr(ld, name), (rel, imp)
    finally:
        sys.path.pop(0)
        for k in [k for k in sys.modules if k == "decoder" or k.startswith("decoder.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_every_call_shape_binds_and_normalises():
    import llm_decoder as ld
    out_list = []
    sent = {"Name": [5, 6, 7], "Attribute": 3, "out_list": out_list, "empty_list": []}
    n_gen = 0
    for s in SITES["sites"]:
        cls = ld.INT8Decoder if s["callee"] in ("INT8Decoder", "load_quantized_weights", "quantize_weights") else ld.CUDADecoder
        args = [_value(a, sent) for a in s["args"]]
        kwargs = {k: _value(v, sent) for k, v in s["kwargs"].items()}
        if s["callee"] == "generate":
            args = _generate_args(s, [5, 6, 7], out_list, 3, 0.7)
        if s["callee"] in ("CUDADecoder", "INT8Decoder"):
            b = inspect.signature(cls.__init__).bind(None, *args, **kwargs)
            got = [b.arguments[k] for k in ("num_layers", "num_heads", "head_dim", "hidden_dim", "vocab_size", "max_seq_len")]
            assert got == [12, 12, 64, 768, 50257, 2048], s          # bindings.cpp:6 argument order
        elif s["callee"] == "generate":
            # positional slot 2 is the prompt (a Name); slot 3 is always the C++ out-parameter list in the callers
            assert s["args"][1]["kind"] == "out_list", s
            inspect.signature(cls.generate).bind(None, *args, **kwargs)
            ids, ol, max_len, temp = cls.normalize_generate_args(*args, **kwargs)
            assert ids == [5, 6, 7] and ol is out_list
            want_len = kwargs.get("max_gen_len", args[2] if len(args) > 2 else None)
            want_t = kwargs.get("temperature", args[3] if len(args) > 3 else 1.0)
            assert max_len == want_len and temp == float(want_t), s
            n_gen += 1
        else:
            inspect.signature(getattr(cls, s["callee"])).bind(None, *args, **kwargs)
    assert n_gen >= 10


class _Recorder:
    """Stands in for the device loop only: constructor / load_weights record their arguments; `generate` is the
    REAL front end of llm_decoder (argument normalisation, out-parameter handling) over a fake generate_batch."""
    calls = []

    def __init__(self, num_layers, num_heads, head_dim, hidden_dim, vocab_size, max_seq_len):
        type(self).calls.append(("ctor", num_layers, num_heads, head_dim, hidden_dim, vocab_size, max_seq_len))

    def load_weights(self, path):
        type(self).calls.append(("load_weights", path))

    def generate_batch(self, prompts, max_len, temperature=1.0, **kw):
        type(self).calls.append(("generate_batch", [list(p) for p in prompts], max_len, temperature))
        return [list(p) + [100 + i for i in range(max_len)] for p in prompts]


def test_reference_api_router_runs_unmodified():
    """Execute /root/reference/api/router.py as shipped.  Stubbed: api.tokenizer (downloads GPT-2), reranker.reranker
    (out of scope), and the decoder's device loop (no GPU here)."""
    if not os.path.isdir(REF):
        pytest.skip("no /root/reference")
    pytest.importorskip("fastapi")
    from llm_decoder.decoders import _DecoderBase
    rec = type("CUDADecoder", (_Recorder,), {"generate": _DecoderBase.generate,
                                             "normalize_generate_args": staticmethod(_DecoderBase.normalize_generate_args),
                                             "calls": []})
    tok = types.ModuleType("api.tokenizer")

    class Tokenizer:
        def __init__(self, name="gpt2"):
            pass

        @classmethod
        def get(cls, name="gpt2"):
            return cls(name)

        def encode(self, text):
            return [ord(c) % 50 for c in text][:8] or [1]

        def decode(self, ids):
            return " ".join(str(i) for i in ids)
    tok.Tokenizer = Tokenizer
    rr = types.ModuleType("reranker.reranker")
    rr.Reranker = type("Reranker", (), {"__init__": lambda self, p: None, "select_best": lambda self, c, b: 0})
    dec_pkg, dec_mod = types.ModuleType("decoder"), types.ModuleType("decoder.cuda_decoder")
    dec_mod.CUDADecoder = rec
    stubs = {"api.tokenizer": tok, "reranker": types.ModuleType("reranker"), "reranker.reranker": rr,
             "decoder": dec_pkg, "decoder.cuda_decoder": dec_mod}
    saved = {k: sys.modules.get(k) for k in list(stubs) + ["api", "api.schema", "api.router"]}
    sys.modules.update(stubs)
    sys.path.insert(0, REF)
    for k in ("api", "api.schema", "api.router"):
        sys.modules.pop(k, None)
    try:
        router = importlib.import_module("api.router")
        assert rec.calls[0] == ("ctor", 12, 12, 64, 768, 50257, 2048)        # api/router.py:14 (keywords)
        assert rec.calls[1] == ("load_weights", "weights")                   # api/router.py:15
        from api.schema import GenerateRequest
        resp = router.generate(GenerateRequest(input_ids=[4, 5, 6], max_tokens=3, temperature=0.5))  # router.py:19-25
        assert resp.output_ids == [4, 5, 6, 100, 101, 102]                   # prompt + generated, via the out-list
        assert rec.calls[-1] == ("generate_batch", [[4, 5, 6]], 3, 0.5)
        # the streaming route re-feeds the growing context one token at a time (router.py:27-41)
        body = router.stream_generate(GenerateRequest(input_ids=[9], max_tokens=2, temperature=1.0)).body_iterator
        import asyncio

        async def drain():
            return [json.loads(x) async for x in body]
        chunks = asyncio.run(drain())
        assert [c["token"] for c in chunks] == [100, 100, None]
        assert rec.calls[-1] == ("generate_batch", [[9, 100]], 1, 1.0)
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ------------------------------------------------------------------------------------------------ GPU
def _tiny_tree(tmp_path, L, hid, V, rng):
    from test_decoders import make_weights, write_fp32_tree
    w = make_weights(rng, L, hid, V)
    write_fp32_tree(w, str(tmp_path / "weights"), packed_mlp=False)
    return w


@pytest.mark.gpu
def test_reference_call_shapes_on_the_real_decoders(tmp_path, monkeypatch):
    """Replay every call site of the reference's callers on the real classes.  The constructor runs with the
    callers' own literal dimensions (GPT-2 small, max_seq_len 2048); load_weights / generate run on a small model
    of the same family (the callers' weight tree does not exist here) from the callers' relative path "weights"."""
    import llm_decoder as ld
    sys.path.insert(0, HERE)
    rng = np.random.default_rng(41)
    L, H, D, V, S = 2, 2, 64, 97, 96
    hid = H * D
    _tiny_tree(tmp_path, L, hid, V, rng)
    small = ld.CUDADecoder(num_layers=L, num_heads=H, head_dim=D, hidden_dim=hid, vocab_size=V, max_seq_len=S)
    monkeypatch.chdir(tmp_path)
    prompt = [3, 1, 4, 1, 5]
    sent_attr = 4      # req.max_tokens / args.max_tokens
    seen_ctor = set()
    baseline = None
    for s in SITES["sites"]:
        out_list = []
        sent = {"Name": list(prompt), "Attribute": sent_attr, "out_list": out_list, "empty_list": []}
        args = [_value(a, sent) for a in s["args"]]
        kwargs = {k: _value(v, sent) for k, v in s["kwargs"].items()}
        if s["callee"] == "generate":
            args = _generate_args(s, list(prompt), out_list, sent_attr, 0.9)
        if s["callee"] in ("CUDADecoder", "INT8Decoder"):
            key = (s["callee"], bool(kwargs))
            if key not in seen_ctor:          # positional and keyword forms, once each (allocates the 2048-token caches)
                seen_ctor.add(key)
                dec = getattr(ld, s["callee"])(*args, **kwargs)
                assert (dec.num_layers_, dec.num_heads_, dec.head_dim_, dec.hidden_dim_, dec.vocab_size_,
                        dec.max_seq_len_) == (12, 12, 64, 768, 50257, 2048)
                del dec
        elif s["callee"] == "load_weights":
            small.load_weights(*args, **kwargs)            # "weights", relative to the caller's cwd
        elif s["callee"] == "generate":
            ret = small.generate(*args, **kwargs)
            assert ret is out_list and out_list[:len(prompt)] == prompt
            n = kwargs.get("max_gen_len", args[2] if len(args) > 2 else None)
            assert len(out_list) == len(prompt) + n and all(0 <= t < V for t in out_list)
            if n == 4 and len(args) > 3:
                baseline = baseline or list(out_list)
                assert out_list == baseline                # same call, same tokens (greedy)
                assert out_list == small.generate(prompt, 4, 0.9)   # == the pybind form (bindings.cpp:8-15)
    assert ("CUDADecoder", True) in seen_ctor and ("CUDADecoder", False) in seen_ctor
