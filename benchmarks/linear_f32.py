#!/usr/bin/env python
"""pa_linear_f32 at the fp32 decoder's layer shapes: tcgen05 3xTF32 kernel against the fp32 SIMT kernels
(PA_LINEAR_TC=0), with the max error of each against a float64 product on sampled outputs.

    python benchmarks/linear_f32.py            # prints one JSON line
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                "pagedattention-based-transformer-decoder-inference-framework_b200"))
from llm_decoder import _cabi  # noqa: E402


def run(rows, K, N, act, mode, iters=20):
    os.environ["PA_LINEAR_TC"] = "1" if mode == "tc" else "0"
    lib = _cabi.lib()
    g = torch.Generator(device="cuda").manual_seed(rows + K + N)
    x = torch.randn((rows, K), device="cuda", generator=g)
    W = torch.randn((K, N), device="cuda", generator=g) / K ** 0.5
    b = torch.randn((N,), device="cuda", generator=g)
    o = torch.empty((rows, N), device="cuda")
    need = lib.pa_linear_workspace_bytes(rows, K, N)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if mode == "packed":
        Wp = torch.empty(lib.pa_linear_pack_bytes(K, N) // 4, device="cuda")
        _cabi.check(lib.pa_linear_pack_f32(W.data_ptr(), Wp.data_ptr(), K, N, _cabi.stream()))
        call = lambda: _cabi.check(lib.pa_linear_f32_packed(x.data_ptr(), Wp.data_ptr(), b.data_ptr(), rows, K, N, act,
                                                            o.data_ptr(), ws.data_ptr(), need, _cabi.stream()))
    else:
        call = lambda: _cabi.check(lib.pa_linear_f32(x.data_ptr(), W.data_ptr(), b.data_ptr(), rows, K, N, act, o.data_ptr(),
                                                     ws.data_ptr(), need, _cabi.stream()))
    for _ in range(3):
        call()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    ridx = torch.randint(0, rows, (8,), device="cuda", generator=g)
    exp = x[ridx].double() @ W.double() + b.double()
    if act:
        exp = exp.clamp_min(0)
    den = x[ridx].double().abs() @ W.double().abs() + b.double().abs()
    err = float(((o[ridx].double() - exp).abs() / den).max())
    return {"us": round(us, 1), "tflops": round(2.0 * rows * K * N / us / 1e6, 1),
            "weights_gbs": round(K * N * 4 / us / 1e3, 1), "max_err_over_abs_products": err}


if __name__ == "__main__":
    if len(sys.argv) > 1:   # linear_f32.py rows K N act: one shape on the tensor-core kernel (for ncu captures)
        rows, K, N, act = (int(a) for a in sys.argv[1:5])
        print(json.dumps(run(rows, K, N, act, sys.argv[5] if len(sys.argv) > 5 else "tc", iters=3)))
        sys.exit(0)
    shapes = [("C2 fc1, batch 64", 64, 4096, 11008, 1), ("C2 fc2, batch 64", 64, 11008, 4096, 0),
              ("C2 projection, batch 64", 64, 4096, 4096, 0), ("C2 fc1, batch 256", 256, 4096, 11008, 1),
              ("prefill fc1, 2048 tokens", 2048, 4096, 11008, 1), ("prefill fc2, 2048 tokens", 2048, 11008, 4096, 0),
              ("C1 fc1, batch 64", 64, 768, 3072, 1), ("C1 fc1 prefill 64x448", 28672, 768, 3072, 1)]
    res = {}
    for name, rows, K, N, act in shapes:
        res[name] = {m: run(rows, K, N, act, m) for m in ("packed", "tc", "simt")}
    print(json.dumps({"workload": "pa_linear_f32 (fp32 x [rows,K] . W [K,N] + bias, relu on fc1)", "results": res}))
