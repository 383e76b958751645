#!/usr/bin/env python
"""In-kernel cycle sums of the tcgen05 fp32 linear kernel (pa_debug_linear_probe): per CTA, where the producer, the MMA
issuer and a split warp spend their time.  python benchmarks/linear_probe.py rows K N [packed] [streamk 0|1]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                "pagedattention-based-transformer-decoder-inference-framework_b200"))
from llm_decoder import _cabi  # noqa: E402

rows, K, N = (int(a) for a in sys.argv[1:4])
packed = len(sys.argv) > 4 and sys.argv[4] == "packed"
os.environ["PA_LINEAR_TC"] = "1"
lib = _cabi.lib()
raw = ctypes.CDLL(lib._name) if hasattr(lib, "_name") else lib
raw.pa_debug_linear_probe.argtypes = [ctypes.c_void_p]
raw.pa_debug_linear_probe.restype = None
x = torch.randn((rows, K), device="cuda")
W = torch.randn((K, N), device="cuda") / K ** 0.5
b = torch.randn((N,), device="cuda")
o = torch.empty((rows, N), device="cuda")
need = lib.pa_linear_workspace_bytes(rows, K, N)
ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
Wp = None
if packed:
    Wp = torch.empty(lib.pa_linear_pack_bytes(K, N) // 4, device="cuda")
    _cabi.check(lib.pa_linear_pack_f32(W.data_ptr(), Wp.data_ptr(), K, N, None))


def call():
    if packed:
        _cabi.check(lib.pa_linear_f32_packed(x.data_ptr(), Wp.data_ptr(), b.data_ptr(), rows, K, N, 1, o.data_ptr(), ws.data_ptr(), need, None))
    else:
        _cabi.check(lib.pa_linear_f32(x.data_ptr(), W.data_ptr(), b.data_ptr(), rows, K, N, 1, o.data_ptr(), ws.data_ptr(), need, None))


for _ in range(3):
    call()
torch.cuda.synchronize()
cnt = torch.zeros((4096, 8), dtype=torch.int64, device="cuda")
raw.pa_debug_linear_probe(ctypes.c_void_p(cnt.data_ptr()))
call()
torch.cuda.synchronize()
raw.pa_debug_linear_probe(None)
c = cnt.cpu().numpy()
c = c[c[:, 5] > 0]
blocks = c[:, 7].astype(np.float64)
names = ["producer_wait_empty", "mma_wait_split", "mma_issue", "split_wait_full", "split_work", "cta_total", "epilogue", "blocks"]
print(json.dumps({"ctas": int(len(c)), "blocks_per_cta_mean": float(blocks.mean()),
                  "per_block_cycles_mean": {n: round(float((c[:, i] / np.maximum(blocks, 1)).mean()), 1) for i, n in enumerate(names[:5])},
                  "cta_total_mean": float(c[:, 5].mean()), "cta_total_per_block": round(float((c[:, 5] / np.maximum(blocks, 1)).mean()), 1),
                  "epilogue_mean": float(c[:, 6].mean())}))
