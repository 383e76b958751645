import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pagedattention-based-transformer-decoder-inference-framework_b200"))
from llm_decoder import _cabi
os.environ["PA_LINEAR_TC"] = "1"
lib = _cabi.lib()
def run(x, W):
    rows, K = x.shape; N = W.shape[1]
    dx, dW = torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda()
    o = torch.full((rows, N), float("nan"), device="cuda")
    _cabi.check(lib.pa_linear_f32(dx.data_ptr(), dW.data_ptr(), None, rows, K, N, 0, o.data_ptr(), None, 0, None))
    torch.cuda.synchronize()
    return o.cpu().numpy()
rows, K, N = 128, 32, 128
x = np.ones((rows, K), np.float32); W = np.ones((K, N), np.float32)
o = run(x, W); print("ones:", o[0, :4], o[5, :4], o[127, 124:], "expect", K)
x = np.zeros((rows, K), np.float32); x[:, 0] = np.arange(rows)
o = run(x, W); print("row id:", o[:4, 0], o[60:64, 0], o[124:, 5])
x = np.ones((rows, K), np.float32); W = np.zeros((K, N), np.float32); W[0, :] = np.arange(N)
o = run(x, W); print("col id:", o[0, :6], o[0, 30:36], o[0, 124:])
x = np.zeros((rows, K), np.float32); x[:, 3] = 1; W = np.zeros((K, N), np.float32); W[3, :] = 2; W[4, :] = 100
o = run(x, W); print("k=3 match:", o[0, :4], "expect 2")
rng = np.random.default_rng(0)
x = rng.standard_normal((rows, K)).astype(np.float32); W = rng.standard_normal((K, N)).astype(np.float32)
o = run(x, W); e = x.astype(np.float64) @ W
print("rand K=32 maxerr", np.abs(o - e).max(), "rel", np.abs(o - e).max() / (np.abs(x) @ np.abs(W)).max())
rows, K, N = 64, 4096, 1024
x = rng.standard_normal((rows, K)).astype(np.float32); W = rng.standard_normal((K, N)).astype(np.float32)
o = run(x, W); e = x.astype(np.float64) @ W
print("rand big maxerr", np.abs(o - e).max(), "rel", (np.abs(o - e) / (np.abs(x).astype(np.float64) @ np.abs(W))).max())
rows, K, N = 128, 32, 128
x = (np.arange(rows)[:, None] * 100 + np.arange(K)[None, :]).astype(np.float32)
W = (np.arange(K)[:, None] * 1000 + np.arange(N)[None, :]).astype(np.float32)
for d in (1, 2, 3, 4):
    os.environ["PA_TC_DEBUG"] = str(d)
    o = run(x, W)
    print("debug", d, "row0", o[0, :12], "row1", o[1, :12], "row9", o[9, :8], "row 32:", o[32, :4], o[33, :4])
