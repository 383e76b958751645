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
np.set_printoptions(linewidth=200, precision=3, suppress=True)
rows, K, N = 128, 32, 128
x = np.ones((rows, K), np.float32); W = np.ones((K, N), np.float32)
o = run(x, W)
print("ones: nonzero count", np.count_nonzero(o), "nan", np.isnan(o).sum(), "unique", np.unique(o)[:10])
nz = np.argwhere(o != 0)
print("first nonzero idx", nz[:5], "rows with nz", np.unique(nz[:, 0])[:20], "cols with nz", np.unique(nz[:, 1])[:40])
x = np.zeros((rows, K), np.float32); x[:, 0] = np.arange(rows) + 1
o = run(x, W); print("row id: col0", o[:, 0][:40]); print(" row0", o[0, :40])
x = np.ones((rows, K), np.float32); W = np.zeros((K, N), np.float32); W[0, :] = np.arange(N) + 1
o = run(x, W); print("col id: row0", o[0, :]); print("   col0", o[:16, 0])
x = np.zeros((rows, K), np.float32); x[:, 0] = 1; W = np.zeros((K, N), np.float32); W[:, 0] = np.arange(K) + 1
o = run(x, W); print("k id (x k=0, W col0 = k+1): ", o[0, :8])
for kk in (1, 7, 8, 31):
    x = np.zeros((rows, K), np.float32); x[:, kk] = 1
    o = run(x, W); print(" x k=%d -> " % kk, o[0, :4], "expect", kk + 1)
