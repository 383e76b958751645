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
for d in (20, 0):
    os.environ["PA_TC_DEBUG"] = str(d)
    o = run(x, W)
    print("debug", d, "ones: nonzero count", np.count_nonzero(o), "unique", np.unique(o)[:10])
