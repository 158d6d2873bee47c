"""Tiny driver for ncu: a handful of approxNDCG launches at the BASELINE configs[1] shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from ltr_b200 import _lib, _runtime as rt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda", 0)
lib = _lib.load()
s, y, _ = bench.make_batch(2020, B=B, n=n)
s, y = s.to(dev), y.to(dev)
grad = torch.empty_like(s)
slate = torch.empty(B, device=dev)
loss = torch.empty(1, device=dev)
tab, ws = rt.disc_prefix(dev), rt.workspace(dev)
for _ in range(reps):
    _lib.check(lib.ltr_approx_ndcg_fwdbwd(s.data_ptr(), y.data_ptr(), B, n, 10.0, 1e-10, -1.0, tab.data_ptr(), tab.numel(),
                                          slate.data_ptr(), loss.data_ptr(), grad.data_ptr(), None, 0, ws.data_ptr(),
                                          rt.stream_ptr()), "approx")
torch.cuda.synchronize()
print("loss", float(loss), "grad_abs_sum", float(grad.abs().sum()))
