"""Time the headline kernel only (CUDA graph over a ring of batches, like bench.py) -- for A/B library builds:
    LTR_B200_LIB=<path to a variant .so> python tools/time_approx.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from ltr_b200 import _lib, _runtime as rt  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 960
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
B, n, RING = 4096, 128, 48
data = [bench.make_batch(2020 + i) for i in range(RING)]
ds = [s.to(dev) for s, _, _ in data]
dy = [y.to(dev) for _, y, _ in data]
dg = [torch.empty(B, n, device=dev) for _ in range(RING)]
slate, loss = torch.empty(B, device=dev), torch.empty(1, device=dev)
tab = rt.disc_prefix(dev)
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    ws = rt.workspace(dev)
    st = rt.stream_ptr()

    def launch(k):
        rc = lib.ltr_approx_ndcg_fwdbwd(ds[k].data_ptr(), dy[k].data_ptr(), B, n, 10.0, 1e-10, -1.0, tab.data_ptr(),
                                        tab.numel(), slate.data_ptr(), loss.data_ptr(), dg[k].data_ptr(), None, 0,
                                        ws.data_ptr(), st)
        assert rc == 0, rc
    for k in range(RING):
        launch(k)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for k in range(RING):
            launch(k)
g.replay()
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps // RING):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / (steps // RING * RING) * 1e3)
print(f"{os.environ.get('LTR_B200_LIB', 'default')}: {best:.2f} us/launch  loss {float(loss):.6f}")
