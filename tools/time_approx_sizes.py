"""Launch-size sweep of the headline kernel: per-launch time vs number of slates (fixed cost vs per-slate cost)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from ltr_b200 import _lib, _runtime as rt  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
n = 128
tab = rt.disc_prefix(dev)
SIZES = [int(a) for a in sys.argv[1:]] or [1, 148, 592, 1184, 2368, 4096, 8192, 16384, 65536]
print(os.environ.get("LTR_B200_LIB", "default"))
for B in SIZES:
    RING = 8 if B <= 16384 else 2
    data = [bench.make_batch(2020 + i, B=B) for i in range(RING)]
    ds = [s.to(dev) for s, _, _ in data]
    dy = [y.to(dev) for _, y, _ in data]
    dg = [torch.empty(B, n, device=dev) for _ in range(RING)]
    slate, loss = torch.empty(B, device=dev), torch.empty(1, device=dev)
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
    reps = max(4, 2000 // RING if B <= 4096 else 40)
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / (reps * RING) * 1e3)
    print(f"B={B:6d}: {best:8.2f} us/launch  {best * 1e3 / B:8.2f} ns/slate")
