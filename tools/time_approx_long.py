"""approxNDCG on long slates through the C-ABI (CUDA graph): time per call for several batch sizes, with and without
the gradient pass.    python tools/time_approx_long.py [n] [B,B,...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from ltr_b200 import _lib, _runtime as rt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Bs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 64, 512]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
tab = rt.disc_prefix(dev)
print(os.environ.get("LTR_B200_LIB", "default"))
for B in Bs:
    s, y, lens = bench.make_batch(900 + B, B=B, n=n)
    pairs = int((lens * (lens - 1)).sum())
    s, y = s.to(dev), y.to(dev)
    grad, slate, loss = torch.empty_like(s), torch.empty(B, device=dev), torch.empty(1, device=dev)
    nscr = lib.ltr_approx_ndcg_scratch_bytes(B, n)
    scr = torch.empty(max(nscr, 1), dtype=torch.uint8, device=dev)
    for want_grad in (True, False):
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            ws = rt.workspace(dev)
            st = rt.stream_ptr()

            def launch():
                rc = lib.ltr_approx_ndcg_fwdbwd(s.data_ptr(), y.data_ptr(), B, n, 10.0, 1e-10, -1.0, tab.data_ptr(), tab.numel(),
                                                slate.data_ptr(), loss.data_ptr(), grad.data_ptr() if want_grad else None,
                                                scr.data_ptr(), nscr, ws.data_ptr(), st)
                assert rc == 0, lib.ltr_last_error()
            launch()
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(10):
                    launch()
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 50 * 1e3)
        print(f"n={n} B={B:5d} grad={int(want_grad)}: {best:8.2f} us/call  {2 * pairs / best / 1e6:7.3f} T SFU-op/s  loss {float(loss):.6f}")
