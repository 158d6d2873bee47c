"""Device time of ltr_triple_approx_ndcg_fwdbwd alone (C-ABI call, CUDA events), BASELINE configs[1] shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200"), os.path.join(ROOT, "tests")]
import torch
from ltr_b200 import _lib, _runtime as rt
from ltr_testing import make_slates
lib = _lib.load(); dev = torch.device("cuda", 0); torch.cuda.set_device(0)
B, n, F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 128, 136
ring = 3
Xs = [torch.rand(B, n, F, device=dev) for _ in range(ring)]
ys = [make_slates(B, n, seed=i, min_len=64)[1].to(dev) for i in range(ring)]
Wc = torch.randn(32, F, device=dev) * 0.1; bc = torch.randn(32, device=dev) * 0.1; w3 = torch.randn(32, device=dev)
scores = torch.empty(B, n, device=dev); slate = torch.empty(B, device=dev); out = torch.zeros(1 + 32 * F + 65, device=dev)
nscr = lib.ltr_triple_scratch_bytes(B, n, F); scr = torch.zeros(nscr, dtype=torch.uint8, device=dev)
tab = rt.disc_prefix(dev); ws = rt.workspace(dev); st = rt.stream_ptr()
def step(i):
    rc = lib.ltr_triple_approx_ndcg_fwdbwd(Xs[i % ring].data_ptr(), ys[i % ring].data_ptr(), Wc.data_ptr(), bc.data_ptr(), w3.data_ptr(), 0.1, B, n, F, 10.0, 1e-10, -1.0,
         tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), out[0:1].data_ptr(), out[1:1+32*F].data_ptr(), out[1+32*F:33+32*F].data_ptr(),
         out[33+32*F:65+32*F].data_ptr(), out[65+32*F:].data_ptr(), scr.data_ptr(), nscr, ws.data_ptr(), st)
    assert rc == 0
for i in range(5): step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 30
a.record()
for i in range(iters): step(i)
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / iters * 1e3
print(f"B={B}: {us:.1f} us per call, {B / us:.2f} M slates/s, X bytes {B*n*F*4/1e6:.0f} MB -> {B*n*F*4/us/1e3:.0f} GB/s")
