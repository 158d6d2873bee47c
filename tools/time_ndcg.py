"""Time the NDCG@k sweep through the C-ABI only (CUDA graph over a ring of batches larger than L2):
    [LTR_B200_LIB=<variant .so>] python tools/time_ndcg.py [Q] [n] [k,k,...] [tie-stress 0/1]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch  # noqa: E402

from ltr_b200 import _lib, ops, _runtime as rt  # noqa: E402

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ks = [int(k) for k in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 5, 10]
ties = len(sys.argv) > 4 and sys.argv[4] == "1"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
RING = max(3, int(400e6 // (Q * n * 8)) + 1)
g = torch.Generator().manual_seed(7)
ys = [torch.randint(0, 5, (Q, n), generator=g).float().to(dev) for _ in range(RING)]
if ties:
    ss = [(torch.randint(0, 6, (Q, n), generator=g).float() / 3).to(dev) for _ in range(RING)]
else:
    ss = [torch.randn(Q, n, generator=g).to(dev) for _ in range(RING)]
d32, _ = ops._disc_tables(n, dev)
out = torch.empty(Q, len(ks), device=dev)
flags = torch.empty(Q + 1, dtype=torch.int32, device=dev)
karr = (ctypes.c_int * len(ks))(*ks)
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    st = rt.stream_ptr()

    def launch(i):
        rc = lib.ltr_ndcg_at_k(ys[i].data_ptr(), ss[i].data_ptr(), Q, n, karr, len(ks), ops.TIES_STDSORT, d32.data_ptr(),
                               out.data_ptr(), flags.data_ptr(), st)
        assert rc == 0, lib.ltr_last_error()
    for i in range(RING):
        launch(i)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for i in range(RING):
            launch(i)
graph.replay()
torch.cuda.synchronize()
reps = max(3, 300 // RING)
best = 1e9
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / (reps * RING) * 1e3)
marked = int((flags[1:] != 0).sum())
gbs = Q * (8 * n + 4 * len(ks)) / best / 1e3
print(f"{os.environ.get('LTR_B200_LIB', 'default')}: Q={Q} n={n} ks={ks} ties={ties}: {best:.2f} us/call, {Q / best:.1f} M q/s, "
      f"{gbs:.0f} GB/s algorithmic, {marked} queries marked for the full sort, mean {float(out.nan_to_num().mean()):.6f}")
