"""What the TMA stream of feature tiles delivers on its own (ltr_probe_tma_stream): GB/s against copies in flight per SM,
tensor-map boxes (mode 0: [128 docs][32 floats], the fused ranker's forward copies; mode 2: [32 docs][5 x 32 floats] slots of 4 KB
boxes, its gradient copies) against contiguous 4 KB bulk copies (mode 1); from HBM
(4096 slates = 285 MB, three buffers cycled) and from L2 (296 slates = 20 MB read 14 times over)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch
from ltr_b200 import _lib, _runtime as rt
lib = _lib.load(); torch.cuda.set_device(0)
F = 136
for B, reps, what in ((4096, 1, "HBM"), (296, 14, "L2")):
    Xs = [torch.rand(B, 128, F, device="cuda") for _ in range(3 if reps == 1 else 1)]
    for mode in (0, 2, 1):
        for slots in (2, 3, 5, 8, 13) if mode != 2 else (2, 3, 4, 6, 10):
            def run(i):
                assert lib.ltr_probe_tma_stream(Xs[i % len(Xs)].data_ptr(), B, F, slots, mode, reps, rt.stream_ptr()) == 0
            for i in range(3): run(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(20): run(i)
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 20 * 1e3
            print(f"{what:3s} mode {mode} slots {slots:2d}: {us:7.1f} us  {reps*B*128*F*4/us/1e3:7.0f} GB/s  ({reps*B*128*F*4/us/1e3/148:5.1f} GB/s per SM)")
