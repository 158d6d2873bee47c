"""Timeline of one CTA of the fused ranker kernel (RF_TRACE build: `build_variant("trace", ["RF_TRACE"])`, run with
LTR_B200_LIB pointing at it): globaltimer stamps of CTA 0's compute groups, forward issuer and gradient issuer per slate.
Prints, per slate, when each stage started / ended relative to the kernel's first stamp, and the mean duration of every
stage over the steady state.

    LTR_B200_LIB=.../variants/trace/libltr_b200.so python tools/trace_ranker.py
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from ltr_b200 import _lib, _runtime as rt
from ltr_testing import make_slates
lib = _lib.load(); dev = torch.device("cuda", 0); torch.cuda.set_device(0)
raw = ctypes.CDLL(_lib.LIB_PATH)
B, n, F, G = 4096, 128, 136, 3
X = torch.rand(B, n, F, device=dev)
y = make_slates(B, n, seed=1, min_len=64)[1].to(dev)
Wc = torch.randn(32, F, device=dev) * 0.1; bc = torch.randn(32, device=dev) * 0.1; w3 = torch.randn(32, device=dev)
scores = torch.empty(B, n, device=dev); slate = torch.empty(B, device=dev); out = torch.zeros(1 + 32 * F + 65, device=dev)
nscr = lib.ltr_triple_scratch_bytes(B, n, F); scr = torch.zeros(nscr, dtype=torch.uint8, device=dev)
tab = rt.disc_prefix(dev); ws = rt.workspace(dev); st = rt.stream_ptr()
trace = torch.zeros(6 * 8 * 64, dtype=torch.int64, device=dev)
def step():
    rc = lib.ltr_triple_approx_ndcg_fwdbwd(X.data_ptr(), y.data_ptr(), Wc.data_ptr(), bc.data_ptr(), w3.data_ptr(), 0.1, B, n, F, 10.0, 1e-10, -1.0,
         tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), out[0:1].data_ptr(), out[1:1+32*F].data_ptr(), out[1+32*F:33+32*F].data_ptr(),
         out[33+32*F:65+32*F].data_ptr(), out[65+32*F:].data_ptr(), scr.data_ptr(), nscr, ws.data_ptr(), st)
    assert rc == 0
for _ in range(3): step()
torch.cuda.synchronize()
raw.ltr_debug_ranker_trace.argtypes = [ctypes.c_void_p]
raw.ltr_debug_ranker_trace(ctypes.c_void_p(trace.data_ptr()))
step(); torch.cuda.synchronize()
raw.ltr_debug_ranker_trace(ctypes.c_void_p(0))
t = trace.cpu().numpy().reshape(6, 8, 64).astype(np.float64)
t0 = t[t > 0].min()
t = np.where(t > 0, (t - t0) / 1e3, np.nan)            # us
nsl = 28
print("slate grp | F: trig-wait  data-wait  done | group: zfull-wait  epilogue  loss  dz-wait  dz-write  bfly | D: dzfull-wait  q0 q1 q2 q3 done   (us since kernel start)")
for k in range(nsl):
    g = k % G
    f = t[4, :, k]; d = t[5, :, k]; c = t[g, :, k]
    print(f"{k:3d}  g{g}  | F {f[0]:6.1f} {f[1]:6.1f} {f[2]:6.1f} {f[3]:6.1f} | grp {c[0]:6.1f} {c[1]:6.1f} {c[2]:6.1f} {c[3]:6.1f} {c[4]:6.1f} {c[5]:6.1f} {c[6]:6.1f} | D {d[0]:6.1f} {d[1]:6.1f} {d[2]:6.1f} {d[3]:6.1f} {d[4]:6.1f} {d[5]:6.1f} {d[6]:6.1f}")
ks = np.arange(6, nsl - 3)
def mean(a): return float(np.nanmean(a))
print("steady state means (us):")
print(" group: wait Z %.2f | epilogue %.2f | loss %.2f | wait own gradient job %.2f | dZ write %.2f | butterfly %.2f | slate period %.2f" % (
    mean([t[k % G, 1, k] - t[k % G, 0, k] for k in ks]), mean([t[k % G, 2, k] - t[k % G, 1, k] for k in ks]),
    mean([t[k % G, 3, k] - t[k % G, 2, k] for k in ks]), mean([t[k % G, 4, k] - t[k % G, 3, k] for k in ks]),
    mean([t[k % G, 5, k] - t[k % G, 4, k] for k in ks]), mean([t[k % G, 6, k] - t[k % G, 5, k] for k in ks]),
    mean([t[k % G, 0, k + G] - t[k % G, 0, k] for k in ks])))
print(" F issuer: wait trigger %.2f | wait data %.2f | issue %.2f | job-to-job %.2f" % (
    mean([t[4, 1, k] - t[4, 0, k] for k in ks]), mean([t[4, 2, k] - t[4, 1, k] for k in ks]), mean([t[4, 3, k] - t[4, 2, k] for k in ks]),
    mean([t[4, 0, k + 1] - t[4, 0, k] for k in ks])))
print(" D issuer: wait dZ %.2f | first quarter's data %.2f | q1 %.2f q2 %.2f q3 %.2f | done %.2f | job-to-job %.2f" % (
    mean([t[5, 1, k] - t[5, 0, k] for k in ks]), mean([t[5, 2, k] - t[5, 1, k] for k in ks]), mean([t[5, 3, k] - t[5, 2, k] for k in ks]),
    mean([t[5, 4, k] - t[5, 3, k] for k in ks]), mean([t[5, 5, k] - t[5, 4, k] for k in ks]), mean([t[5, 6, k] - t[5, 5, k] for k in ks]),
    mean([t[5, 0, k + 1] - t[5, 0, k] for k in ks])))
print(" Z ready (F done) -> the group takes it: %.2f us before the group asks" % mean([t[k % G, 0, k] - t[4, 3, k] for k in ks]))
print(" dZ written -> gradient job done: %.2f us" % mean([t[5, 6, k] - t[k % G, 5, k] for k in ks]))
