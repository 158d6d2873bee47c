import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200"), os.path.join(ROOT, "tests")]
import torch, numpy as np
from ltr_b200.architeture.tripleLayer import TripleLayerNet, _FusedTripleApproxNDCG
from ltr_b200 import _lib, _runtime as rt
from ltr_testing import make_slates
from oracle import ltr_oracle as O
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
net = TripleLayerNet(136)
g = torch.Generator().manual_seed(100)
X = torch.rand(B, 128, 136, generator=g)
_, y = make_slates(B, 128, seed=0, min_len=64)
W1, b1, W2, b2, W3, b3 = [p.detach() for p in (net.l1.weight, net.l1.bias, net.l2.weight, net.l2.bias, net.l3.weight, net.l3.bias)]
Wc = (W2 @ W1).clone().requires_grad_(True); bc = (W2 @ b1 + b2).clone().requires_grad_(True)
w3 = W3.reshape(-1).clone().requires_grad_(True); b3r = b3.clone().requires_grad_(True)
z = X @ Wc.t() + bc; a = torch.sigmoid(z); s = a @ w3 + b3r
loss = O.approx_ndcg_loss(s, y, alpha=10.0); loss.backward()
lib = _lib.load(); dev = torch.device("cuda", 0)
Xd, yd = X.cuda(), y.cuda()
F = 136
scores = torch.empty(B, 128, device=dev); slate = torch.empty(B, device=dev)
out = torch.zeros(1 + 32 * F + 65, device=dev)
nscr = lib.ltr_triple_scratch_bytes(B, 128, F); scr = torch.zeros(nscr, dtype=torch.uint8, device=dev)
tab = rt.disc_prefix(dev)
Wcd, bcd, w3d = Wc.detach().cuda().contiguous(), bc.detach().cuda().contiguous(), w3.detach().cuda().contiguous()
rc = lib.ltr_triple_approx_ndcg_fwdbwd(Xd.data_ptr(), yd.data_ptr(), Wcd.data_ptr(), bcd.data_ptr(), w3d.data_ptr(), float(b3), B, 128, F, 10.0, 1e-10, -1.0,
     tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), out[0:1].data_ptr(), out[1:1+32*F].data_ptr(), out[1+32*F:33+32*F].data_ptr(),
     out[33+32*F:65+32*F].data_ptr(), out[65+32*F:].data_ptr(), scr.data_ptr(), nscr, rt.workspace(dev).data_ptr(), rt.stream_ptr())
torch.cuda.synchronize()
print("rc", rc, "loss", float(out[0]), float(loss))
gWc = out[1:1+32*F].view(32, F).cpu(); gbc = out[1+32*F:33+32*F].cpu(); gw3 = out[33+32*F:65+32*F].cpu(); gb3 = out[65+32*F:].cpu()
def rel(a, b): return float((a - b).abs().max() / b.abs().max())
print("gWc", rel(gWc, Wc.grad), "gbc", rel(gbc, bc.grad), "gw3", rel(gw3, w3.grad), "gb3", float(gb3), float(b3r.grad))
part = scr.view(torch.float32)[: 64 * 160].view(64, 160).cpu()
exp = torch.zeros(64, 160); exp[:32, :136] = Wc.grad; exp[:32, 136] = bc.grad; exp[32:, 136] = w3.grad
np.set_printoptions(precision=3, suppress=False, linewidth=200)
print("got  rows 0,1,16,32 cols 0..7, 32..35, 128..139"); 
for r in (0, 1, 16, 32):
    print(r, part[r, :8].numpy(), part[r, 32:36].numpy(), part[r, 128:140].numpy())
print("exp")
for r in (0, 1, 16, 32):
    print(r, exp[r, :8].numpy(), exp[r, 32:36].numpy(), exp[r, 128:140].numpy())
# where does exp[0,0] appear in part?
for (r, c) in ((0, 0), (1, 0), (0, 1), (5, 77)):
    d = (part - exp[r, c]).abs(); i = int(d.argmin()); print("exp", (r, c), float(exp[r, c]), "closest in got at", divmod(i, 160), float(part.view(-1)[i]))
if os.environ.get("LTR_B200_LIB"):
    s.retain_grad() if False else None
    sg = torch.autograd.grad(O.approx_ndcg_loss((X @ Wc.detach().t() + bc.detach()).sigmoid() @ w3.detach() + b3, y, alpha=10.0).sum(), [], allow_unused=True) if False else None
    s2 = ((X @ Wc.detach().t() + bc.detach()).sigmoid() @ w3.detach() + b3).clone().requires_grad_(True)
    O.approx_ndcg_loss(s2, y, alpha=10.0).backward()
    print("ds got", scores[0, :6].cpu().numpy(), "exp", s2.grad[0, :6].numpy(), "rel", rel(scores.cpu(), s2.grad))
if "m64" in os.environ.get("LTR_B200_LIB", ""):
    Zp = (X[0, :64, :128] @ Wc.detach()[:, :128].t())
    print("m64 experiment: got rows 0,1,17,40 cols 0..5"); 
    for r in (0, 1, 17, 40): print(r, part[r, :6].numpy(), "exp", Zp[r, :6].numpy())
    print("rel", rel(part[:64, :32], Zp))
print("nonzero count in part", int((part != 0).sum()), "absmax", float(part.abs().max()))
