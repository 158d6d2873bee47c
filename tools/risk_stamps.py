"""Phase timeline of the cooperative GeoRisk kernel (CTA 0's globaltimer stamps in the workspace)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch
from ltr_b200 import ops, _runtime as rt
torch.cuda.set_device(0)
B, n, S = 8192, 128, 2
g = torch.Generator().manual_seed(400)
s = torch.randn(B, n, generator=g).cuda().requires_grad_(True)
y = torch.randint(0, 5, (B, n), generator=g).float().cuda()
base = torch.randn(B, n, S, generator=g).cuda()
dev = s.device
for it in range(6):
    ops.risk_listnet(s, y, base, ops.KIND_GEO, 5.0, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 1, 1.0, False)
    torch.cuda.synchronize()
    ws = rt.workspace(dev)
    st = ws.view(torch.int64)[96:103].cpu().tolist()
    print("rows %.2f | rdv1 %.2f | statsB %.2f | rdv2 %.2f | consts %.2f | bwd %.2f | total %.2f us" % tuple(
        [(st[i + 1] - st[i]) / 1e3 for i in range(6)] + [(st[6] - st[0]) / 1e3]))
