"""Race hunt: every kernel family on a fixed input, many times; any bit that changes between runs is reported."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from ltr_testing import make_slates  # noqa: E402

torch.cuda.set_device(0)
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 40


def repeat(name, fn):
    ref = None
    bad = 0
    for _ in range(REPS):
        out = [o.clone() for o in fn()]
        if ref is None:
            ref = out
        elif not all(torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)) for a, b in zip(ref, out)):
            bad += 1
    print(f"{name:40s} {'ok' if bad == 0 else f'{bad} of {REPS} runs differ'}")
    return bad


def approx(B, n, **kw):
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    s, y = make_slates(B, n, seed=1, min_len=n // 2, **kw)
    sd, yd = s.cuda(), y.cuda()

    def run():
        x = sd.clone().requires_grad_(True)
        loss = approxNDCGLoss(x, yd, alpha=10.0)
        loss.backward()
        return loss.detach(), x.grad
    return run


def lam(B, n):
    from ltr_b200.losses.lambdaL import lambdaLoss
    s, y = make_slates(B, n, seed=2, min_len=n // 2)
    sd, yd = s.cuda(), y.cuda()

    def run():
        x = sd.clone().requires_grad_(True)
        loss = lambdaLoss(x, yd, weighing_scheme="ndcgLoss2PP_scheme")
        loss.backward()
        return loss.detach(), x.grad
    return run


def ndcg(Q, n, **kw):
    from ltr_b200.utils.metrics import ndcg_sweep
    s, y = make_slates(Q, n, seed=3, **kw)
    sd, yd = s.cuda(), y.cuda()
    return lambda: (ndcg_sweep(yd, sd, ks=(1, 5, 10)),)


def risk(B, n, S):
    from ltr_b200.losses.riskLosses.riskLosses import geoRiskListnetLoss
    g = torch.Generator().manual_seed(4)
    s = torch.randn(B, n, generator=g).cuda()
    y = torch.randint(0, 5, (B, n), generator=g).float().cuda()
    base = torch.randn(B, n, S, generator=g).cuda()

    def run():
        x = s.clone().requires_grad_(True)
        loss = geoRiskListnetLoss(x, y, base)
        loss.backward()
        return loss.detach(), x.grad
    return run


def ranker(B, F=136, min_len=64):
    from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_approx_ndcg
    torch.manual_seed(B)
    net = TripleLayerNet(F).cuda()
    g = torch.Generator().manual_seed(B)
    X = torch.rand(B, 128, F, generator=g).cuda()
    y = make_slates(B, 128, seed=B, min_len=min_len)[1].cuda()

    def run():
        net.zero_grad(set_to_none=True)
        loss, scores = fused_triple_approx_ndcg(net, X, y, alpha=10.0, return_scores=True)
        loss.backward()
        return [loss.detach(), scores] + [p.grad for p in net.parameters()]
    return run


bad = 0
bad += repeat("fused ranker step 4096 x 128 x 136", ranker(4096))
bad += repeat("fused ranker step 1500 x 128 x 136, short slates", ranker(1500, min_len=1))
bad += repeat("fused ranker step 149 x 128 x 64", ranker(149, F=64))
bad += repeat("approxNDCG 4096 x 128 (sym)", approx(4096, 128))
bad += repeat("approxNDCG 20000 x 128 (warp per slate)", approx(20000, 128))
bad += repeat("approxNDCG 64 x 1024 (cluster)", approx(64, 1024))
bad += repeat("approxNDCG 300 x 400 (cluster / general)", approx(300, 400))
bad += repeat("approxNDCG 512 x 128 ties", approx(512, 128, ties=True))
bad += repeat("LambdaLoss 1024 x 256", lam(1024, 256))
bad += repeat("LambdaLoss 3000 x 100", lam(3000, 100))
bad += repeat("NDCG 65536 x 128", ndcg(65536, 128))
bad += repeat("NDCG 100000 x 128", ndcg(100000, 128))
bad += repeat("NDCG 8192 x 1024", ndcg(8192, 1024))
bad += repeat("NDCG 4096 x 128 ties", ndcg(4096, 128, ties=True))
bad += repeat("geoRiskListnet 8192 x 128 x 2", risk(8192, 128, 2))
sys.exit(1 if bad else 0)
