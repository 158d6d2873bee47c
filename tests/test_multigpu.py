"""Query sharding on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`).
Every rank runs the kernels on its own contiguous query range; the risk losses all-reduce their phase states over NCCL.
The sharded result must equal the single-GPU result on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import PKG_ROOT, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, p2p, out_q):
    import sys
    for p in (ROOT, PKG_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from ltr_b200 import sharding
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    from ltr_b200.losses.lambdaL import lambdaLoss
    from ltr_b200.losses.riskLosses import riskLosses as RL
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LTR_RISK_P2P"] = "1" if p2p else "0"      # in-kernel exchange over NVLink peer memory / one all-gather per round
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = torch.Generator().manual_seed(11)
        B, n, S = 301, 128, 2
        s = torch.randn(B, n, generator=g)
        y = torch.randint(0, 5, (B, n), generator=g).float()
        base = torch.randn(B, n, S, generator=g)
        lo, hi = sharding.shard_bounds(B, rank, world)
        res = {}
        for name, fn, kw in (("geoListnet", RL.geoRiskListnetLoss, dict(listnet_transformation=1, return_strategy=1)),
                             ("zListnet_rs3", RL.zRiskListnetLoss, dict(listnet_transformation=3, return_strategy=3,
                                                                        add_ideal_ranking_to_mat=2)),
                             ("geoLambda", RL.geoRiskLambdaLoss, dict(listnet_transformation=1, return_strategy=2,
                                                                      add_ideal_ranking_to_mat=2))):
            sl = s[lo:hi].cuda().requires_grad_(True)
            loss = fn(sl, y[lo:hi].cuda(), base[lo:hi].cuda(), process_group=dist.group.WORLD, **kw)
            loss.sum().backward()
            res[name] = (loss.detach().cpu().numpy(), sl.grad.cpu().numpy())
        # per-slate losses: no collective on the data path, only scalar bookkeeping
        sl = s[lo:hi].cuda().requires_grad_(True)
        yl = y[lo:hi].cuda()
        la = approxNDCGLoss(sl, yl, alpha=10.0)
        res["approx_mean"] = float(sharding.global_mean(la, hi - lo, dist.group.WORLD))
        ll = lambdaLoss(sl, yl, weighing_scheme="ndcgLoss2PP_scheme")
        res["lambda_sum"] = float(sharding.global_sum(ll, dist.group.WORLD))
        lm = lambdaLoss(sl, yl, weighing_scheme="ndcgLoss2PP_scheme", reduction="mean", process_group=dist.group.WORLD)
        res["lambda_mean"] = float(sharding.global_sum(lm, dist.group.WORLD))
        ex = sharding.risk_peer_exchange(dist.group.WORLD, torch.device("cuda", rank))
        res["p2p_used"] = ex is not None
        res["status"] = ex.status() if ex is not None else 0
        out_q.put((rank, (lo, hi), res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("p2p", [True, False])
def test_query_sharded_losses_match_single_gpu(p2p):
    import torch.multiprocessing as mp
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    from ltr_b200.losses.lambdaL import lambdaLoss
    from ltr_b200.losses.riskLosses import riskLosses as RL
    from ltr_testing import assert_grad_close, assert_loss_close
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, p2p, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(11)
    B, n, S = 301, 128, 2
    s = torch.randn(B, n, generator=g)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    base = torch.randn(B, n, S, generator=g)
    single = {}
    for name, fn, kw in (("geoListnet", RL.geoRiskListnetLoss, dict(listnet_transformation=1, return_strategy=1)),
                         ("zListnet_rs3", RL.zRiskListnetLoss, dict(listnet_transformation=3, return_strategy=3,
                                                                    add_ideal_ranking_to_mat=2)),
                         ("geoLambda", RL.geoRiskLambdaLoss, dict(listnet_transformation=1, return_strategy=2,
                                                                  add_ideal_ranking_to_mat=2))):
        sd = s.cuda().requires_grad_(True)
        loss = fn(sd, y.cuda(), base.cuda(), **kw)
        loss.sum().backward()
        single[name] = (loss.detach().cpu().numpy(), sd.grad.cpu().numpy())
    for rank, (lo, hi), res in outs:
        for name, (loss1, grad1) in single.items():
            assert_loss_close(res[name][0], loss1, rtol=2e-6)
            assert_grad_close(res[name][1], grad1[lo:hi], rtol=1e-5)
        a = approxNDCGLoss(s.cuda(), y.cuda(), alpha=10.0)
        assert abs(res["approx_mean"] - float(a)) <= 1e-6 * abs(float(a))
        ls = lambdaLoss(s.cuda(), y.cuda(), weighing_scheme="ndcgLoss2PP_scheme")
        assert abs(res["lambda_sum"] - float(ls)) <= 1e-5 * abs(float(ls))
        lm = lambdaLoss(s.cuda(), y.cuda(), weighing_scheme="ndcgLoss2PP_scheme", reduction="mean")
        assert abs(res["lambda_mean"] - float(lm)) <= 1e-5 * abs(float(lm))     # global pair count (lambdaL.py:88-89)
        assert res["p2p_used"] == p2p and res["status"] == 0


def _grad_worker(rank, world, port, state, out_q):
    import sys
    for p in (ROOT, PKG_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from ltr_b200 import sharding
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        res = {}
        # (1) the raw collective: several segments, odd lengths, many rounds (both parities of the double-buffered inbox)
        ex = sharding.grad_peer_exchange(dist.group.WORLD, dev)
        res["p2p_used"] = ex is not None
        if ex is not None:
            rounds = []
            for it in range(7):
                g = torch.Generator().manual_seed(100 * it + rank)
                segs = [torch.randn(m, generator=g).to(dev) for m in (1, 8704, 3, 2049, 33)]
                ex.average(segs)
                rounds.append([t.cpu().numpy() for t in segs])
            big = torch.full((ex.max_floats,), float(rank + 1), device=dev)
            ex.average([big])
            res["rounds"], res["big"], res["status"] = rounds, big.cpu().numpy(), ex.status()
            with pytest.raises(Exception):
                ex.average([torch.zeros(ex.max_floats + 1, device=dev)])
        # (2) the fused step with process_group=: gradients come back averaged over the ranks
        B, n, F = 300, 128, 136
        g = torch.Generator().manual_seed(7 + rank)
        X = torch.rand(B, n, F, generator=g).to(dev)
        y = torch.randint(0, 5, (B, n), generator=g).float().to(dev)
        net = TripleLayerNet(F).to(dev)
        net.load_state_dict(state)
        for kind in ("approx", "listnet"):
            net.zero_grad(set_to_none=True)
            if kind == "approx":
                loss = net.approx_ndcg_loss(X, y, alpha=10.0, process_group=dist.group.WORLD)
            else:
                loss = net.listnet_loss(X, y, process_group=dist.group.WORLD)
            loss.backward()
            res[kind] = (float(loss), {k: v.grad.cpu().numpy() for k, v in net.named_parameters()})
        out_q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_peer_memory_gradient_average_and_data_parallel_fused_step():
    """`ltr_peer_allreduce_avg` (sum in rank order, x 1/world: bit-identical on every rank and to the same sum done on the
    host) and the fused ranker step with process_group= against two single-GPU steps averaged by hand."""
    import torch.multiprocessing as mp
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    world = 2
    torch.manual_seed(3)
    ref = TripleLayerNet(136).cuda()
    state = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, state, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=500) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0]["p2p_used"] and got[1]["p2p_used"], "two GPUs of one box map each other's memory"
    assert got[0]["status"] == 0 and got[1]["status"] == 0
    for it in range(7):
        gens = [torch.Generator().manual_seed(100 * it + r) for r in range(world)]
        for j, m in enumerate((1, 8704, 3, 2049, 33)):
            parts = [torch.randn(m, generator=gens[r]).numpy() for r in range(world)]
            want = (parts[0] + parts[1]) * np.float32(0.5)
            for r in range(world):
                assert np.array_equal(got[r]["rounds"][it][j], want), (it, j, r)
    assert np.all(got[0]["big"] == 1.5) and np.all(got[1]["big"] == 1.5)
    # the data-parallel step: every rank's own step, averaged here
    for kind in ("approx", "listnet"):
        own = []
        for r in range(world):
            g = torch.Generator().manual_seed(7 + r)
            X = torch.rand(300, 128, 136, generator=g).cuda()
            y = torch.randint(0, 5, (300, 128), generator=g).float().cuda()
            ref.zero_grad(set_to_none=True)
            loss = ref.approx_ndcg_loss(X, y, alpha=10.0) if kind == "approx" else ref.listnet_loss(X, y)
            loss.backward()
            own.append((float(loss), {k: v.grad.cpu().numpy() for k, v in ref.named_parameters()}))
        for r in range(world):
            assert got[r][kind][0] == own[r][0], "the loss stays the rank's own"
            for k in own[0][1]:
                want = (own[0][1][k] + own[1][1][k]) * np.float32(0.5)
                assert np.array_equal(got[r][kind][1][k], want), (kind, k, r)
