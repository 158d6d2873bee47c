"""The reference's training loop under DistributedDataParallel with the B200 losses (BASELINE configs 3 / 5 are
DDP runs of `main_batch_execution.py:120-171`): stock TripleLayerNet, query-sharded batch, `loss.backward()` through DDP's
gradient all-reduce.  Two ranks share cuda:0 over the gloo backend, so the test runs on a one-GPU box; what it checks is
the arithmetic of the integration, not the transport:
  * lambdaLoss(reduction="sum") is a SUM over queries while DDP AVERAGES gradients over ranks: the per-rank loss is scaled
    by the world size, and the DDP gradients then equal the single-process gradients of the whole batch;
  * approxNDCGLoss is a MEAN over the rank's queries: with equal shards DDP's average of the ranks' gradients is the
    whole-batch gradient as it stands;
  * lambdaLoss(reduction="mean") divides by the GLOBAL pair count (`process_group=`), then the same world-size scaling."""
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


def _make(seed=5, B=64, n=128, F=136):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(B, n, F, generator=g)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    lens = torch.randint(n // 2, n + 1, (B,), generator=g)
    y[torch.arange(n)[None, :] >= lens[:, None]] = -1.0
    return X, y


def _losses():
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    from ltr_b200.losses.lambdaL import lambdaLoss
    return {
        "lambda_sum": lambda s, y, world, group: world * lambdaLoss(s, y, weighing_scheme="ndcgLoss2PP_scheme"),
        "approx_mean": lambda s, y, world, group: approxNDCGLoss(s, y, alpha=10.0),
        "lambda_mean": lambda s, y, world, group: world * lambdaLoss(s, y, weighing_scheme="lamdbaRank_scheme", reduction="mean",
                                                                   **({"process_group": group} if group is not None else {})),
    }


def _worker(rank, world, port, state, out_q):
    import sys
    for p in (ROOT, PKG_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from ltr_b200 import sharding
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, y = _make()
        lo, hi = sharding.shard_bounds(X.shape[0], rank, world)
        net = TripleLayerNet(X.shape[2]).cuda()
        net.load_state_dict(state)
        ddp = DDP(net, device_ids=[0])
        res = {}
        for name, fn in _losses().items():
            ddp.zero_grad(set_to_none=True)
            scores = ddp(X[lo:hi].cuda(), None, None).squeeze(-1)
            loss = fn(scores, y[lo:hi].cuda(), world, dist.group.WORLD)
            loss.backward()
            res[name] = {k: v.grad.detach().cpu().numpy() for k, v in net.named_parameters()}   # (by value through the queue)
        out_q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_gradients_equal_the_single_process_whole_batch_gradients():
    import torch.multiprocessing as mp
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    world = 2
    torch.manual_seed(3)
    ref_net = TripleLayerNet(136).cuda()
    state = {k: v.detach().cpu() for k, v in ref_net.state_dict().items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, state, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, y = _make()
    for name, fn in _losses().items():
        ref_net.zero_grad(set_to_none=True)
        scores = ref_net(X.cuda(), None, None).squeeze(-1)
        fn(scores, y.cuda(), 1, None).backward()                  # whole batch, one process, no scaling
        for k, v in ref_net.named_parameters():
            want = v.grad.detach().cpu().numpy()
            # l3.bias: d loss / d b3 = sum of d loss / d score = 0 up to rounding (these losses are shift invariant):
            # measured against the scale of the neighbouring l3.weight gradient
            scale = float(np.abs(want).max()) if k != "l3.bias" else float(ref_net.l3.weight.grad.abs().max())
            for r in range(world):
                assert np.array_equal(got[r][name][k], got[0][name][k]), "DDP leaves every rank with the same gradients"
            err = float(np.abs(got[0][name][k] - want).max())
            assert err <= 2e-4 * scale + 1e-9, (name, k, err, scale)


def _fused_worker(rank, world, port, state, out_q):
    import sys
    for p in (ROOT, PKG_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from ltr_b200 import sharding
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LTR_GRAD_P2P"] = "0"        # two processes time-slice ONE GPU here: the torch.distributed fallback of the average
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, y = _make()
        lo, hi = sharding.shard_bounds(X.shape[0], rank, world)
        net = TripleLayerNet(X.shape[2]).cuda()
        net.load_state_dict(state)
        res = {}
        for kind in ("approx", "listnet"):
            net.zero_grad(set_to_none=True)
            if kind == "approx":
                loss = net.approx_ndcg_loss(X[lo:hi].cuda(), y[lo:hi].cuda(), alpha=10.0, process_group=dist.group.WORLD)
            else:
                loss = net.listnet_loss(X[lo:hi].cuda(), y[lo:hi].cuda(), process_group=dist.group.WORLD)
            loss.backward()
            res[kind] = {k: v.grad.detach().cpu().numpy() for k, v in net.named_parameters()}
        out_q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_fused_step_with_process_group_delivers_the_whole_batch_gradients():
    """`net.approx_ndcg_loss(..., process_group=)`: the plain module, gradients averaged over the ranks inside the step (here by
    torch.distributed; the NVLink peer-memory form needs two GPUs, tests/test_multigpu.py) == the one-process gradients of the
    whole batch (approxNDCG is a mean over slates and the shards are equal; ListNet is a sum, so its average is 1 / world of it)."""
    import torch.multiprocessing as mp
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    world = 2
    torch.manual_seed(4)
    ref_net = TripleLayerNet(136).cuda()
    state = {k: v.detach().cpu() for k, v in ref_net.state_dict().items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, world, port, state, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, y = _make()
    for kind in ("approx", "listnet"):
        ref_net.zero_grad(set_to_none=True)
        loss = ref_net.approx_ndcg_loss(X.cuda(), y.cuda(), alpha=10.0) if kind == "approx" else ref_net.listnet_loss(X.cuda(), y.cuda())
        loss.backward()
        for k, v in ref_net.named_parameters():
            want = v.grad.detach().cpu().numpy()
            scale = float(np.abs(want).max()) if k != "l3.bias" else float(ref_net.l3.weight.grad.abs().max())
            assert np.array_equal(got[0][kind][k], got[1][kind][k]), "every rank holds the same averaged gradients"
            # listnetLoss is a SUM over slates (losses/listnet.py:16): the ranks' average is 1 / world of the whole-batch gradient
            err = float(np.abs(got[0][kind][k] * (world if kind == "listnet" else 1) - want).max())
            assert err <= 1e-4 * scale + 1e-9, (kind, k, err, scale)
