"""N > 1 host logic on CPU: world_size-2 gloo process groups exercise the query partition and the risk losses'
two exchange rounds (ltr_b200/sharding.py) against the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_ROOT, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mat, alpha, kind, invert, strategy, out_q):
    import sys
    for p in (ROOT, PKG_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from ltr_b200 import sharding
    from oracle import risk_phases as RP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        group = dist.group.WORLD
        assert sharding.is_sharded(group)
        lo, hi = sharding.shard_bounds(mat.shape[0], rank, world)
        local = mat[lo:hi]
        s1 = torch.from_numpy(RP.state1(local))
        sharding.allreduce_risk_state1(s1, group)
        s2 = torch.from_numpy(RP.state2(local, s1.numpy(), invert, alpha, strategy))
        sharding.allreduce_risk_state2(s2, group)
        val = RP.final_value(s1.numpy(), s2.numpy(), kind, invert, strategy, 1.0, C=mat.shape[1])
        # scalar helpers
        mean = sharding.global_mean(torch.tensor(float(local[:, 0].mean())), hi - lo, group)
        total = sharding.global_sum(torch.tensor(float(local.sum())), group)
        out_q.put((rank, float(val), float(mean), float(total), (lo, hi)))
    finally:
        dist.destroy_process_group()


def _run(world, mat, alpha, kind, invert, strategy):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mat, alpha, kind, invert, strategy, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_shard_bounds_partition():
    from ltr_b200.sharding import shard_bounds
    for total in (0, 1, 7, 8, 4096, 8193):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("kind,invert,strategy", [(1, False, 1), (1, True, 1), (0, True, 2), (1, True, 3), (0, False, 1)])
def test_risk_exchange_rounds_match_unsharded_reference(kind, invert, strategy):
    """2 ranks, 37 queries x 4 systems: all-reduced phase states give the reference's value on the whole matrix."""
    from oracle import ltr_oracle as O
    rng = np.random.default_rng(7)
    raw = rng.random((37, 4)) * 0.5 + 0.05
    alpha = 3.0
    res = _run(2, raw, alpha, kind, invert, strategy)
    X = torch.from_numpy(raw.max() - raw if invert else raw).float()
    fn = O.geo_risk if kind == 1 else O.z_risk
    r0, rl = float(fn(X, alpha, i=0)), float(fn(X, alpha, i=-1))
    ref = r0 if strategy == 1 else (rl - r0 if strategy == 2 else (rl - r0) ** 2)
    for rank, val, mean, total, span in res:
        assert abs(val - ref) <= 2e-5 * max(1.0, abs(ref)), (rank, val, ref)
        assert abs(mean - raw[:, 0].mean()) < 1e-7         # returned in the dtype of the local mean (fp32)
        assert abs(total - raw.sum()) < 1e-4          # fp32 scalar round trip
    assert res[0][4] == (0, 19) and res[1][4] == (19, 37)
