"""N > 1 host logic on CPU: world_size-2 / -3 gloo process groups exercise the query partition and the risk losses'
two exchange rounds (ltr_b200/sharding.py: one all-gather per round, states combined in rank order) against the
unsharded oracle -- including a rank whose shard is EMPTY."""
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
        # the exchange block of this rank (fp64 words): inbox[round][rank][16]; one all-gather per round
        xchg = torch.zeros(1024, dtype=torch.float64)
        sharding.risk_inbox(xchg, 0)[rank] = torch.from_numpy(RP.state1(local))
        sharding.allgather_risk_round(xchg, 0, group)
        g1 = RP.combine(sharding.risk_inbox(xchg, 0)[:world].numpy(), 0)
        sharding.risk_inbox(xchg, 1)[rank] = torch.from_numpy(RP.state2(local, g1, invert, alpha, strategy))
        sharding.allgather_risk_round(xchg, 1, group)
        g2 = RP.combine(sharding.risk_inbox(xchg, 1)[:world].numpy(), 1)
        val = RP.final_value(g1, g2, kind, invert, strategy, 1.0, C=mat.shape[1])
        pair_mean = sharding.global_pair_mean(torch.tensor(float(local.sum())), torch.tensor(float(local.size)), group)
        # scalar helpers
        mean = sharding.global_mean(torch.tensor(float(local[:, 0].mean()) if hi > lo else 0.0), hi - lo, group)
        total = sharding.global_sum(torch.tensor(float(local.sum())), group)
        out_q.put((rank, float(val), float(mean), float(total), (lo, hi), float(pair_mean)))
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
    for rank, val, mean, total, span, pair_mean in res:
        assert abs(val - ref) <= 2e-5 * max(1.0, abs(ref)), (rank, val, ref)
        assert abs(pair_mean - raw.mean()) < 1e-6
        assert abs(mean - raw[:, 0].mean()) < 1e-7         # returned in the dtype of the local mean (fp32)
        assert abs(total - raw.sum()) < 1e-4          # fp32 scalar round trip
    assert res[0][4] == (0, 19) and res[1][4] == (19, 37)


def test_risk_exchange_with_an_empty_shard():
    """3 ranks, 2 queries: the last rank owns nothing, contributes the neutral state and still joins both rounds."""
    from oracle import ltr_oracle as O
    rng = np.random.default_rng(11)
    raw = rng.random((2, 3)) * 0.5 + 0.05
    res = _run(3, raw, 2.0, 1, True, 1)
    X = torch.from_numpy(raw.max() - raw).float()
    ref = float(O.geo_risk(X, 2.0, i=0))
    assert [r[4] for r in res] == [(0, 1), (1, 2), (2, 2)]
    for r in res:
        assert abs(r[1] - ref) <= 2e-5 * max(1.0, abs(ref)), (r[0], r[1], ref)
