"""Query sharding across the GPUs of one box.

Slates are independent, so every loss / metric kernel runs on a contiguous range of queries per rank with no
data-path collective.  The only exchange steps on the path are
  * the risk losses' matrix-wide statistics: two tiny all-reduces between the three phases of `ltr_risk_phase`
    (MAX of the matrix maximum, SUM of the column / row totals and of the z-risk partial sums), and
  * scalar bookkeeping when a global value is wanted (mean of per-rank means, global pair count of
    lambdaLoss(reduction="mean")).
Everything here is host-side plumbing over `torch.distributed` (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

# layout of the two 16-double state blocks of ltr_risk_phase (include/ltr_b200.h)
STATE1_MAX = slice(0, 1)      # max(mat)                                  -> MAX
STATE1_SUMS = slice(1, 5)     # sum mat[:,0], sum mat[:,C-1], sum(mat), Q -> SUM
STATE2_SUMS = slice(0, 13)    # z-risk partial sums, #maxima, t-risk sums -> SUM


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous, balanced range [lo, hi) of `total` queries owned by `rank` (first `total % world` ranks get one more)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def is_sharded(group) -> bool:
    return group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allreduce_risk_state1(state1: torch.Tensor, group) -> None:
    """In place: phase-1 state of every rank -> the global one."""
    dist.all_reduce(state1[STATE1_MAX], op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(state1[STATE1_SUMS], op=dist.ReduceOp.SUM, group=group)


def allreduce_risk_state2(state2: torch.Tensor, group) -> None:
    dist.all_reduce(state2[STATE2_SUMS], op=dist.ReduceOp.SUM, group=group)


def global_mean(local_mean: torch.Tensor, local_count: int, group) -> torch.Tensor:
    """Mean over all ranks' slates from per-rank means (approxNDCG / orderScore / exactNDCG divide by the LOCAL batch)."""
    buf = torch.stack([local_mean.detach().reshape(()).double() * local_count,
                       torch.tensor(float(local_count), dtype=torch.float64, device=local_mean.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[0] / buf[1]).to(local_mean.dtype)


def global_sum(local: torch.Tensor, group) -> torch.Tensor:
    """Sum over ranks of a per-rank sum (listnetLoss, lambdaLoss(reduction="sum"))."""
    buf = local.detach().clone()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf
