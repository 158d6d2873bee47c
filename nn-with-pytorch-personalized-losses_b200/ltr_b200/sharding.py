"""Query sharding across the GPUs of one box.

Slates are independent, so every loss / metric kernel runs on a contiguous range of queries per rank with no
data-path collective.  The only exchange steps on the path are
  * the risk losses' matrix-wide statistics: two rounds of 16 doubles per rank (round 0: MAX of the matrix maximum,
    SUM of the column / row totals and the query count; round 1: SUM of the z-risk partial sums).  Either the
    kernels exchange them themselves over NVLink peer memory (`RiskPeerExchange`: every rank's exchange block is
    mapped into every process with CUDA IPC, the last CTA of a phase stores its state into all peers' inboxes and
    the next phase spins on sequence-numbered flags -- no collective call at all), or, without peer memory, ONE
    all-gather per round between the phase launches (`allgather_risk_round`: NCCL on the GPUs, gloo in the CPU tests);
  * scalar bookkeeping when a global value is wanted (mean of per-rank means, the global pair count of
    lambdaLoss(reduction="mean")).
Everything here is host-side plumbing over `torch.distributed` and the C-ABI.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib

RISK_STATE = 16                       # doubles per phase state (include/ltr_b200.h)
RISK_MAX_WORLD = _lib.RISK_MAX_WORLD  # ranks an exchange block has room for


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous, balanced range [lo, hi) of `total` queries owned by `rank` (first `total % world` ranks get one
    more).  With total < world the last ranks get an EMPTY range: the risk losses accept it (the rank contributes the
    neutral state to both exchange rounds), every other loss is simply not called for it."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def is_sharded(group) -> bool:
    return group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def rank_world(group):
    if not is_sharded(group):
        return 0, 1
    world = dist.get_world_size(group)
    if world > RISK_MAX_WORLD:
        raise ValueError(f"the risk exchange block has room for {RISK_MAX_WORLD} ranks, got {world}")
    return dist.get_rank(group), world


def risk_inbox(xchg: torch.Tensor, slot: int) -> torch.Tensor:
    """[RISK_MAX_WORLD, 16] view of round `slot` of an exchange block (fp64 words)."""
    n = RISK_MAX_WORLD * RISK_STATE
    return xchg[slot * n:(slot + 1) * n].view(RISK_MAX_WORLD, RISK_STATE)


def allgather_risk_round(xchg: torch.Tensor, slot: int, group) -> None:
    """In place: row `rank` of round `slot` of every rank's exchange block -> rows 0..world-1 of everybody's.
    One collective (the kernels of the next phase combine the rows themselves, MAX / SUM in rank order, so every
    rank gets bit-identical totals)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    inbox = risk_inbox(xchg, slot)
    mine = inbox[rank]
    if xchg.device.type != "cuda":
        mine = mine.clone()                       # (gloo copies the input into the output first: no aliasing there)
    dist.all_gather_into_tensor(inbox[:world].reshape(-1), mine, group=group)


class RiskPeerExchange:
    """Every rank's exchange block mapped into this process (CUDA IPC over NVLink / PCIe peer access), so that
    `ltr_risk_listnet_fwdbwd` exchanges the phase states inside its kernels.  Collective constructor: every rank of
    `group` must build it at the same point (the 64-byte IPC handles travel by all_gather_object, any backend)."""

    def __init__(self, group, device: torch.device):
        lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        self._lib = lib
        local = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.ltr_risk_xchg_alloc(ctypes.byref(local)), "ltr_risk_xchg_alloc")
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.ltr_ipc_export(local, handle), "ltr_ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.local_ptr = local.value
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(local.value)
                    continue
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                peer = ctypes.c_void_p()
                _lib.check(lib.ltr_ipc_open(buf, ctypes.byref(peer)), "ltr_ipc_open")
                ptrs.append(peer.value)
        self._ptrs = ptrs
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        dist.barrier(group=group)                 # nobody publishes before everybody has mapped everybody

    def status(self) -> int:
        st = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ltr_risk_xchg_status(self.local_ptr, ctypes.byref(st)), "ltr_risk_xchg_status")
        return st.value

    def close(self):
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for r, p in enumerate(self._ptrs):
                if r != self.rank and p:
                    self._lib.ltr_ipc_close(p)
            self._lib.ltr_risk_xchg_free(self.local_ptr)
        self._ptrs = []


class GradPeerExchange:
    """Peer-memory average of small tensors over the ranks of `group` (`ltr_peer_allreduce_avg`): what follows the fused
    ranker step under data parallelism instead of DistributedDataParallel's NCCL all-reduce of ~13 k gradient floats.
    Every rank's exchange block is mapped into every process with CUDA IPC; the collective constructor must run on every
    rank at the same point.  `average(tensors)` averages contiguous fp32 CUDA tensors IN PLACE (at most 16,384 floats in
    total, at most 8 tensors); every rank must make the same calls in the same order."""

    def __init__(self, group, device: torch.device):
        lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise _lib.LtrError("GradPeerExchange: at most 8 ranks (one box)")
        self.device, self._lib = device, lib
        local = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.ltr_peer_xchg_alloc(ctypes.byref(local)), "ltr_peer_xchg_alloc")
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.ltr_ipc_export(local, handle), "ltr_ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.local_ptr = local.value
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(local.value)
                    continue
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                peer = ctypes.c_void_p()
                _lib.check(lib.ltr_ipc_open(buf, ctypes.byref(peer)), "ltr_ipc_open")
                ptrs.append(peer.value)
        self._ptrs, self._args = ptrs, {}
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.max_floats = int(lib.ltr_peer_xchg_max_floats())
        dist.barrier(group=group)                 # nobody pushes before everybody has mapped everybody

    def average(self, tensors, stream_ptr=None):
        from . import _runtime as rt
        key = tuple((t.data_ptr(), t.numel()) for t in tensors)
        args = self._args.get(key)
        if args is None:                                   # (a training loop repeats one argument list: built once)
            for t in tensors:
                if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                    raise ValueError("GradPeerExchange.average takes contiguous fp32 tensors on the exchange's device")
            n = len(key)
            args = ((ctypes.c_void_p * n)(*[k[0] for k in key]), (ctypes.c_int * n)(*[k[1] for k in key]), n)
            if len(self._args) > 64:
                self._args.clear()
            self._args[key] = args
        stream = rt.stream_ptr() if stream_ptr is None else stream_ptr
        if torch.cuda.current_device() == self.device.index:
            rc = self._lib.ltr_peer_allreduce_avg(*args, self.local_ptr, self.peer_ptrs, self.rank, self.world, stream)
        else:
            with torch.cuda.device(self.device):
                rc = self._lib.ltr_peer_allreduce_avg(*args, self.local_ptr, self.peer_ptrs, self.rank, self.world, stream)
        _lib.check(rc, "ltr_peer_allreduce_avg")

    def status(self) -> int:
        st = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ltr_peer_xchg_status(self.local_ptr, ctypes.byref(st)), "ltr_peer_xchg_status")
        return st.value

    def close(self):
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for r, p in enumerate(self._ptrs):
                if r != self.rank and p:
                    self._lib.ltr_ipc_close(p)
            self._lib.ltr_peer_xchg_free(self.local_ptr)
        self._ptrs = []


_grad_exchanges = {}


def grad_peer_exchange(group, device: torch.device):
    """The (cached) GradPeerExchange of `group` on `device`, or None when the peers' memory cannot be mapped (the caller then
    all-reduces with torch.distributed).  The first call per group is collective."""
    if os.environ.get("LTR_GRAD_P2P", "1") == "0":
        return None
    key = (id(group), device.index)
    if key not in _grad_exchanges:
        try:
            ok, ex = 1, GradPeerExchange(group, device)
        except _lib.LtrError:
            ok, ex = 0, None
        flag = torch.tensor([ok], dtype=torch.int32, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0 and ex is not None:
            ex.close()
            ex = None
        _grad_exchanges[key] = ex
    return _grad_exchanges[key]


_peer_exchanges = {}


def risk_peer_exchange(group, device: torch.device):
    """The (cached) peer exchange of `group` on `device`, or None when in-kernel exchange is switched off
    (LTR_RISK_P2P=0) or the peers' memory cannot be mapped -- the callers then all-gather between the phases.
    The first call per group is collective."""
    if os.environ.get("LTR_RISK_P2P", "1") == "0":
        return None
    key = (id(group), device.index)
    if key not in _peer_exchanges:
        try:
            ok, ex = 1, RiskPeerExchange(group, device)
        except _lib.LtrError:
            ok, ex = 0, None
        # all or nothing: one rank without peer access sends everybody to the all-gather path
        flag = torch.tensor([ok], dtype=torch.int32, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0 and ex is not None:
            ex.close()
            ex = None
        _peer_exchanges[key] = ex
    return _peer_exchanges[key]


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


def global_pair_mean(local_sum: torch.Tensor, local_pairs: torch.Tensor, group) -> torch.Tensor:
    """lambdaLoss(reduction="mean") over query-sharded slates (lambdaL.py:88-89 divides by the GLOBAL number of masked
    pairs): from every rank's (-sum of pair log-probabilities, number of masked pairs) -> the global mean loss.
    One all-reduce of two doubles."""
    buf = torch.stack([local_sum.detach().reshape(()).double(), local_pairs.detach().reshape(()).double()])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[0] / buf[1]).to(local_sum.dtype)
