"""autograd Functions over the C-ABI (include/ltr_b200.h).

Every kernel is fused forward+backward: `forward` computes the loss AND stashes d loss / d y_pred;
`backward` only scales the stash by grad_output, so it survives `backward(retain_graph=True)` as
the reference drivers call it (main_batch_execution.py:170).  PyTorch provides device memory, the
current stream and (for query-sharded GeoRisk) the NCCL collectives; all arithmetic of the path
happens in libltr_b200.so.  There is no CPU fallback anywhere in this module.
"""
from __future__ import annotations

import ctypes
import functools

import numpy as np
import torch
from torch.autograd.function import once_differentiable

from . import _lib, _runtime as rt, sharding

TIES_STABLE, TIES_STDSORT, TIES_REVERSED = 0, 1, 2

LAMBDA_SCHEMES = {
    None: 0,
    "rankNet_scheme": 0,
    "ndcgLoss1_scheme": 1,
    "ndcgLoss2_scheme": 2,
    "lamdbaRank_scheme": 3,      # [sic] the reference's spelling (losses/lambdaL.py:109)
    "ndcgLoss2PP_scheme": 4,
    "rankNetWeightedByGTDiff_scheme": 5,
    "rankNetWeightedByGTDiffPowed_scheme": 6,
}


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _guard(fn):
    """Run `fn` with the device of its first tensor argument current: streams, workspaces and launches follow the
    data, not the caller's current device (the reference callables are device-agnostic)."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        first = args[0] if isinstance(args[0], torch.Tensor) else args[1]      # (ctx, tensor, ...) or (tensor, ...)
        if first.device.index == torch.cuda.current_device():                 # the common case: nothing to switch
            return fn(*args, **kwargs)
        with torch.cuda.device(first.device):
            return fn(*args, **kwargs)
    return wrapped


def _scaled(stash, grad_out):
    return stash * grad_out if stash is not None else None


# ------------------------------------------------------------------------------------------------
# approxNDCG                                                           losses/approxNDCG.py:7-53
# ------------------------------------------------------------------------------------------------
class _ApproxNDCG(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true, eps, pad, alpha):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        need_grad = ctx.needs_input_grad[0]
        slate = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(y_pred) if need_grad else None
        tab = rt.disc_prefix(dev)
        nscratch = lib.ltr_approx_ndcg_scratch_bytes(B, n)
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=dev) if nscratch else None
        _lib.check(lib.ltr_approx_ndcg_fwdbwd(
            y_pred.data_ptr(), y_true.data_ptr(), B, n, float(alpha), float(eps), float(pad),
            tab.data_ptr(), tab.numel(), slate.data_ptr(), loss.data_ptr(),
            _ptr(grad), _ptr(scratch), nscratch, rt.workspace(dev).data_ptr(), rt.stream_ptr()),
            "ltr_approx_ndcg_fwdbwd")
        ctx.grad = grad
        ctx.slate_ndcg = slate
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        g = ctx.grad * grad_out if ctx.grad is not None else None
        return g, None, None, None, None


def approx_ndcg(y_pred, y_true, eps=1e-10, pad=-1.0, alpha=1.0):
    """Device entry: y_pred/y_true are CUDA fp32 contiguous [B, n]."""
    return _ApproxNDCG.apply(y_pred, y_true, eps, pad, alpha)


# ------------------------------------------------------------------------------------------------
# ListNet                                                                 losses/listnet.py:5-16
# ------------------------------------------------------------------------------------------------
class _ListNet(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true, apply_sigmoid):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(y_pred) if need_grad else None
        _lib.check(lib.ltr_listnet_fwdbwd(y_true.data_ptr(), y_pred.data_ptr(), B, n, int(bool(apply_sigmoid)), None,
                                          loss.data_ptr(), _ptr(grad), rt.workspace(dev).data_ptr(), rt.stream_ptr()),
                   "ltr_listnet_fwdbwd")
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        return (ctx.grad * grad_out if ctx.grad is not None else None), None, None


def listnet(y_true, y_pred, apply_sigmoid=False):
    return _ListNet.apply(y_pred, y_true, apply_sigmoid)


# ------------------------------------------------------------------------------------------------
# ordinal                                                                 losses/ordinal.py:27-53
# ------------------------------------------------------------------------------------------------
class _Ordinal(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true, levels, pad):
        lib = _lib.load()
        B, n, L = y_pred.shape
        dev = y_pred.device
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        grad = torch.empty_like(y_pred) if need_grad else None
        _lib.check(lib.ltr_ordinal_fwdbwd(y_pred.data_ptr(), y_true.data_ptr(), B, n, L, float(pad), loss.data_ptr(),
                                          status.data_ptr(), _ptr(grad), rt.workspace(dev).data_ptr(), rt.stream_ptr()),
                   "ltr_ordinal_fwdbwd")
        ctx.grad = grad
        ctx.mark_non_differentiable(status)
        return loss.reshape(()), status

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out, _grad_status):
        return (ctx.grad * grad_out if ctx.grad is not None else None), None, None, None


def ordinal(y_pred, y_true, levels, pad=-1.0):
    """Returns (loss 0-dim, status [1] int32 on the device: bit 0 a target < 0 reached the BCE, bit 1 an input outside [0, 1])."""
    return _Ordinal.apply(y_pred, y_true, levels, pad)


# ------------------------------------------------------------------------------------------------
# LambdaLoss                                                            losses/lambdaL.py:67-93
# ------------------------------------------------------------------------------------------------
class _Lambda(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true, eps, pad, scheme, k, sigma, mu, mean, natural):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        need_grad = ctx.needs_input_grad[0]
        sums = torch.empty(2, B, dtype=torch.float32, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        grad = torch.empty_like(y_pred) if need_grad else None
        _lib.check(lib.ltr_lambda_fwdbwd(y_pred.data_ptr(), y_true.data_ptr(), B, n, float(eps), float(pad), int(scheme),
                                         int(k), float(sigma), float(mu), int(mean), int(natural), sums[0].data_ptr(),
                                         sums[1].data_ptr(), out[0:1].data_ptr(), out[1:2].data_ptr(), _ptr(grad),
                                         rt.workspace(dev).data_ptr(), rt.stream_ptr()), "ltr_lambda_fwdbwd")
        ctx.grad = grad
        ctx.slate_sums = sums
        count = out[1].reshape(())
        ctx.mark_non_differentiable(count)
        return out[0].reshape(()), count

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out, _grad_count):
        return ((ctx.grad * grad_out if ctx.grad is not None else None),) + (None,) * 9


def lambda_loss(y_pred, y_true, eps, pad, scheme, k, sigma, mu, mean, natural):
    return _Lambda.apply(y_pred, y_true, eps, pad, scheme, k, sigma, mu, mean, natural)[0]


def lambda_loss_and_count(y_pred, y_true, eps, pad, scheme, k, sigma, mu, natural):
    """(-sum of the masked pair log-probabilities, attached to autograd; number of masked pairs) of the local slates."""
    return _Lambda.apply(y_pred, y_true, eps, pad, scheme, k, sigma, mu, False, natural)


# ------------------------------------------------------------------------------------------------
# orderScore / exactNDCG                              losses/orderScore.py:5-24, exactNDCG.py:5-29
# ------------------------------------------------------------------------------------------------
class _OrderScore(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        need_grad = ctx.needs_input_grad[0]
        slate = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(y_pred) if need_grad else None
        _lib.check(lib.ltr_order_score_fwdbwd(y_true.data_ptr(), y_pred.data_ptr(), B, n, slate.data_ptr(),
                                              loss.data_ptr(), _ptr(grad), rt.stream_ptr()), "ltr_order_score_fwdbwd")
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        return (ctx.grad * grad_out if ctx.grad is not None else None), None


def order_score(y_true, y_pred):
    return _OrderScore.apply(y_pred, y_true)


@_guard
def exact_ndcg(y_true, y_pred):
    """Forward only: the reference's ndcgLoss gives y_pred no gradient (exactNDCG.py:7 'not differentiable')."""
    lib = _lib.load()
    B, n = y_pred.shape
    dev = y_pred.device
    slate = torch.empty(B, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    _lib.check(lib.ltr_exact_ndcg(y_true.data_ptr(), y_pred.data_ptr(), B, n, slate.data_ptr(), loss.data_ptr(),
                                  rt.stream_ptr()), "ltr_exact_ndcg")
    return loss.reshape(())


# ------------------------------------------------------------------------------------------------
# sorting and NDCG@k                                                     utils/metrics.py:48-104
# ------------------------------------------------------------------------------------------------
@_guard
def sort_desc(keys, tie_rule=TIES_STDSORT, want_values=True):
    """[B, n] CUDA fp32 -> (sorted values or None, int32 permutation) with the reference's tie order."""
    lib = _lib.load()
    B, n = keys.shape
    perm = torch.empty(B, n, dtype=torch.int32, device=keys.device)
    vals = torch.empty_like(keys) if want_values else None
    _lib.check(lib.ltr_sort_desc(keys.data_ptr(), B, n, int(tie_rule), perm.data_ptr(), _ptr(vals), rt.stream_ptr()),
               "ltr_sort_desc")
    return vals, perm


_disc_cache = {}


def _disc_tables(n, device):
    """disc32[r] = float32(np.log2(2 + r)) and disc64[r] = np.log2(2 + r): bit-for-bit the reference's divisors
    (metrics.py:65 and :91, where the fp32 tensor is divided by the numpy scalar rounded to fp32)."""
    key = (n, device.index)
    if key not in _disc_cache:
        d64 = np.log2(np.arange(n, dtype=np.float64) + 2.0)
        _disc_cache[key] = (torch.from_numpy(d64.astype(np.float32)).to(device), torch.from_numpy(d64).to(device))
    return _disc_cache[key]


def _ks_array(ks):
    arr = (ctypes.c_int * len(ks))(*[int(k) if k is not None else 0 for k in ks])
    return arr


@_guard
def ndcg_at_k(y_true, y_pred, ks, tie_rule=TIES_STDSORT):
    """torchNdcg arithmetic (fp32, gains 2^l - 1).  Returns [Q, len(ks)] fp32 on the device (NaN for 0/0)."""
    lib = _lib.load()
    Q, n = y_pred.shape
    d32, _ = _disc_tables(n, y_pred.device)
    out = torch.empty(Q, len(ks), dtype=torch.float32, device=y_pred.device)
    flags = torch.empty(Q + 1, dtype=torch.int32, device=y_pred.device)
    need = int(lib.ltr_ndcg_at_k_scratch_bytes(Q, n))          # long rows of whole 128-doc chunks: per-chunk top lists
    if need:
        scratch = torch.empty(need, dtype=torch.uint8, device=y_pred.device)
        _lib.check(lib.ltr_ndcg_at_k_ws(y_true.data_ptr(), y_pred.data_ptr(), Q, n, _ks_array(ks), len(ks), int(tie_rule),
                                        d32.data_ptr(), out.data_ptr(), flags.data_ptr(), scratch.data_ptr(), need,
                                        rt.stream_ptr()), "ltr_ndcg_at_k_ws")
        return out
    _lib.check(lib.ltr_ndcg_at_k(y_true.data_ptr(), y_pred.data_ptr(), Q, n, _ks_array(ks), len(ks), int(tie_rule),
                                 d32.data_ptr(), out.data_ptr(), flags.data_ptr(), rt.stream_ptr()), "ltr_ndcg_at_k")
    return out


@_guard
def ndcg_at_k_f64(y_true, y_pred, ks, tie_rule=TIES_STABLE, exponential=False, empty_value=1.0):
    """metrics.dcg/ndcg arithmetic (fp64).  Returns [Q, len(ks)] fp64 on the device."""
    lib = _lib.load()
    Q, n = y_pred.shape
    _, d64 = _disc_tables(n, y_pred.device)
    out = torch.empty(Q, len(ks), dtype=torch.float64, device=y_pred.device)
    flags = torch.empty(Q + 1, dtype=torch.int32, device=y_pred.device)
    _lib.check(lib.ltr_ndcg_at_k_f64(y_true.data_ptr(), y_pred.data_ptr(), Q, n, _ks_array(ks), len(ks), int(tie_rule),
                                     int(bool(exponential)), float(empty_value), d64.data_ptr(), out.data_ptr(),
                                     flags.data_ptr(), rt.stream_ptr()), "ltr_ndcg_at_k_f64")
    return out


@_guard
def dcg_at_k_f64(y_true, y_pred, ks, tie_rule=TIES_STABLE, exponential=False):
    """metrics.dcg arithmetic (fp64, un-normalised).  Returns [Q, len(ks)] fp64 on the device."""
    lib = _lib.load()
    Q, n = y_pred.shape
    _, d64 = _disc_tables(n, y_pred.device)
    out = torch.empty(Q, len(ks), dtype=torch.float64, device=y_pred.device)
    flags = torch.empty(Q + 1, dtype=torch.int32, device=y_pred.device)
    _lib.check(lib.ltr_dcg_at_k_f64(y_true.data_ptr(), y_pred.data_ptr(), Q, n, _ks_array(ks), len(ks), int(tie_rule),
                                    int(bool(exponential)), d64.data_ptr(), out.data_ptr(), flags.data_ptr(),
                                    rt.stream_ptr()), "ltr_dcg_at_k_f64")
    return out


def georisk_default_f64(mat, alpha):
    """getGeoRiskDefault (metrics.py:8-45): fp64 [Q, C] (any device) -> fp64 [C] on the GPU."""
    lib = _lib.load()
    rt.require_cuda()
    if mat.device.type != "cuda":
        mat = mat.to(torch.device("cuda", torch.cuda.current_device()))
    mat = mat.to(torch.float64).contiguous()
    Q, C = mat.shape
    with torch.cuda.device(mat.device):
        out = torch.empty(C, dtype=torch.float64, device=mat.device)
        _lib.check(lib.ltr_georisk_default_f64(mat.data_ptr(), Q, C, float(alpha), out.data_ptr(), rt.stream_ptr()),
                   "ltr_georisk_default_f64")
    return out


# ------------------------------------------------------------------------------------------------
# risk functions and the ListNet-style risk losses       losses/riskLosses/riskFunctions.py, riskLosses.py
# ------------------------------------------------------------------------------------------------
KIND_Z, KIND_GEO, KIND_T = 0, 1, 2


@_guard
def _risk_phases(mat, kind, invert, strategy, quirk, alpha, negative, want_gq, want_gmat, group=None):
    """Runs the three phases of ltr_risk_phase on the local rows `mat` [Q, C] (Q = 0: an empty shard); with a process
    group the ranks' phase states are all-gathered between the phases (one collective per round, on the current
    stream) and every rank combines them in rank order.  Returns (loss[1], gq[Q] or None, gmat[Q, C] or None)."""
    lib = _lib.load()
    Q, C = mat.shape
    dev = mat.device
    xchg = rt.risk_xchg(dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    gq = torch.empty(Q, dtype=torch.float32, device=dev) if want_gq else None
    gmat = torch.empty(Q, C, dtype=torch.float32, device=dev) if want_gmat else None
    sharded = sharding.is_sharded(group)
    rank, world = sharding.rank_world(group)

    def phase(i):
        _lib.check(lib.ltr_risk_phase(i, mat.data_ptr() if Q else None, Q, C, int(kind), int(invert), int(strategy),
                                      int(quirk), float(alpha), float(negative), xchg.data_ptr(), rank, world,
                                      loss.data_ptr(), _ptr(gq) if Q else None, _ptr(gmat) if Q else None,
                                      rt.stream_ptr()), "ltr_risk_phase")

    phase(1)
    if sharded:
        sharding.allgather_risk_round(xchg, 0, group)
    phase(2)
    if sharded:
        sharding.allgather_risk_round(xchg, 1, group)
    phase(3)
    return loss, gq, gmat


class _RiskFn(torch.autograd.Function):
    """zRisk / geoRisk of column 0 of `mat` (callers permute the requested column to the front)."""

    @staticmethod
    @_guard
    def forward(ctx, mat, kind, alpha):
        need = ctx.needs_input_grad[0]
        loss, _, gmat = _risk_phases(mat, kind, 0, 1, 0, alpha, 1.0, False, need)
        ctx.gmat = gmat
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        return (ctx.gmat * grad_out if ctx.gmat is not None else None), None, None


def risk_fn(mat, kind, alpha):
    return _RiskFn.apply(mat, kind, alpha)


class _RiskListnet(torch.autograd.Function):
    """geoRisk / zRisk / tRisk ListNet losses: `ltr_risk_listnet_fwdbwd`, three launches (rows + phase-1 state, z-risk
    sums, loss + backward rows).  Query-sharded (`group`): the ranks' phase states are exchanged either inside the
    kernels over NVLink peer memory (sharding.RiskPeerExchange: no collective, one library call) or, when peer
    memory is not available, by one all-gather per round between the launches."""

    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true, y_base, kind, alpha, lt, strategy, negative, add_ideal, quirk, group):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        S = 0 if y_base is None else y_base.shape[2]
        C = 1 + S + (1 if add_ideal else 0)
        stream = rt.stream_ptr()
        ws, xchg, mat, rowstate = rt.risk_scratch(dev, stream, B, C)   # per (device, stream): stream-ordered reuse is safe
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        # mat <- max(mat) - mat: riskLosses.py:49-50 (lt 1, 3) and :274-275 (t-risk: lt 1 only)
        invert = (lt == 1) if kind == KIND_T else (lt in (1, 3))
        need = ctx.needs_input_grad[0]
        grad = torch.empty_like(y_pred) if need else None
        sharded = sharding.is_sharded(group)
        rank, world = sharding.rank_world(group) if sharded else (0, 1)
        peers = sharding.risk_peer_exchange(group, dev) if sharded else None
        ptrs = (y_pred.data_ptr(), y_true.data_ptr(), _ptr(y_base), mat.data_ptr(), _ptr(grad)) if B else (None,) * 5
        args = (B, n, S, int(lt), int(bool(add_ideal)), int(kind), int(invert), int(strategy), int(quirk), float(alpha),
                float(negative), ptrs[3], rowstate.data_ptr(), loss.data_ptr(), ptrs[4])

        def call(phases, xchg_ptr, peer_ptrs):
            rc = lib.ltr_risk_listnet_fwdbwd(ptrs[0], ptrs[1], ptrs[2], *args, phases, xchg_ptr, rank, world, peer_ptrs,
                                             ws.data_ptr(), stream)
            if rc:
                _lib.check(rc, "ltr_risk_listnet_fwdbwd")

        if peers is not None:
            call(7, peers.local_ptr, peers.peer_ptrs)
        elif sharded:
            call(1, xchg.data_ptr(), None)
            sharding.allgather_risk_round(xchg, 0, group)
            call(2, xchg.data_ptr(), None)
            sharding.allgather_risk_round(xchg, 1, group)
            call(4, xchg.data_ptr(), None)
        else:
            call(7, xchg.data_ptr(), None)
        ctx.grad = grad
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        return ((ctx.grad * grad_out if ctx.grad is not None else None),) + (None,) * 10


def risk_listnet(y_pred, y_true, y_base, kind, alpha, lt, strategy, negative, add_ideal, quirk=False, group=None):
    return _RiskListnet.apply(y_pred, y_true, y_base, kind, alpha, lt, strategy, negative, add_ideal, quirk, group)


class _RiskLambda(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true, y_base, kind, alpha, lt, strategy, negative, add_ideal, scheme, group):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        S = 0 if y_base is None else y_base.shape[2]
        C = 1 + S + (1 if add_ideal else 0)
        need = ctx.needs_input_grad[0]
        mat = torch.empty(B, C, dtype=torch.float32, device=dev)
        dvec = torch.empty(B, n, dtype=torch.float32, device=dev) if need else None
        if B > 0:                                            # (B = 0: an empty shard still joins both exchange rounds)
            _lib.check(lib.ltr_risk_lambda_rows(y_pred.data_ptr(), y_true.data_ptr(), _ptr(y_base), B, n, S, int(lt),
                                                int(bool(add_ideal)), int(kind == KIND_GEO), int(scheme), mat.data_ptr(),
                                                _ptr(dvec), rt.stream_ptr()), "ltr_risk_lambda_rows")
        invert = (lt == 1)                                   # riskLosses.py:114-115, :335
        loss, gq, _ = _risk_phases(mat, kind, invert, strategy, 0, alpha, negative, need, False, group)
        grad = None
        if need:
            grad = torch.empty_like(y_pred)
        if need and B > 0:
            _lib.check(lib.ltr_risk_lambda_rows_bwd(y_pred.data_ptr(), y_true.data_ptr(), B, n, int(scheme),
                                                    gq.data_ptr(), dvec.data_ptr(), grad.data_ptr(), rt.stream_ptr()),
                       "ltr_risk_lambda_rows_bwd")
        ctx.grad = grad
        ctx.mat = mat
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        return ((ctx.grad * grad_out if ctx.grad is not None else None),) + (None,) * 10


def risk_lambda(y_pred, y_true, y_base, kind, alpha, lt, strategy, negative, add_ideal, scheme, group=None):
    return _RiskLambda.apply(y_pred, y_true, y_base, kind, alpha, lt, strategy, negative, add_ideal, scheme, group)


class _LambdaPairMatrix(torch.autograd.Function):
    """lambdaMask's [B, n, n] pair log-probabilities (rank order), differentiable w.r.t. y_pred (lambdaL.py:7-64)."""

    @staticmethod
    @_guard
    def forward(ctx, y_pred, y_true, eps, pad, scheme, k, sigma, mu, natural, want_mask):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        logp = torch.empty(B, n, n, dtype=torch.float32, device=dev)
        mask = torch.empty(B, n, n, dtype=torch.uint8, device=dev) if want_mask else None
        _lib.check(lib.ltr_lambda_pair_matrix(y_pred.data_ptr(), y_true.data_ptr(), B, n, float(eps), float(pad), int(scheme),
                                              int(k), float(sigma), float(mu), int(natural), logp.data_ptr(), _ptr(mask),
                                              rt.stream_ptr()), "ltr_lambda_pair_matrix")
        ctx.save_for_backward(y_pred, y_true)
        ctx.args = (float(eps), float(pad), int(scheme), int(k), float(sigma), float(mu), int(natural))
        if mask is None:
            return logp
        ctx.mark_non_differentiable(mask)
        return logp, mask

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_logp, _grad_mask=None):
        y_pred, y_true = ctx.saved_tensors
        eps, pad, scheme, k, sigma, mu, natural = ctx.args
        B, n = y_pred.shape
        with torch.cuda.device(y_pred.device):
            g = grad_logp.contiguous()
            grad = torch.empty_like(y_pred)
            _lib.check(_lib.load().ltr_lambda_pair_matrix_bwd(y_pred.data_ptr(), y_true.data_ptr(), B, n, eps, pad, scheme, k,
                                                              sigma, mu, natural, g.data_ptr(), grad.data_ptr(),
                                                              rt.stream_ptr()), "ltr_lambda_pair_matrix_bwd")
        return (grad,) + (None,) * 9


def lambda_pair_matrix(y_pred, y_true, eps, pad, scheme, k, sigma, mu, natural, want_mask):
    """[B, n, n] pair log-probabilities in rank order (+ the lambdaLoss pair mask as uint8), attached to autograd."""
    if want_mask:
        return _LambdaPairMatrix.apply(y_pred, y_true, eps, pad, scheme, k, sigma, mu, natural, True)
    return _LambdaPairMatrix.apply(y_pred, y_true, eps, pad, scheme, k, sigma, mu, natural, False), None
