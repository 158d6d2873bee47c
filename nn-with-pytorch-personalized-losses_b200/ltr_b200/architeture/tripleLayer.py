"""TripleLayerNet -- drop-in for reference architeture/tripleLayer.py:5-17 (same module: l1 136->64 with NO activation,
sigmoid(l2) 64->32, l3 32->1; `forward(x, mask, indices)` returns [B, n, 1]) -- plus the fused training step of BASELINE
configs[1]: ranker forward, approxNDCG forward + backward and all six parameter gradients from ONE pass over the
feature tensor (`ltr_triple_approx_ndcg_fwdbwd`: TMA-fed tcgen05 tf32 MMAs, the loss in registers in between).

    net = TripleLayerNet(136).cuda()
    loss = net.approx_ndcg_loss(X, y_true, alpha=10.)       # == approxNDCGLoss(net(X, None, None).squeeze(-1), y_true, alpha=10.)
    loss.backward()                                           # l1 / l2 / l3 .weight / .bias gradients

The dense layers of `forward` stay stock `torch.nn` (cuBLAS), as north_star asks; only the fused step is custom.
Operands are rounded to tf32 (10-bit mantissa) inside the tensor cores: scores / loss / gradients agree with the
fp32 modules to ~1e-3 relative.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd.function import once_differentiable

from .. import _lib, _runtime as rt

SLATE = 128          # docs per slate of the fused path (one MMA tile)
MAX_FEATURES = 160   # the fused kernel stages up to 5 chunks of 32 features (MSLR-WEB10K: 136)


_scratch_cache = {}          # (device, batch, features) -> uint8 scratch of the step (partials, folded weights): reused, never returned


def _scratch(lib, dev, B, n, F):
    key = (dev, B, F)
    buf = _scratch_cache.get(key)
    if buf is None:
        buf = torch.empty(lib.ltr_triple_step_scratch_bytes(B, n, F), dtype=torch.uint8, device=dev)
        _scratch_cache.clear()                                # one shape at a time: a training loop has one
        _scratch_cache[key] = buf
    return buf


class _FusedTripleApproxNDCG(torch.autograd.Function):
    """forward = the whole step (`ltr_triple_approx_ndcg_step`: fold, fused kernel, unfold -- four launches, ONE library call,
    nothing read back to the host); the six parameter gradients come out of one flat buffer and are scaled by the incoming
    gradient of the loss in backward."""

    @staticmethod
    def forward(ctx, X, y_true, W1, b1, W2, b2, W3, b3, eps, pad, alpha, kind, group=None):
        lib = _lib.load()
        B, n, F = X.shape
        dev = X.device
        with torch.cuda.device(dev):
            prm = [t.detach() for t in (W1, b1, W2, b2, W3, b3)]
            if any(t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev for t in prm):
                prm = [t.to(dev, torch.float32).contiguous() for t in prm]
            sizes = [t.numel() for t in prm]
            # [loss | pad to 4 | gW1 | gb1 | gW2 | gb2 | gW3 | gb3]; scores and per-slate values separately
            flat = torch.empty(4 + sum(sizes), dtype=torch.float32, device=dev)
            offs, o = [], 4
            for sz in sizes:
                offs.append(o)
                o += sz
            scores = torch.empty(B, n, dtype=torch.float32, device=dev)
            slate = torch.empty(B, dtype=torch.float32, device=dev)
            scratch = _scratch(lib, dev, B, n, F)
            tab = rt.disc_prefix(dev)
            base = flat.data_ptr()
            if kind == "listnet":
                _lib.check(lib.ltr_triple_listnet_step(
                    X.data_ptr(), y_true.data_ptr(), *[t.data_ptr() for t in prm], B, n, F, tab.data_ptr(), tab.numel(),
                    scores.data_ptr(), slate.data_ptr(), base, *[base + 4 * o for o in offs],
                    scratch.data_ptr(), scratch.numel(), rt.workspace(dev).data_ptr(), rt.stream_ptr()), "ltr_triple_listnet_step")
            else:
                _lib.check(lib.ltr_triple_approx_ndcg_step(
                    X.data_ptr(), y_true.data_ptr(), *[t.data_ptr() for t in prm], B, n, F, float(alpha), float(eps), float(pad),
                    tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), base, *[base + 4 * o for o in offs],
                    scratch.data_ptr(), scratch.numel(), rt.workspace(dev).data_ptr(), rt.stream_ptr()), "ltr_triple_approx_ndcg_step")
            if group is not None:
                # data parallel: the six gradients (one contiguous run of `flat`) are averaged over the ranks right behind the
                # step -- through peer memory (two small launches, no collective call) or, without peer access, by NCCL
                import torch.distributed as dist
                from .. import sharding
                if dist.get_world_size(group) > 1:
                    ex = sharding.grad_peer_exchange(group, dev)
                    if ex is not None and flat.numel() - 4 <= ex.max_floats:
                        ex.average([flat[4:]])
                    else:
                        dist.all_reduce(flat[4:], group=group)
                        flat[4:].mul_(1.0 / dist.get_world_size(group))
        ctx.flat, ctx.offs, ctx.shapes = flat, offs, [t.shape for t in prm]
        ctx.scores = scores
        ctx.slate_ndcg = slate
        ctx.mark_non_differentiable(scores)
        return flat[0].clone(), scores

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss, _grad_scores):
        g = ctx.flat * grad_loss                                  # one launch for all six
        out = [g[o:o + s.numel()].view(s) for o, s in zip(ctx.offs, ctx.shapes)]
        return (None, None, out[0], out[1], out[2], out[3], out[4], out[5], None, None, None, None, None)


def fused_triple_approx_ndcg(net, X, y_true, eps=1e-10, padded_value_indicator=-1, alpha=1., return_scores=False,
                             process_group=None):
    """approxNDCGLoss(net(X), y_true) with its backward through `net` (a TripleLayerNet-shaped module: l1, l2, l3), fused.
    X: [batch, 128, F <= 160] CUDA fp32, y_true: [batch, 128]; returns the 0-dim loss (and the [batch, 128] scores).
    `process_group` (data parallel, extension): the parameter gradients that `backward()` delivers are already AVERAGED over
    the group's ranks, as DistributedDataParallel would leave them -- use the plain module, not a DDP wrapper; the returned
    loss stays this rank's own."""
    rt.require_cuda()
    if X.dim() != 3 or X.shape[1] != SLATE or X.shape[2] > MAX_FEATURES or X.shape[2] % 4:
        raise ValueError(f"the fused step takes [batch, {SLATE}, F] features with F <= {MAX_FEATURES}, F % 4 == 0; got {tuple(X.shape)}")
    if X.device.type != "cuda" or X.dtype != torch.float32:
        raise ValueError("the fused step takes CUDA fp32 features")
    X = X.detach().contiguous()
    y = y_true.detach().to(X.device, dtype=torch.float32).contiguous()
    loss, scores = _FusedTripleApproxNDCG.apply(X, y, net.l1.weight, net.l1.bias, net.l2.weight, net.l2.bias, net.l3.weight,
                                                net.l3.bias, eps, float(padded_value_indicator), alpha, "approx_ndcg", process_group)
    return (loss, scores) if return_scores else loss


def fused_triple_listnet(net, X, y_true, apply_sigmoid=False, return_scores=False, process_group=None):
    """listnetLoss(y_true, net(X).squeeze(-1)) (losses/listnet.py:5-16, note the reference's argument order) with its backward
    through `net`, fused like `fused_triple_approx_ndcg`.  X: [batch, 128, F <= 160] CUDA fp32, y_true: [batch, 128].
    The loss is a SUM over slates: with `process_group=` the averaged gradients are 1 / world of the whole-batch ones (as under
    DistributedDataParallel; scale the learning rate or the loss by the world size for the reference's single-process step)."""
    rt.require_cuda()
    if apply_sigmoid:
        raise NotImplementedError("the fused ListNet step implements the default apply_sigmoid=False (losses/listnet.py:16)")
    if X.dim() != 3 or X.shape[1] != SLATE or X.shape[2] > MAX_FEATURES or X.shape[2] % 4:
        raise ValueError(f"the fused step takes [batch, {SLATE}, F] features with F <= {MAX_FEATURES}, F % 4 == 0; got {tuple(X.shape)}")
    if X.device.type != "cuda" or X.dtype != torch.float32:
        raise ValueError("the fused step takes CUDA fp32 features")
    X = X.detach().contiguous()
    y = y_true.detach().to(X.device, dtype=torch.float32).contiguous()
    loss, scores = _FusedTripleApproxNDCG.apply(X, y, net.l1.weight, net.l1.bias, net.l2.weight, net.l2.bias, net.l3.weight,
                                                net.l3.bias, 0.0, -1.0, 1.0, "listnet", process_group)
    return (loss, scores) if return_scores else loss


class TripleLayerNet(nn.Module):
    def __init__(self, N_features):
        super().__init__()
        self.l1 = nn.Linear(N_features, 64)
        self.l2 = nn.Linear(64, 32)
        self.l3 = nn.Linear(32, 1)

    def forward(self, x, mask=None, indices=None):
        return self.l3(torch.sigmoid(self.l2(self.l1(x))))

    def approx_ndcg_loss(self, x, y_true, eps=1e-10, padded_value_indicator=-1, alpha=1., process_group=None):
        """The fused training step (extension): same value as approxNDCGLoss(self(x).squeeze(-1), y_true, ...)."""
        return fused_triple_approx_ndcg(self, x, y_true, eps, padded_value_indicator, alpha, process_group=process_group)

    def listnet_loss(self, x, y_true, process_group=None):
        """The fused training step with ListNet: same value as listnetLoss(y_true, self(x).squeeze(-1))."""
        return fused_triple_listnet(self, x, y_true, process_group=process_group)
