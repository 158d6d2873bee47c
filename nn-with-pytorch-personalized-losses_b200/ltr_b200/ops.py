"""autograd Functions over the C-ABI: forward computes the loss AND stashes d loss / d y_pred
(the kernels are fused fwd+bwd); backward only scales the stash by grad_output, so it survives
`backward(retain_graph=True)` as the reference drivers call it (main_batch_execution.py:170)."""
from __future__ import annotations

import torch

from . import _lib, _runtime as rt


class _ApproxNDCG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, eps, pad, alpha):
        lib = _lib.load()
        B, n = y_pred.shape
        dev = y_pred.device
        need_grad = ctx.needs_input_grad[0]
        slate = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(y_pred) if need_grad else None
        tab = rt.disc_prefix(dev)
        _lib.check(lib.ltr_approx_ndcg_fwdbwd(
            y_pred.data_ptr(), y_true.data_ptr(), B, n, float(alpha), float(eps), float(pad),
            tab.data_ptr(), tab.numel(), slate.data_ptr(), loss.data_ptr(),
            grad.data_ptr() if need_grad else None, rt.workspace(dev).data_ptr(), rt.stream_ptr()),
            "ltr_approx_ndcg_fwdbwd")
        ctx.grad = grad
        ctx.slate_ndcg = slate
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.grad * grad_out if ctx.grad is not None else None
        return g, None, None, None, None


def approx_ndcg(y_pred, y_true, eps=1e-10, pad=-1.0, alpha=1.0):
    """Device entry: y_pred/y_true are CUDA fp32 contiguous [B, n]."""
    return _ApproxNDCG.apply(y_pred, y_true, eps, pad, alpha)
