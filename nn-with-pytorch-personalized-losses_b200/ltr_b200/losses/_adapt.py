"""Shared adapter: the reference callables take CPU tensors; ours run on the GPU.

Inputs on the CPU are copied to the current CUDA device (through autograd, so `.grad` lands on
the caller's CPU leaf) and the scalar loss is returned on the input's device.  CUDA inputs stay
where they are.  Nothing is ever computed on the CPU.
"""
from __future__ import annotations

import torch

from .. import _runtime as rt


def to_device(t: torch.Tensor) -> torch.Tensor:
    if t.device.type == "cuda":
        return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()
    rt.require_cuda()
    return t.to(torch.device("cuda", torch.cuda.current_device()), dtype=torch.float32, non_blocking=True).contiguous()


def back_to(loss: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return loss if like.device.type == "cuda" else loss.to(like.device)


def to_device_2d(t: torch.Tensor) -> torch.Tensor:
    """[B, n] view for the row kernels; a trailing singleton dim (the rankers emit [B, n, 1]) is dropped."""
    t = to_device(t)
    if t.dim() == 3 and t.shape[2] == 1:
        t = t.squeeze(2)
    if t.dim() != 2:
        raise ValueError(f"expected a [batch, slate] tensor, got shape {tuple(t.shape)}")
    return t.contiguous()
