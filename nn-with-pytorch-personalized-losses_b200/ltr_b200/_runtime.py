"""Device-side plumbing: per-(device, stream) workspaces, discount tables, tensor checks.

PyTorch is used for device memory and streams only; all arithmetic happens in libltr_b200.so.
"""
from __future__ import annotations

import torch

from . import _lib

_DISC_LEN = 16384 + 1
_workspaces = {}
_risk_blocks = {}
_disc_tables = {}


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.LtrError("ltr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def workspace(device: torch.device) -> torch.Tensor:
    """Zero-initialised scratch (counters) owned by one (device, stream); kernels leave it zeroed."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr())
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(_lib.WORKSPACE_BYTES // 4, dtype=torch.int32, device=device)
        _workspaces[key] = ws
    return ws


def risk_xchg(device: torch.device) -> torch.Tensor:
    """Zero-initialised exchange block of the risk phases (include/ltr_b200.h: LTR_RISK_XCHG_BYTES) owned by one
    (device, stream), as fp64 words: [0 : 2*16*16] is inbox[round][rank][16]."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr())
    blk = _risk_blocks.get(key)
    if blk is None:
        blk = torch.zeros(_lib.RISK_XCHG_BYTES // 8, dtype=torch.float64, device=device)
        _risk_blocks[key] = blk
    return blk


_risk_scratch = {}


def risk_scratch(device: torch.device, stream: int, B: int, C: int):
    """(workspace, exchange block, mat [B, C] fp32, rowstate [max(B, 1), 3] fp64) of the fused risk losses for one
    (device, stream): pure scratch between the launches of one call, so calls on the same stream may share it."""
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev_index, stream)
    ent = _risk_scratch.get(key)
    if ent is None or ent[2].shape[0] < B or ent[2].shape[1] != C:
        ws = _workspaces.get(key)
        if ws is None:
            ws = torch.zeros(_lib.WORKSPACE_BYTES // 4, dtype=torch.int32, device=device)
            _workspaces[key] = ws
        blk = _risk_blocks.get(key)
        if blk is None:
            blk = torch.zeros(_lib.RISK_XCHG_BYTES // 8, dtype=torch.float64, device=device)
            _risk_blocks[key] = blk
        rows = max(B, 1)
        ent = (ws, blk, torch.empty(rows, C, dtype=torch.float32, device=device),
               torch.empty(rows, 3, dtype=torch.float64, device=device))
        _risk_scratch[key] = ent
    return ent


def disc_prefix(device: torch.device) -> torch.Tensor:
    """fp64 table P[r] = sum_{q<=r} 1/log2(1+q), built on the device by the library."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    tab = _disc_tables.get(key)
    if tab is None:
        with torch.cuda.device(device):          # the library also fills a per-device __constant__ copy
            tab = torch.empty(_DISC_LEN, dtype=torch.float64, device=device)
            _lib.check(_lib.load().ltr_fill_disc_prefix(tab.data_ptr(), _DISC_LEN, stream_ptr()), "ltr_fill_disc_prefix")
        _disc_tables[key] = tab
    return tab


def as_device_f32(t: torch.Tensor, device=None) -> torch.Tensor:
    """fp32, contiguous, on the GPU.  CPU tensors (the reference's calling convention) are copied over."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if t.device.type != "cuda":
        require_cuda()
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        t = t.to(dev, dtype=torch.float32, non_blocking=True)
    elif t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
