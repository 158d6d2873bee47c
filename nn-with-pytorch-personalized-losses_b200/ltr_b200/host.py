"""Host-buffer front end: pinned CPU tensors in, pinned CPU results out, every copy and the fused
kernel queued on one CUDA stream by a single C-ABI call (`ltr_approx_ndcg_fwdbwd_host`).

This is the path a CPU-resident caller (the reference's drivers keep everything on the CPU) uses
when it wants loss AND gradient without going through autograd."""
from __future__ import annotations

import torch

from . import _lib, _runtime as rt


class ApproxNDCGHost:
    """Reusable plan for one (B, n): owns device scratch + pinned outputs for `depth` in-flight steps."""

    def __init__(self, B: int, n: int, alpha: float = 1.0, eps: float = 1e-10, pad: float = -1.0,
                 device=None, depth: int = 3, want_grad: bool = True):
        rt.require_cuda()
        self.lib = _lib.load()
        self.B, self.n, self.alpha, self.eps, self.pad = B, n, float(alpha), float(eps), float(pad)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        nbytes = self.lib.ltr_approx_ndcg_host_scratch_bytes(B, n)
        self.scratch_bytes = nbytes
        with torch.cuda.device(self.device):
            self.tab = rt.disc_prefix(self.device)
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
            self.scratch = [torch.empty(nbytes, dtype=torch.uint8, device=self.device) for _ in range(depth)]
            self.ws = [torch.zeros(_lib.WORKSPACE_BYTES // 4, dtype=torch.int32, device=self.device) for _ in range(depth)]
        self.loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.grad = [torch.empty(B, n, dtype=torch.float32).pin_memory() for _ in range(depth)] if want_grad else None
        self._slot = 0

    def submit(self, y_pred: torch.Tensor, y_true: torch.Tensor):
        """Queue one fwd+bwd on the next slot; returns (slot, loss_pinned, grad_pinned). Asynchronous:
        call `wait(slot)` (or synchronize the slot's stream) before reading the outputs."""
        assert y_pred.device.type == "cpu" and y_true.device.type == "cpu"
        assert y_pred.dtype == torch.float32 and y_pred.is_contiguous() and y_true.is_contiguous()
        k = self._slot
        self._slot = (k + 1) % self.depth
        st = self.streams[k]
        g = self.grad[k] if self.grad is not None else None
        with torch.cuda.device(self.device):           # the plan's device, whatever the caller's current one is
            _lib.check(self.lib.ltr_approx_ndcg_fwdbwd_host(
                y_pred.data_ptr(), y_true.data_ptr(), self.B, self.n, self.alpha, self.eps, self.pad,
                self.tab.data_ptr(), self.tab.numel(), self.loss[k].data_ptr(),
                g.data_ptr() if g is not None else None, None,
                self.scratch[k].data_ptr(), self.scratch_bytes, self.ws[k].data_ptr(), st.cuda_stream),
                "ltr_approx_ndcg_fwdbwd_host")
        return k, self.loss[k], g

    def wait(self, slot: int):
        self.streams[slot].synchronize()

    def wait_all(self):
        for st in self.streams:
            st.synchronize()


class NdcgEvalHost:
    """Reusable plan for an NDCG@k sweep over host-resident evaluation data (the reference keeps validation labels
    and predictions on the CPU and calls torchNdcg / mNdcg once per epoch, main.py:145): pinned [Q, n] labels and
    scores in, pinned [Q, len(ks)] fp32 NDCG out, every copy and the sweep kernel queued on one of `depth` streams
    by a single C-ABI call (`ltr_ndcg_at_k_host`)."""

    def __init__(self, Q: int, n: int, ks=(1, 5, 10), tie_rule: int = 1, device=None, depth: int = 3):
        import ctypes
        from . import ops
        rt.require_cuda()
        self.lib = _lib.load()
        self.Q, self.n, self.ks, self.tie_rule = Q, n, list(ks), int(tie_rule)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        self._karr = (ctypes.c_int * len(self.ks))(*[int(k) if k is not None else 0 for k in self.ks])
        self.scratch_bytes = self.lib.ltr_ndcg_at_k_host_scratch_bytes(Q, n, len(self.ks))
        with torch.cuda.device(self.device):
            self.d32, _ = ops._disc_tables(n, self.device)
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
            self.scratch = [torch.empty(self.scratch_bytes, dtype=torch.uint8, device=self.device) for _ in range(depth)]
        self.out = [torch.empty(Q, len(self.ks), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self._slot = 0

    def submit(self, y_true: torch.Tensor, y_pred: torch.Tensor):
        """Queue one sweep on the next slot; returns (slot, out_pinned).  Asynchronous: `wait(slot)` before reading."""
        assert y_pred.device.type == "cpu" and y_true.device.type == "cpu"
        assert y_pred.dtype == torch.float32 and y_pred.is_contiguous() and y_true.is_contiguous()
        k = self._slot
        self._slot = (k + 1) % self.depth
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ltr_ndcg_at_k_host(y_true.data_ptr(), y_pred.data_ptr(), self.Q, self.n, self._karr,
                                                   len(self.ks), self.tie_rule, self.d32.data_ptr(), self.out[k].data_ptr(),
                                                   self.scratch[k].data_ptr(), self.scratch_bytes,
                                                   self.streams[k].cuda_stream), "ltr_ndcg_at_k_host")
        return k, self.out[k]

    def wait(self, slot: int):
        self.streams[slot].synchronize()

    def wait_all(self):
        for st in self.streams:
            st.synchronize()
