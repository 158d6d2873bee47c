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
