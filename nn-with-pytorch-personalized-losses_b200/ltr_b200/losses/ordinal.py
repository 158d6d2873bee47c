"""Ordinal loss -- drop-in for reference losses/ordinal.py:9-53, computed by `ltr_ordinal_fwdbwd`."""
from __future__ import annotations

import torch

from .. import ops
from ._adapt import back_to, to_device

PADDED_Y_VALUE = -1


def with_ordinals(y, n, padded_value_indicator=PADDED_Y_VALUE):
    """Labels [batch, slate] -> ordinal targets [batch, slate, n]: target l is 1 when label >= l + 1, and every
    target of a doc whose label equals the padding value is the padding value (reference ordinal.py:9-24).
    Comparison / broadcast plumbing only; the loss kernel forms the same targets on the fly."""
    thresholds = torch.arange(1, n + 1, dtype=torch.float, device=y.device).view(1, 1, n)
    labels = y.unsqueeze(-1).expand(-1, -1, n)
    targets = (labels >= thresholds).to(torch.float)
    return torch.where(labels == padded_value_indicator, torch.full_like(targets, padded_value_indicator), targets)


def ordinalLoss(y_pred, y_true, n, padded_value_indicator=PADDED_Y_VALUE):
    """y_pred [batch, slate, n] probabilities of "label > level", y_true [batch, slate] labels, n ordinal levels;
    returns the masked BCE sum divided by the number of docs with an unmasked target (0-dim tensor).
    Like the reference on the installed torch this raises RuntimeError when a label equals -1
    (`with_ordinals` is called with its default padding value and BCELoss rejects the resulting -1 targets,
    ordinal.py:39-44) and when an element of y_pred lies outside [0, 1]; both conditions are detected inside the
    loss kernel and read back as one status word.
    """
    yp = to_device(y_pred)
    yt = to_device(y_true.detach())
    if yp.dim() != 3 or yp.shape[2] != n or tuple(yp.shape[:2]) != tuple(yt.shape):
        raise ValueError(f"Using a target size ({tuple(yt.shape) + (n,)}) that is different to the input size "
                         f"({tuple(yp.shape)}) is deprecated. Please ensure they have the same size.")
    loss, status = ops.ordinal(yp, yt, n, float(padded_value_indicator))
    st = int(status)
    if st & 2:
        raise RuntimeError("all elements of input should be between 0 and 1")
    if st & 1:
        raise RuntimeError("all elements of target should be between 0 and 1")
    return back_to(loss, y_pred)
