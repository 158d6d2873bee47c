"""Ordinal loss -- drop-in for reference losses/ordinal.py:9-53, computed by `ltr_ordinal_fwdbwd`."""
from __future__ import annotations

import torch

from .. import ops
from ._adapt import back_to, to_device

PADDED_Y_VALUE = -1


def with_ordinals(y, n, padded_value_indicator=PADDED_Y_VALUE):
    """
    Helper of the reference (ordinal.py:9-24): labels [batch, slate] -> ordinals [batch, slate, n].
    Index plumbing only (comparisons against 1..n); the loss kernel forms the same targets on the fly.
    """
    one_to_n = torch.arange(start=1, end=n + 1, dtype=torch.float, device=y.device)
    unsqueezed = y.unsqueeze(2).repeat(1, 1, n)
    mask = unsqueezed == padded_value_indicator
    ordinals = (unsqueezed >= one_to_n).type(torch.float)
    ordinals[mask] = padded_value_indicator
    return ordinals


def ordinalLoss(y_pred, y_true, n, padded_value_indicator=PADDED_Y_VALUE):
    """
    :param y_pred: predictions from the model, shape [batch_size, slate_length, n]
    :param y_true: ground truth labels, shape [batch_size, slate_length]
    :param n: number of ordinal values, int
    :param padded_value_indicator: an indicator of the y_true index containing a padded item, e.g. -1
    :return: loss value, a torch.Tensor
    Like the reference this raises ValueError when a label equals -1: `with_ordinals` is called
    with its default padding value and torch's BCELoss rejects the resulting -1 targets (ordinal.py:39-44).
    """
    yp = to_device(y_pred)
    yt = to_device(y_true.detach())
    if yp.dim() != 3 or yp.shape[2] != n or tuple(yp.shape[:2]) != tuple(yt.shape):
        raise ValueError(f"Using a target size ({tuple(yt.shape) + (n,)}) that is different to the input size "
                         f"({tuple(yp.shape)}) is deprecated. Please ensure they have the same size.")
    if bool((yt == PADDED_Y_VALUE).any()):
        raise ValueError("all elements of target should be between 0 and 1")
    loss = ops.ordinal(yp, yt, n, float(padded_value_indicator))
    return back_to(loss, y_pred)
