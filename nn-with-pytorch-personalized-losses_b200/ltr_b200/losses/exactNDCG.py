"""exact NDCG "loss" -- drop-in for reference losses/exactNDCG.py:5-29 (y_true FIRST).  Forward only: in the
reference the value reaches y_predicted through argsort alone, so y_predicted gets no gradient
(exactNDCG.py:7 "Todo loss not differentiable")."""
from __future__ import annotations

from .. import ops
from ._adapt import back_to, to_device


def ndcgLoss(y_true, y_predicted):
    loss = ops.exact_ndcg(to_device(y_true.detach()), to_device(y_predicted.detach()))
    return back_to(loss, y_predicted)
