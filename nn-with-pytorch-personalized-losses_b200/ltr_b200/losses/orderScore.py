"""orderScore loss -- drop-in for reference losses/orderScore.py:5-24 (y_true FIRST), computed by
`ltr_order_score_fwdbwd` (label sort with torch.argsort's tie order + log-softmax, fused fwd+bwd)."""
from __future__ import annotations

from .. import ops
from ._adapt import back_to, to_device


def orderScoreLoss(y_true, y_predicted):
    loss = ops.order_score(to_device(y_true.detach()), to_device(y_predicted))
    return back_to(loss, y_predicted)
