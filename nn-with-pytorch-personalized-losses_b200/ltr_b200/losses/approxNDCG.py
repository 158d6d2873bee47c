"""approxNDCG loss -- drop-in for reference losses/approxNDCG.py:7-53 (same name, argument
order and defaults), computed by the fused sm_100a kernel `ltr_approx_ndcg_fwdbwd`."""
from __future__ import annotations

from .. import ops
from ._adapt import back_to, to_device


def approxNDCGLoss(y_pred, y_true, eps=1e-10, padded_value_indicator=-1, alpha=1.):
    """
    :param y_pred: predictions from the model, shape [batch_size, slate_length]
    :param y_true: ground truth labels, shape [batch_size, slate_length]
    :param eps: epsilon value, used for numerical stability
    :param padded_value_indicator: label value marking a padded item, e.g. -1
    :param alpha: score difference weight used in the sigmoid function
    :return: loss value, a 0-dim torch.Tensor attached to autograd (approxNDCG.py:53)
    Inputs are not modified (the reference clones them, approxNDCG.py:19-20).
    """
    loss = ops.approx_ndcg(to_device(y_pred), to_device(y_true.detach()), eps, float(padded_value_indicator), alpha)
    return back_to(loss, y_pred)
