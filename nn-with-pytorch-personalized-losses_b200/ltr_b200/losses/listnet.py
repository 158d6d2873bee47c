"""ListNet loss -- drop-in for reference losses/listnet.py:5-16 (same name, argument order -- y_true FIRST --
and defaults), computed by the streaming sm_100a kernel `ltr_listnet_fwdbwd`."""
from __future__ import annotations

from .. import ops
from ._adapt import back_to, to_device, to_device_2d


def listnetLoss(y_true, y_predicted, apply_sigmoid=False):
    """
        Default ListNet loss: -sum(softmax(y_true) * log(softmax(y_predicted))), softmax over dim 1.
        apply_sigmoid: -sum(sigmoid(softmax(y_true) * log(softmax(y_predicted))))   (listnet.py:13-15)
    :return: 0-dim torch.Tensor attached to autograd
    """
    loss = ops.listnet(to_device_2d(y_true.detach()), to_device_2d(y_predicted), apply_sigmoid)
    return back_to(loss, y_predicted)
