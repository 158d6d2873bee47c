"""LambdaLoss framework -- drop-in for reference losses/lambdaL.py:67-93 (lambdaLoss; same name, argument
order, defaults and exceptions), computed by the fused sm_100a kernel `ltr_lambda_fwdbwd`."""
from __future__ import annotations

from .. import ops
from ._adapt import back_to, to_device

PADDED_Y_VALUE = -1
DEFAULT_EPS = 1e-10


def _scheme_id(weighing_scheme):
    if weighing_scheme is None:
        return 0
    name = weighing_scheme if isinstance(weighing_scheme, str) else getattr(weighing_scheme, "__name__", weighing_scheme)
    if name not in ops.LAMBDA_SCHEMES:
        raise KeyError(name)             # the reference resolves the name with globals()[...] (lambdaL.py:46)
    return ops.LAMBDA_SCHEMES[name]


def lambdaMask(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE, weighing_scheme=None, k=None,
               sigma=1., mu=10., reduction="sum", reduction_log="binary", return_losses=False):
    """
    The pair log-probabilities lambdaLoss reduces (lambdaL.py:7-64): the 1-D tensor of the masked pairs, in
    (slate, rank_i, rank_j) order, or -- return_losses=True -- the unmasked [batch, slate, slate] tensor.
    Forward only (the fused training path is lambdaLoss); the boolean gather is index plumbing on the device.
    """
    scheme = _scheme_id(weighing_scheme)
    if reduction_log not in ("natural", "binary"):
        raise ValueError("Reduction logarithm base can be either natural or binary")
    logp, mask = ops.lambda_pair_matrix(to_device(y_pred.detach()), to_device(y_true.detach()), eps,
                                        float(padded_value_indicator), scheme, 0 if k is None else int(k), sigma, mu,
                                        reduction_log == "natural", not return_losses)
    if return_losses:
        return back_to(logp, y_pred)
    return back_to(logp[mask.bool()], y_pred)


def lambdaLoss(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE, weighing_scheme=None, k=None,
               sigma=1., mu=10., reduction="sum", reduction_log="binary"):
    """LambdaLoss (Wang et al., "The LambdaLoss Framework for Ranking Metric Optimization"): -sum / -mean over the
    pairs (i, j) of real docs with label_i > label_j and both rank positions < k of log_b(sigmoid(sigma (s_i - s_j))^w_ij),
    w_ij given by `weighing_scheme` (None, ndcgLoss1_scheme, ndcgLoss2_scheme, lamdbaRank_scheme, ndcgLoss2PP_scheme
    with weight `mu`, rankNet_scheme, rankNetWeightedByGTDiff_scheme, rankNetWeightedByGTDiffPowed_scheme).
    y_pred, y_true: [batch, slate]; labels equal to `padded_value_indicator` mark padding; `reduction` is "sum" or
    "mean", `reduction_log` "binary" or "natural".  Returns a 0-dim tensor attached to autograd.
    """
    scheme = _scheme_id(weighing_scheme)
    if reduction_log not in ("natural", "binary"):
        raise ValueError("Reduction logarithm base can be either natural or binary")      # lambdaL.py:57
    if reduction not in ("sum", "mean"):
        raise ValueError("Reduction method can be either sum or mean")                    # lambdaL.py:91
    loss = ops.lambda_loss(to_device(y_pred), to_device(y_true.detach()), eps, float(padded_value_indicator), scheme,
                           0 if k is None else int(k), sigma, mu, reduction == "mean", reduction_log == "natural")
    return back_to(loss, y_pred)


# names of the weighing schemes, importable like the reference's functions (they are selected by name)
def ndcgLoss1_scheme(*_a, **_k):
    raise NotImplementedError("weighing schemes are evaluated inside the fused kernel; pass the name to lambdaLoss")


ndcgLoss2_scheme = lamdbaRank_scheme = ndcgLoss2PP_scheme = rankNet_scheme = ndcgLoss1_scheme
rankNetWeightedByGTDiff_scheme = rankNetWeightedByGTDiffPowed_scheme = ndcgLoss1_scheme
