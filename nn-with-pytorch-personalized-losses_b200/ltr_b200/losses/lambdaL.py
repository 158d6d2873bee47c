"""LambdaLoss framework -- drop-in for reference losses/lambdaL.py (lambdaMask :7-64, lambdaLoss :67-93, the weighing
schemes :96-127; same names, argument order, defaults and exceptions).  lambdaLoss runs the fused sm_100a kernel
`ltr_lambda_fwdbwd`; lambdaMask the pair-matrix kernels `ltr_lambda_pair_matrix(_bwd)`."""
from __future__ import annotations

import torch

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


def _k_arg(k):
    """k=None -> -1 (no truncation); k=0 stays 0: the reference's `[:0, :0]` mask selects no pair (lambdaL.py:29-30)."""
    return -1 if k is None else int(k)


def lambdaMask(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE, weighing_scheme=None, k=None,
               sigma=1., mu=10., reduction="sum", reduction_log="binary", return_losses=False):
    """
    The pair log-probabilities lambdaLoss reduces (lambdaL.py:7-64): the 1-D tensor of the masked pairs, in
    (slate, rank_i, rank_j) order, or -- return_losses=True -- the unmasked [batch, slate, slate] tensor.
    Like the reference's, the result is attached to autograd: d / d y_pred flows through the score differences
    (`ltr_lambda_pair_matrix_bwd`), the weights carry no gradient; the boolean gather is index plumbing on the device.
    """
    scheme = _scheme_id(weighing_scheme)
    if reduction_log not in ("natural", "binary"):
        raise ValueError("Reduction logarithm base can be either natural or binary")
    logp, mask = ops.lambda_pair_matrix(to_device(y_pred), to_device(y_true.detach()), eps,
                                        float(padded_value_indicator), scheme, _k_arg(k), sigma, mu,
                                        reduction_log == "natural", not return_losses)
    if return_losses:
        return back_to(logp, y_pred)
    return back_to(logp[mask.bool()], y_pred)


def lambdaLoss(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE, weighing_scheme=None, k=None,
               sigma=1., mu=10., reduction="sum", reduction_log="binary", process_group=None):
    """LambdaLoss (Wang et al., "The LambdaLoss Framework for Ranking Metric Optimization"): -sum / -mean over the
    pairs (i, j) of real docs with label_i > label_j and both rank positions < k of log_b(sigmoid(sigma (s_i - s_j))^w_ij),
    w_ij given by `weighing_scheme` (None, ndcgLoss1_scheme, ndcgLoss2_scheme, lamdbaRank_scheme, ndcgLoss2PP_scheme
    with weight `mu`, rankNet_scheme, rankNetWeightedByGTDiff_scheme, rankNetWeightedByGTDiffPowed_scheme).
    y_pred, y_true: [batch, slate]; labels equal to `padded_value_indicator` mark padding; `reduction` is "sum" or
    "mean", `reduction_log` "binary" or "natural".  Returns a 0-dim tensor attached to autograd.
    `process_group` (extension, query-sharded runs): with reduction="mean" the reference divides by the
    number of masked pairs of the WHOLE batch (lambdaL.py:88-89); every rank then returns its own sum divided by the
    all-reduced global pair count, so that the ranks' values (and gradients) add up to the single-process ones.
    With reduction="sum" the ranks' values add up as they are (under DistributedDataParallel, which AVERAGES
    gradients, scale the loss by the world size in both cases).
    """
    scheme = _scheme_id(weighing_scheme)
    if reduction_log not in ("natural", "binary"):
        raise ValueError("Reduction logarithm base can be either natural or binary")      # lambdaL.py:57
    if reduction not in ("sum", "mean"):
        raise ValueError("Reduction method can be either sum or mean")                    # lambdaL.py:91
    from .. import sharding
    if reduction == "mean" and sharding.is_sharded(process_group):
        total, pairs = ops.lambda_loss_and_count(to_device(y_pred), to_device(y_true.detach()), eps,
                                                 float(padded_value_indicator), scheme, _k_arg(k), sigma, mu,
                                                 reduction_log == "natural")
        return back_to(total / sharding.global_sum(pairs, process_group), y_pred)
    loss = ops.lambda_loss(to_device(y_pred), to_device(y_true.detach()), eps, float(padded_value_indicator), scheme,
                           _k_arg(k), sigma, mu, reduction == "mean", reduction_log == "natural")
    return back_to(loss, y_pred)


# ---------------------------------------------------------------------------------------------------------------------
# The weighing schemes as callables (lambdaL.py:96-127).  lambdaLoss / lambdaMask select a scheme BY NAME and evaluate
# it inside the kernels (csrc/lambda_common.cuh: pair_weight); these functions exist for callers that import them and
# build a weight table themselves.  G: [batch, slate] gains in rank order, D: [1, slate] discounts log2(1 + rank),
# args = (mu, labels in rank order clamped at 0).  Broadcast arithmetic on the tensors' own device.
# ---------------------------------------------------------------------------------------------------------------------
def _pairwise_abs_diff(t):
    return (t.unsqueeze(2) - t.unsqueeze(1)).abs()


def ndcgLoss1_scheme(G, D, *args):
    return (G / D).unsqueeze(2)                                   # weight of the FIRST doc of the pair only


def ndcgLoss2_scheme(G, D, *args):
    n = G.shape[1]
    inv = D[0].abs().reciprocal()                                 # 1 / log2(1 + r), r = 1..n
    rank = torch.arange(n, device=G.device)
    gap = (rank.unsqueeze(1) - rank.unsqueeze(0)).abs()           # |i - j|
    delta = (inv[gap - 1] - inv[gap]).abs()                       # |1/log2(1 + gap) - 1/log2(2 + gap)|
    delta = delta.masked_fill(gap == 0, 0.)                       # the diagonal's gap - 1 = -1 wraps around: zeroed (:104)
    return delta.unsqueeze(0) * _pairwise_abs_diff(G)


def lamdbaRank_scheme(G, D, *args):
    return _pairwise_abs_diff(D.reciprocal().expand_as(G)) * _pairwise_abs_diff(G)


def ndcgLoss2PP_scheme(G, D, *args):
    return args[0] * ndcgLoss2_scheme(G, D) + lamdbaRank_scheme(G, D)


def rankNet_scheme(G, D, *args):
    return 1.


def rankNetWeightedByGTDiff_scheme(G, D, *args):
    return _pairwise_abs_diff(args[1])


def rankNetWeightedByGTDiffPowed_scheme(G, D, *args):
    return _pairwise_abs_diff(args[1] ** 2)
