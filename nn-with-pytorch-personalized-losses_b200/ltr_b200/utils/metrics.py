"""NDCG@k / GeoRisk evaluation -- drop-in for reference utils/metrics.py (torchNdcg :83-104, dcg / ndcg / mNdcg
:48-80, getGeoRiskDefault :8-45).  Sorting, top-k gathers and DCG sums run in `ltr_ndcg_at_k(_f64)`; one call
serves every query (and, through `ndcg_sweep`, every k of a sweep) instead of the reference's Python loops."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..losses._adapt import to_device


def _as_2d(t):
    t = to_device(torch.as_tensor(np.asarray(t)) if not isinstance(t, torch.Tensor) else t.detach())
    if t.dim() == 3 and t.shape[2] == 1:
        t = t.squeeze(2)
    if t.dim() == 1:
        t = t.unsqueeze(0)
    return t.contiguous()


def torchNdcg(ys_true, ys_pred, k=None, return_type='list'):
    """Per-query NDCG@k with gains 2^l - 1 (metrics.py:83-104).  'list': python list of 0-dim fp32 tensors (NaN when
    the ideal DCG is 0); 'tensor': 1-D fp32 tensor with NaN mapped to 0 (:99-102)."""
    yt, yp = _as_2d(ys_true), _as_2d(ys_pred)
    vals = ops.ndcg_at_k(yt, yp, [k], ops.TIES_STDSORT)[:, 0]
    dev = ys_true.device if isinstance(ys_true, torch.Tensor) else torch.device("cpu")
    vals = vals.to(dev)
    if return_type == 'tensor':
        vals = vals.clone()
        vals[torch.isnan(vals)] = 0
        return vals
    return list(vals.unbind(0))


def ndcg_sweep(ys_true, ys_pred, ks=(1, 5, 10), tie_rule=ops.TIES_STDSORT):
    """Extension: NDCG@k for every k of `ks` from one read and one sort of each query -> [Q, len(ks)] (device)."""
    return ops.ndcg_at_k(_as_2d(ys_true), _as_2d(ys_pred), list(ks), tie_rule)


def _f64(true_relevance, pred_relevance, k, gains, use_numpy, empty_value):
    if gains not in ("exponential", "linear"):
        raise ValueError("Invalid gains option.")                                  # metrics.py:63
    yt, yp = _as_2d(true_relevance), _as_2d(pred_relevance)
    # python sorted(reverse=True) is stable (:51); np.argsort(...)[::-1] (:53) is reproduced as the reversed
    # stable order (numpy's default quicksort has no defined tie order)
    tie = ops.TIES_REVERSED if use_numpy else ops.TIES_STABLE
    return ops.ndcg_at_k_f64(yt, yp, [k], tie, gains == "exponential", empty_value)


def dcg(true_relevance, pred_relevance, k=5, gains='linear', use_numpy=False):
    """DCG@k of ONE query (metrics.py:48-66): sum over the first min(k, n) docs in predicted order of
    gain(label) / log2(rank + 1), fp64; gains 'linear' (the default) or 'exponential' (2^l - 1).  Returns a float."""
    if gains not in ("exponential", "linear"):
        raise ValueError("Invalid gains option.")                                  # metrics.py:63
    yt, yp = _as_2d(true_relevance), _as_2d(pred_relevance)
    tie = ops.TIES_REVERSED if use_numpy else ops.TIES_STABLE
    return float(ops.dcg_at_k_f64(yt, yp, [k], tie, gains == "exponential")[0, 0])


def ndcg(true_relevance, pred_relevance, k=5, no_relevant=True, gains='linear', use_numpy=False):
    out = _f64(true_relevance, pred_relevance, k, gains, use_numpy, 1.0 if no_relevant else 0.0)
    return float(out[0, 0])


def mNdcg(true_relevance, pred_relevance, k=5, no_relevant=True, gains='linear', use_numpy=False):
    out = _f64(true_relevance, pred_relevance, k, gains, use_numpy, 1.0 if no_relevant else 0.0)
    return out[:, 0].cpu().tolist()


def getGeoRiskDefault(mat, alpha):
    """GeoRisk of every system column (metrics.py:8-45): numpy fp64 [queries, systems] in, numpy fp64 [systems] out.
    fp64 end to end on the device (`ltr_georisk_default_f64`), including the reference's guard that an entry whose
    expected value is 0 contributes z = 0 (:31-34; all-zero query rows occur with no_relevant=False)."""
    m = torch.as_tensor(np.ascontiguousarray(np.asarray(mat, dtype=np.float64)))
    if m.dim() != 2:
        raise IndexError("tuple index out of range")                               # mat.shape[1] on a 1-D array (:9)
    return ops.georisk_default_f64(m, float(alpha)).cpu().numpy()
