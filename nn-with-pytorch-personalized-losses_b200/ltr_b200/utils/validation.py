"""Validation epilogue on the device -- the per-epoch block of the reference driver (main_batch_execution.py:173-200):

    ndcg_score = mNdcg(y_vali, pred, k)                       # model
    mat = [ndcg_score] + [mNdcg(y_vali, y_baseline[:, :, i], k) for every baseline system]
    georisk = geoRisk(stack(mat).t(), 5)

The reference makes (S + 1) * Q Python-level sorts on `.numpy()` copies.  Here the scores stay on the GPU:
one `ltr_ndcg_at_k_f64` launch per system fills a column of the query x system matrix and the three
`ltr_risk_phase` launches reduce it.  SURVEY 8f row 2 ("next": the direct consumer of K4 + K6b).
"""
from __future__ import annotations

import torch

from .. import ops
from ..losses._adapt import to_device


def ndcg_matrix(y_true, y_pred, y_baselines=None, k=5, no_relevant=True, gains="linear", use_numpy=False):
    """[Q, 1 + S] fp64 matrix of per-query NDCG@k: column 0 the model, then the S baseline systems
    (`y_baselines` is [Q, n, S]).  Same options and tie rules as `utils.metrics.mNdcg`."""
    if gains not in ("exponential", "linear"):
        raise ValueError("Invalid gains option.")
    yt = to_device(y_true.detach())
    yp = to_device(y_pred.detach())
    if yp.dim() == 3 and yp.shape[2] == 1:
        yp = yp.squeeze(2).contiguous()
    systems = [yp]
    if y_baselines is not None:
        yb = to_device(y_baselines.detach())
        systems += [yb[:, :, i].contiguous() for i in range(yb.shape[2])]
    tie = ops.TIES_REVERSED if use_numpy else ops.TIES_STABLE
    cols = [ops.ndcg_at_k_f64(yt, s, [k], tie, gains == "exponential", 1.0 if no_relevant else 0.0)[:, 0]
            for s in systems]
    return torch.stack(cols, dim=1)


def validation_scores(y_true, y_pred, y_baselines=None, k=5, alpha=5, **ndcg_options):
    """(mean NDCG@k of the model, GeoRisk of the model against the baselines), both python floats.
    Like the driver, the NDCG list becomes an fp32 matrix before geoRisk (torch.tensor(list_of_floats))."""
    mat = ndcg_matrix(y_true, y_pred, y_baselines, k=k, **ndcg_options)
    ndcg_mean = float(mat[:, 0].mean())
    geo = ops.risk_fn(mat.float().contiguous(), ops.KIND_GEO, float(alpha))
    if bool(torch.isnan(geo).any()):
        raise ValueError("geoRisk: NaN z-risk (an expected value was 0)")
    return ndcg_mean, float(geo)
