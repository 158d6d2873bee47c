"""Risk-sensitive losses -- drop-in for reference losses/riskLosses/riskLosses.py: the ListNet family
(geoRiskListnetLoss :8-60, zRiskListnetLoss :128-180, tRiskListnetLoss :247-291) and the LambdaLoss family
(geoRiskLambdaLoss :63-125, zRiskLambdaLoss :183-244, tRiskLambdaLoss :294-345).  Same names, argument order,
defaults, return shape ([1]) and quirks; computed by `ltr_risk_{listnet,lambda}_rows` / `ltr_risk_phase` /
`ltr_risk_{listnet,lambda}_rows_bwd`.  `process_group` is an extension: when the queries are sharded over ranks the
matrix-wide max, the column totals and the z sums are all-reduced over NCCL between the phases."""
from __future__ import annotations

import torch

from ... import ops
from .._adapt import back_to, to_device
from ..lambdaL import _scheme_id


def _prep(y_predicted, y_true, y_baselines):
    yp = to_device(y_predicted)
    yt = to_device(y_true.detach())
    yb = None
    if y_baselines is not None:
        yb = to_device(y_baselines.detach())
        if yb.dim() == 2:
            yb = yb.unsqueeze(2).contiguous()
    return yp, yt, yb


def _risk_listnet(kind, quirk, y_predicted, y_true, y_baselines, alpha, lt, return_strategy, negative, add_ideal,
                  process_group):
    if lt not in (1, 2, 3):
        raise UnboundLocalError("local variable 'mat' referenced before assignment")   # riskLosses.py:43
    yp, yt, yb = _prep(y_predicted, y_true, y_baselines)
    if return_strategy not in (1, 2, 3):
        return None                                                                    # riskLosses.py:60
    loss = ops.risk_listnet(yp, yt, yb, kind, float(alpha), int(lt), int(return_strategy), float(negative),
                            add_ideal == 2, quirk, process_group)
    if kind == ops.KIND_GEO and bool(torch.isnan(loss).any()):
        raise ValueError("Expected value argument (Tensor of shape ()) to be within the support (Real()) of the "
                         "distribution Normal(loc: 0.0, scale: 1.0), but found invalid values:\ntensor(nan)")
    return back_to(loss, y_predicted)


def geoRiskListnetLoss(y_predicted, y_true, y_baselines=None, alpha=5, listnet_transformation=1, return_strategy=1,
                       negative=1, add_ideal_ranking_to_mat=1, process_group=None):
    return _risk_listnet(ops.KIND_GEO, False, y_predicted, y_true, y_baselines, alpha, listnet_transformation,
                         return_strategy, negative, add_ideal_ranking_to_mat, process_group)


def zRiskListnetLoss(y_predicted, y_true, y_baselines=None, alpha=5, listnet_transformation=1, return_strategy=1,
                     negative=1, add_ideal_ranking_to_mat=1, process_group=None):
    # return_strategy 2 is `factor * zRisk(i=-1) - zRisk(i=0)` here (operator precedence, riskLosses.py:176)
    return _risk_listnet(ops.KIND_Z, True, y_predicted, y_true, y_baselines, alpha, listnet_transformation,
                         return_strategy, negative, add_ideal_ranking_to_mat, process_group)


def tRiskListnetLoss(y_predicted, y_true, y_baselines, alpha=5, listnet_transformation=1, negative=1,
                     process_group=None):
    if listnet_transformation not in (1, 2, 3):
        raise RuntimeError("stack expects a non-empty TensorList")
    yp, yt, yb = _prep(y_predicted, y_true, y_baselines)
    loss = ops.risk_listnet(yp, yt, yb, ops.KIND_T, float(alpha), int(listnet_transformation), 1, float(negative),
                            False, False, process_group)
    return back_to(loss, y_predicted)


def _risk_lambda(kind, y_predicted, y_true, y_baselines, alpha, lt, return_strategy, negative, add_ideal,
                 weighing_scheme, process_group):
    scheme = _scheme_id(weighing_scheme)
    if lt not in (1, 2):
        raise UnboundLocalError("local variable 'mat' referenced before assignment")   # riskLosses.py:109
    yp, yt, yb = _prep(y_predicted, y_true, y_baselines)
    if return_strategy not in (1, 2, 3):
        return None
    loss = ops.risk_lambda(yp, yt, yb, kind, float(alpha), int(lt), int(return_strategy), float(negative),
                           add_ideal == 2, scheme, process_group)
    if kind == ops.KIND_GEO and bool(torch.isnan(loss).any()):
        raise ValueError("Expected value argument (Tensor of shape ()) to be within the support (Real()) of the "
                         "distribution Normal(loc: 0.0, scale: 1.0), but found invalid values:\ntensor(nan)")
    return back_to(loss, y_predicted)


def geoRiskLambdaLoss(y_predicted, y_true, y_baselines=None, alpha=5, listnet_transformation=1, return_strategy=1,
                      negative=1, add_ideal_ranking_to_mat=1, weighing_scheme="ndcgLoss2PP_scheme", process_group=None):
    return _risk_lambda(ops.KIND_GEO, y_predicted, y_true, y_baselines, alpha, listnet_transformation, return_strategy,
                        negative, add_ideal_ranking_to_mat, weighing_scheme, process_group)


def zRiskLambdaLoss(y_predicted, y_true, y_baselines=None, alpha=5, listnet_transformation=1, return_strategy=1,
                    negative=1, add_ideal_ranking_to_mat=1, weighing_scheme="ndcgLoss2PP_scheme", process_group=None):
    return _risk_lambda(ops.KIND_Z, y_predicted, y_true, y_baselines, alpha, listnet_transformation, return_strategy,
                        negative, add_ideal_ranking_to_mat, weighing_scheme, process_group)


def tRiskLambdaLoss(y_predicted, y_true, y_baselines, alpha=5, listnet_transformation=1, negative=1,
                    weighing_scheme="ndcgLoss2PP_scheme", process_group=None):
    scheme = _scheme_id(weighing_scheme)
    if listnet_transformation not in (1, 2, 3):
        raise RuntimeError("stack expects a non-empty TensorList")
    yp, yt, yb = _prep(y_predicted, y_true, y_baselines)
    loss = ops.risk_lambda(yp, yt, yb, ops.KIND_T, float(alpha), int(listnet_transformation), 1, float(negative),
                           False, scheme, process_group)
    return back_to(loss, y_predicted)
