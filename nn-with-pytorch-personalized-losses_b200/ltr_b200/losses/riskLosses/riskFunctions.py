"""zRisk / geoRisk -- drop-in for reference losses/riskLosses/riskFunctions.py:4-33, computed by the three
phases of `ltr_risk_phase` (fp64 reductions on the device)."""
from __future__ import annotations

import torch

from ... import ops
from .._adapt import back_to, to_device


def _col_first(mat, i):
    """Column i to the front: z-risk only distinguishes 'the column' from 'all columns' (row sums, N)."""
    C = mat.shape[1]
    i = i % C
    if i == 0:
        return mat
    idx = [i] + [c for c in range(C) if c != i]
    return mat[:, idx]


def zRisk(mat, alpha, requires_grad=False, i=0):
    m = _col_first(to_device(mat), i).contiguous()
    return back_to(ops.risk_fn(m, ops.KIND_Z, float(alpha)).reshape(()), mat)


def geoRisk(mat, alpha, requires_grad=False, i=0):
    m = _col_first(to_device(mat), i).contiguous()
    out = ops.risk_fn(m, ops.KIND_GEO, float(alpha)).reshape(1)     # [1]: Normal(tensor([0.]), tensor([1.])).cdf broadcasts
    if bool(torch.isnan(out).any()):
        # torch.distributions.Normal.cdf validates its argument in the reference (riskFunctions.py:31-32)
        raise ValueError("Expected value argument (Tensor of shape ()) to be within the support (Real()) of the "
                         "distribution Normal(loc: 0.0, scale: 1.0), but found invalid values:\ntensor(nan)")
    return back_to(out, mat)
