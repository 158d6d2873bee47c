"""CPU restatement of the three phases of the risk epilogue (TEST INFRASTRUCTURE ONLY).

numpy fp64 version of what `ltr_risk_phase` (csrc/risk.cu) computes from a shard of the query x system matrix,
with the same 16-double state layout, so that the query-sharded exchange (ltr_b200/sharding.py) can be exercised
on CPU with gloo: per-rank phase-1 / phase-2 states, all-gathered and combined in rank order, must give the value the reference's
riskFunctions.zRisk / geoRisk (losses/riskLosses/riskFunctions.py:4-33) give on the unsharded matrix.
Forward values only; imported by tests/ only.
"""
import math

import numpy as np


def state1(mat):
    mat = np.asarray(mat, dtype=np.float64)
    s = np.zeros(16)
    if mat.shape[0] == 0:                      # an empty shard: the neutral state
        s[0] = -1.0e300
        return s
    s[0] = mat.max()
    s[1] = mat[:, 0].sum()
    s[2] = mat[:, -1].sum()
    s[3] = mat.sum()
    s[4] = mat.shape[0]
    return s


def combine(rows, slot):
    """The ranks' states of one exchange round -> the global state, as the kernels combine them (csrc/risk.cu
    risk_gather): in rank order, entry 0 of round 0 is a MAX (an empty shard contributes -1e300), everything else a SUM."""
    rows = np.asarray(rows, dtype=np.float64)
    out = rows[0].copy()
    for r in rows[1:]:
        if slot == 0:
            out[0] = max(out[0], r[0])
            out[1:] += r[1:]
        else:
            out += r
    return out


def _zcol(mat, g1, last, invert, alpha):
    mat = np.asarray(mat, dtype=np.float64)
    C = mat.shape[1]
    M, Q = g1[0], g1[4]
    X = M - mat if invert else mat
    col = C - 1 if last else 0
    sraw = g1[2] if last else g1[1]
    S = Q * M - sraw if invert else sraw
    N = Q * C * M - g1[3] if invert else g1[3]
    T = X.sum(axis=1)
    e = S * T / N
    z = (X[:, col] - e) / np.sqrt(e)
    c = np.where(z < 0, 1.0 + alpha, 1.0)
    return S, z, c


def state2(mat, g1, invert, alpha, strategy):
    s = np.zeros(16)
    if np.asarray(mat).shape[0] == 0:
        return s
    _, z, c = _zcol(mat, g1, False, invert, alpha)
    s[0] = (z * c).sum()
    if strategy != 1:
        _, z, c = _zcol(mat, g1, True, invert, alpha)
        s[5] = (z * c).sum()
    return s


def final_value(g1, g2, kind, invert, strategy, negative, quirk=False, C=None):
    """kind 0 zRisk, 1 geoRisk.  C = number of columns (only needed with invert)."""
    Q, M = g1[4], g1[0]
    vals = []
    for u, sraw in ((0, g1[1]), (1, g1[2])):
        S = Q * M - sraw if invert else sraw
        zsum = g2[5 * u]
        if kind == 0:
            vals.append(zsum)
        else:
            cdf = 0.5 * math.erfc(-(zsum / Q) / math.sqrt(2.0))
            vals.append(math.sqrt(S / Q * cdf))
    f = negative
    if strategy == 1:
        return f * vals[0]
    if strategy == 2:
        return f * vals[1] - vals[0] if quirk else f * (vals[1] - vals[0])
    return f * (vals[1] - vals[0]) ** 2
