"""The oracle port (oracle/ltr_oracle.py) against golden vectors from the unmodified reference.

CPU only.  This is what pins the oracle: every hot-path callable, every weighing scheme /
transformation / return strategy, the reference's one known-answer test, and torch.sort's
tie order.
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ltr_oracle as O

G = golden()

LOSS_FNS = {
    # golden fn name -> (oracle fn, positional order of the tensor inputs)
    "approxNDCGLoss": (O.approx_ndcg_loss, ("y_pred", "y_true")),
    "lambdaLoss": (O.lambda_loss, ("y_pred", "y_true")),
    "listnetLoss": (O.listnet_loss, ("y_true", "y_pred")),
    "orderScoreLoss": (O.order_score_loss, ("y_true", "y_pred")),
    "ndcgLoss": (O.ndcg_loss, ("y_true", "y_pred")),
    "ordinalLoss": (O.ordinal_loss, ("y_pred", "y_true")),
    "geoRiskListnetLoss": (O.geo_risk_listnet_loss, ("y_pred", "y_true", "y_baselines")),
    "zRiskListnetLoss": (O.z_risk_listnet_loss, ("y_pred", "y_true", "y_baselines")),
    "geoRiskLambdaLoss": (O.geo_risk_lambda_loss, ("y_pred", "y_true", "y_baselines")),
    "zRiskLambdaLoss": (O.z_risk_lambda_loss, ("y_pred", "y_true", "y_baselines")),
    "tRiskListnetLoss": (O.t_risk_listnet_loss, ("y_pred", "y_true", "y_baselines")),
    "tRiskLambdaLoss": (O.t_risk_lambda_loss, ("y_pred", "y_true", "y_baselines")),
}
LOSS_CASES = [n for n, c in G.cases.items() if c["fn"] in LOSS_FNS]


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_matches_reference(name):
    case = G.cases[name]
    fn, order = LOSS_FNS[case["fn"]]
    ins = G.inputs(name)
    leaf = {k: torch.from_numpy(v.copy()).requires_grad_(k == "y_pred") for k, v in ins.items()}
    out = fn(*[leaf[k] for k in order if k in leaf], **G.kwargs(name))
    ref = G.out(name)
    assert tuple(out.shape) == tuple(ref.shape), "return shape (0-dim vs [1]) must match the reference"
    # same torch ops on the same bits: the port is expected to be (nearly) bit-identical
    np.testing.assert_allclose(out.detach().numpy(), ref, rtol=2e-6, atol=1e-7)
    if G.has_grad(name, "y_pred"):
        out.sum().backward()
        gref = G.grad(name, "y_pred")
        scale = max(float(np.abs(gref).max()), 1e-30)
        np.testing.assert_allclose(leaf["y_pred"].grad.numpy() / scale, gref / scale, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("name", G.names("lambda_mask_full/"))
def test_lambda_mask_full(name):
    ins = G.inputs(name)
    out = O.lambda_mask(torch.from_numpy(ins["y_pred"]), torch.from_numpy(ins["y_true"]), **G.kwargs(name))
    np.testing.assert_allclose(out.numpy(), G.out(name), rtol=2e-6, atol=1e-7)


def test_georisk_known_answer():
    """reference tests/georiskTorchTest.py:5-12 -- the only golden value the reference itself holds."""
    mat = torch.from_numpy(G.inputs("georisk_fn/known5x8/i0")["mat"])
    assert abs(float(O.geo_risk(mat, 3)) - 0.31438308416523303) < 5e-8
    np.testing.assert_allclose(O.geo_risk_default(mat.numpy(), 3), G.out("georisk_default/known5x8"), rtol=1e-12)
    assert abs(O.geo_risk_default(mat.numpy().astype(np.float64), 3)[0] - 0.31438308416523303) < 1e-7


@pytest.mark.parametrize("name", G.names("georisk_fn/") + G.names("zrisk_fn/"))
def test_risk_functions(name):
    mat = torch.from_numpy(G.inputs(name)["mat"].copy()).requires_grad_(True)
    fn = O.geo_risk if G.cases[name]["fn"] == "geoRisk" else O.z_risk
    out = fn(mat, requires_grad=True, **G.kwargs(name))
    np.testing.assert_allclose(out.detach().numpy(), G.out(name), rtol=2e-6, atol=1e-7)
    if G.has_grad(name, "mat"):
        out.sum().backward()
        np.testing.assert_allclose(mat.grad.numpy(), G.grad(name, "mat"), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name", G.names("torchndcg/"))
def test_torch_ndcg(name):
    ins = G.inputs(name)
    kw = G.kwargs(name)
    out = O.torch_ndcg(torch.from_numpy(ins["ys_true"]), torch.from_numpy(ins["ys_pred"]), **kw)
    got = out.numpy() if kw["return_type"] == "tensor" else np.array([float(v) for v in out], dtype=np.float32)
    ref = G.out(name)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    # fp32 sequential accumulation restated exactly -> bit-identical
    assert np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)])


@pytest.mark.parametrize("name", G.names("mndcg/"))
def test_m_ndcg(name):
    ins = G.inputs(name)
    out = np.array(O.m_ndcg(ins["true"], ins["pred"], **G.kwargs(name)), dtype=np.float64)
    np.testing.assert_allclose(out, G.out(name), rtol=1e-12, atol=0)


# ---- round-2 goldens (tests/golden/make_golden_v2.py) ----------------------------------------------------
@pytest.mark.parametrize("name", G.names("dcg/"))
def test_dcg(name):
    ins = G.inputs(name)
    out = O.dcg(ins["true"], ins["pred"], **G.kwargs(name))
    np.testing.assert_allclose(out, G.out(name), rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", [n for n in G.names("georisk_default/") if n != "georisk_default/known5x8"])
def test_georisk_default_with_zero_rows(name):
    out = O.geo_risk_default(G.inputs(name)["mat"], **G.kwargs(name))
    np.testing.assert_allclose(out, G.out(name), rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", G.names("lambda_mask/"))
def test_lambda_mask_values_and_gradient(name):
    ins = G.inputs(name)
    sp = torch.from_numpy(ins["y_pred"].copy()).requires_grad_(True)
    out = O.lambda_mask(sp, torch.from_numpy(ins["y_true"]), **G.kwargs(name))
    np.testing.assert_allclose(out.detach().numpy(), G.out(name), rtol=2e-6, atol=1e-7)
    (out * torch.from_numpy(ins["w"])).sum().backward()
    gref = G.grad(name, "y_pred")
    scale = max(float(np.abs(gref).max()), 1e-30)
    np.testing.assert_allclose(sp.grad.numpy() / scale, gref / scale, rtol=1e-4, atol=2e-6)


def test_error_behaviour():
    s = torch.randn(2, 6)
    y = torch.randint(0, 3, (2, 6)).float()
    with pytest.raises(ValueError):
        O.lambda_loss(s, y, reduction="median")
    with pytest.raises(ValueError):
        O.lambda_loss(s, y, reduction_log="decimal")
    with pytest.raises(KeyError):
        O.lambda_loss(s, y, weighing_scheme="nope")
    with pytest.raises(ValueError):
        O.dcg([1, 0], [0.3, 0.2], gains="cubic")
    assert O.geo_risk_listnet_loss(s, y, return_strategy=7) is None
    with pytest.raises(ValueError):       # Normal.cdf validation on NaN
        O.geo_risk(torch.zeros(3, 2), 5)
