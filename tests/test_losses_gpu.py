"""Every remaining hot-path callable: CUDA path (through the C-ABI and the drop-in callables) vs the golden
vectors generated from the unmodified reference, and vs the oracle on larger seeded inputs."""
import numpy as np
import pytest
import torch

from conftest import golden
from ltr_testing import assert_grad_close, assert_loss_close, make_slates

pytestmark = pytest.mark.gpu
G = golden()


def _fns():
    from ltr_b200.losses.exactNDCG import ndcgLoss
    from ltr_b200.losses.lambdaL import lambdaLoss
    from ltr_b200.losses.listnet import listnetLoss
    from ltr_b200.losses.orderScore import orderScoreLoss
    from ltr_b200.losses.ordinal import ordinalLoss
    from ltr_b200.losses.riskLosses import riskLosses as RL
    return {
        "lambdaLoss": (lambdaLoss, ("y_pred", "y_true")),
        "listnetLoss": (listnetLoss, ("y_true", "y_pred")),
        "orderScoreLoss": (orderScoreLoss, ("y_true", "y_pred")),
        "ndcgLoss": (ndcgLoss, ("y_true", "y_pred")),
        "ordinalLoss": (ordinalLoss, ("y_pred", "y_true")),
        "geoRiskListnetLoss": (RL.geoRiskListnetLoss, ("y_pred", "y_true", "y_baselines")),
        "zRiskListnetLoss": (RL.zRiskListnetLoss, ("y_pred", "y_true", "y_baselines")),
        "tRiskListnetLoss": (RL.tRiskListnetLoss, ("y_pred", "y_true", "y_baselines")),
        "geoRiskLambdaLoss": (RL.geoRiskLambdaLoss, ("y_pred", "y_true", "y_baselines")),
        "zRiskLambdaLoss": (RL.zRiskLambdaLoss, ("y_pred", "y_true", "y_baselines")),
        "tRiskLambdaLoss": (RL.tRiskLambdaLoss, ("y_pred", "y_true", "y_baselines")),
    }


def _oracle_fns():
    from oracle import ltr_oracle as O
    return {
        "lambdaLoss": O.lambda_loss, "listnetLoss": O.listnet_loss, "orderScoreLoss": O.order_score_loss,
        "ndcgLoss": O.ndcg_loss, "ordinalLoss": O.ordinal_loss, "geoRiskListnetLoss": O.geo_risk_listnet_loss,
        "zRiskListnetLoss": O.z_risk_listnet_loss, "tRiskListnetLoss": O.t_risk_listnet_loss,
        "geoRiskLambdaLoss": O.geo_risk_lambda_loss, "zRiskLambdaLoss": O.z_risk_lambda_loss,
        "tRiskLambdaLoss": O.t_risk_lambda_loss,
    }


LOSS_FNS = ("lambdaLoss", "listnetLoss", "orderScoreLoss", "ndcgLoss", "ordinalLoss", "geoRiskListnetLoss",
            "zRiskListnetLoss", "tRiskListnetLoss", "geoRiskLambdaLoss", "zRiskLambdaLoss", "tRiskLambdaLoss")
CASES = [n for n, c in G.cases.items() if c["fn"] in LOSS_FNS]


def _run(fn_name, tensors, kwargs, device="cuda"):
    fn, order = _fns()[fn_name]
    leaf = {k: torch.as_tensor(v).clone().to(device).requires_grad_(k == "y_pred") for k, v in tensors.items()}
    out = fn(*[leaf[k] for k in order if k in leaf], **kwargs)
    grad = None
    if out is not None and out.requires_grad:
        out.sum().backward()
        grad = leaf["y_pred"].grad.cpu().numpy()
    return (None if out is None else out.detach().cpu().numpy()), grad


def _run_oracle(fn_name, tensors, kwargs, dtype=torch.float32):
    fn = _oracle_fns()[fn_name]
    _, order = _fns()[fn_name]
    leaf = {k: torch.as_tensor(v).clone().to(dtype).requires_grad_(k == "y_pred") for k, v in tensors.items()}
    out = fn(*[leaf[k] for k in order if k in leaf], **kwargs)
    grad = None
    if out.requires_grad:
        out.sum().backward()
        grad = leaf["y_pred"].grad.numpy()
    return out.detach().numpy(), grad


@pytest.mark.parametrize("name", CASES)
def test_golden(name):
    case = G.cases[name]
    ins = G.inputs(name)
    out, grad = _run(case["fn"], ins, G.kwargs(name))
    ref = G.out(name)
    assert tuple(out.shape) == tuple(ref.shape), "return shape (0-dim vs [1]) must match the reference"
    o64 = g64 = None
    if "ambda" in case["fn"] or "Risk" in case["fn"]:
        # pairwise / risk values are differences of nearly equal fp32 sums: allow the reference's own fp32 rounding
        o64, g64 = _run_oracle(case["fn"], ins, G.kwargs(name), dtype=torch.float64)
    assert_loss_close(out, ref, atol=2e-6, ref64=o64)
    if G.has_grad(name, "y_pred"):
        assert_grad_close(grad, G.grad(name, "y_pred"), ref64=g64)


# ---------------------------------------------------------------------------------------------------
# larger seeded inputs against the oracle
# ---------------------------------------------------------------------------------------------------
def _check(fn_name, tensors, kwargs, f64_slack=False, **tol):
    out, grad = _run(fn_name, tensors, kwargs)
    rout, rgrad = _run_oracle(fn_name, tensors, kwargs)
    assert tuple(out.shape) == tuple(rout.shape)
    o64, g64 = _run_oracle(fn_name, tensors, kwargs, dtype=torch.float64) if f64_slack else (None, None)
    assert_loss_close(out, rout, ref64=o64, **tol)
    if rgrad is not None:
        assert_grad_close(grad, rgrad, ref64=g64)


@pytest.mark.parametrize("B,n", [(64, 128), (7, 100), (3, 1000), (1, 37), (300, 5)])
@pytest.mark.parametrize("sig", [False, True])
def test_listnet_vs_oracle(B, n, sig):
    s, y = make_slates(B, n, seed=B + n)
    _check("listnetLoss", {"y_true": y, "y_pred": s}, {"apply_sigmoid": sig})


def test_listnet_full_size_properties():
    """BASELINE config 1 shape at 65,536 slates: the batch loss is the sum of the slate losses, gradient rows sum to 0,
    a 512-slate sample agrees with the oracle."""
    from ltr_b200.losses.listnet import listnetLoss
    B, n = 65536, 128
    s, y = make_slates(B, n, seed=5)
    sd = s.cuda().requires_grad_(True)
    yd = y.cuda()
    loss = listnetLoss(yd, sd)
    loss.backward()
    assert torch.equal(loss, listnetLoss(yd, sd.detach())), "batch sum must be bitwise reproducible"
    assert float(sd.grad.sum(dim=1).abs().max()) < 1e-5          # softmax(s) - softmax(y) sums to 0 per slate
    parts = sum(float(listnetLoss(yd[i:i + 8192], sd.detach()[i:i + 8192])) for i in range(0, B, 8192))
    assert abs(parts - float(loss)) <= 1e-5 * abs(float(loss))
    _check("listnetLoss", {"y_true": y[1000:1512], "y_pred": s[1000:1512]}, {})


SCHEMES = [None, "ndcgLoss1_scheme", "ndcgLoss2_scheme", "lamdbaRank_scheme", "ndcgLoss2PP_scheme", "rankNet_scheme",
           "rankNetWeightedByGTDiff_scheme", "rankNetWeightedByGTDiffPowed_scheme"]


@pytest.mark.parametrize("scheme", SCHEMES)
def test_lambda_schemes_vs_oracle(scheme):
    s, y = make_slates(12, 64, seed=21, min_len=20, pad_anywhere=True)
    _check("lambdaLoss", {"y_pred": s, "y_true": y}, {"weighing_scheme": scheme}, f64_slack=True)


@pytest.mark.parametrize("kw", [
    dict(weighing_scheme="ndcgLoss2PP_scheme", k=10),
    dict(weighing_scheme="ndcgLoss2PP_scheme", reduction="mean", reduction_log="natural"),
    dict(weighing_scheme="ndcgLoss1_scheme", k=5, sigma=2.0),
    dict(weighing_scheme="lamdbaRank_scheme", mu=3.0, sigma=0.5, reduction="mean"),
    dict(weighing_scheme=None, eps=1e-4),
])
def test_lambda_options_vs_oracle(kw):
    s, y = make_slates(9, 50, seed=33, min_len=10)
    _check("lambdaLoss", {"y_pred": s, "y_true": y}, kw, f64_slack=True)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("eps,scale,n", [(1e-3, 3.0, 100), (1e-2, 1.0, 200), (0.3, 0.5, 40)])
def test_lambda_active_eps_clamps_take_the_exact_path(scheme, eps, scale, n):
    """The pair tiles run without the two eps clamps of lambdaL.py:51 and hand a slate that WOULD have clamped to the exact
    row-per-thread loop: large eps / wide scores make the clamps active for many pairs of every slate here."""
    s, y = make_slates(7, n, seed=int(eps * 1e4) + n, min_len=n // 2, score_scale=scale)
    _check("lambdaLoss", {"y_pred": s, "y_true": y}, {"weighing_scheme": scheme, "eps": eps}, f64_slack=True)
    _check("lambdaLoss", {"y_pred": s, "y_true": y}, {"weighing_scheme": scheme, "eps": eps, "reduction": "mean", "k": 30}, f64_slack=True)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_lambda_pair_counts_mean_reduction(scheme):
    """reduction="mean" divides by the number of masked pairs, which the fast path takes from label counts (integer labels 0..4),
    from a direct count (fractional labels), or as lim^2 (ndcgLoss1); with and without truncation, with fully padded slates."""
    s, y = make_slates(10, 130, seed=77, min_len=3, pad_anywhere=True)
    y[3] = -1.0                                                   # no doc at all
    y[4, :] = torch.where(y[4] == -1, y[4], torch.full_like(y[4], 2.0))     # one label value: no pair with different labels
    for kw in (dict(), dict(k=17), dict(k=1)):
        _check("lambdaLoss", {"y_pred": s, "y_true": y}, dict(weighing_scheme=scheme, reduction="mean", **kw), f64_slack=True)
    g = torch.Generator().manual_seed(5)
    yf = torch.where(y == -1, y, torch.round(torch.rand(y.shape, generator=g) * 6.0) / 2.0)        # 0, 0.5, ..., 3
    _check("lambdaLoss", {"y_pred": s, "y_true": yf}, dict(weighing_scheme=scheme, reduction="mean"), f64_slack=True)


def test_lambda_config3_shape_and_ties():
    """BASELINE config 3 shape (256-doc slates, ndcgLoss2PP) on a sample, plus a tie-heavy set where the rank
    positions depend on torch.sort's tie order (up to 1.3e-2 relative if the order were different)."""
    s, y = make_slates(16, 256, seed=2020, min_len=128)
    _check("lambdaLoss", {"y_pred": s, "y_true": y}, {"weighing_scheme": "ndcgLoss2PP_scheme"}, f64_slack=True)
    s, y = make_slates(16, 100, seed=8, min_len=40, ties=True)
    for scheme in ("ndcgLoss2PP_scheme", "lamdbaRank_scheme", "ndcgLoss1_scheme"):
        _check("lambdaLoss", {"y_pred": s, "y_true": y}, {"weighing_scheme": scheme}, f64_slack=True)


def test_lambda_errors():
    from ltr_b200.losses.lambdaL import lambdaLoss
    s, y = make_slates(2, 8, seed=1)
    with pytest.raises(KeyError):
        lambdaLoss(s.cuda(), y.cuda(), weighing_scheme="noSuch_scheme")
    with pytest.raises(ValueError):
        lambdaLoss(s.cuda(), y.cuda(), reduction_log="decimal")
    with pytest.raises(ValueError):
        lambdaLoss(s.cuda(), y.cuda(), reduction="median")


@pytest.mark.parametrize("B,n", [(16, 12), (8, 40), (5, 128), (3, 300)])
def test_order_score_and_exact_ndcg_vs_oracle(B, n):
    s, y = make_slates(B, n, seed=n)
    _check("orderScoreLoss", {"y_true": y, "y_pred": s}, {})
    _check("ndcgLoss", {"y_true": y, "y_pred": s}, {})
    st, _ = make_slates(B, n, seed=n + 1, ties=True)               # tied predictions: argsort tie order matters
    _check("ndcgLoss", {"y_true": y, "y_pred": st}, {})


def test_ordinal_vs_oracle_and_padding_error():
    from ltr_b200.losses.ordinal import ordinalLoss
    g = torch.Generator().manual_seed(4)
    y = torch.randint(0, 5, (6, 30), generator=g).float()
    p = torch.rand(6, 30, 4, generator=g) * 0.98 + 0.01
    _check("ordinalLoss", {"y_pred": p, "y_true": y}, {"n": 4})
    y[0, 3] = -1.0
    with pytest.raises(RuntimeError, match="target should be between 0 and 1"):    # what torch's BCELoss raises
        ordinalLoss(p.cuda(), y.cuda(), 4)
    y[0, 3] = 2.0
    p[1, 2, 3] = 1.5
    with pytest.raises(RuntimeError, match="input should be between 0 and 1"):
        ordinalLoss(p.cuda(), y.cuda(), 4)


@pytest.mark.parametrize("fn", ["geoRiskListnetLoss", "zRiskListnetLoss"])
@pytest.mark.parametrize("lt", [1, 2, 3])
@pytest.mark.parametrize("rs", [1, 2, 3])
def test_risk_listnet_vs_oracle(fn, lt, rs):
    g = torch.Generator().manual_seed(10 * lt + rs)
    B, n, S = 96, 128, 2
    s = torch.randn(B, n, generator=g)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    base = torch.randn(B, n, S, generator=g)
    for ideal in (1, 2):
        _check(fn, {"y_pred": s, "y_true": y, "y_baselines": base},
               dict(listnet_transformation=lt, return_strategy=rs, add_ideal_ranking_to_mat=ideal), atol=1e-6,
               f64_slack=True)


@pytest.mark.parametrize("lt", [1, 2, 3])
def test_trisk_listnet_vs_oracle(lt):
    g = torch.Generator().manual_seed(lt)
    B, n = 64, 100
    s = torch.randn(B, n, generator=g)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    base = torch.randn(B, n, generator=g)
    _check("tRiskListnetLoss", {"y_pred": s, "y_true": y, "y_baselines": base}, dict(listnet_transformation=lt), atol=1e-6)


@pytest.mark.parametrize("fn,lt", [("geoRiskLambdaLoss", 1), ("geoRiskLambdaLoss", 2), ("zRiskLambdaLoss", 1),
                                   ("zRiskLambdaLoss", 2), ("tRiskLambdaLoss", 1), ("tRiskLambdaLoss", 2),
                                   ("tRiskLambdaLoss", 3)])
def test_risk_lambda_vs_oracle(fn, lt):
    g = torch.Generator().manual_seed(100 + lt)
    B, n, S = 12, 40, 2
    s = torch.randn(B, n, generator=g)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    if fn.startswith("tRisk"):
        base = torch.randn(B, n, generator=g)
        _check(fn, {"y_pred": s, "y_true": y, "y_baselines": base}, dict(listnet_transformation=lt), atol=1e-6,
               f64_slack=True)
        return
    base = torch.randn(B, n, S, generator=g)
    for rs, ideal, scheme in ((1, 1, "ndcgLoss2PP_scheme"), (2, 2, "ndcgLoss2PP_scheme"), (3, 2, "lamdbaRank_scheme"),
                              (1, 2, None), (1, 1, "ndcgLoss1_scheme")):
        _check(fn, {"y_pred": s, "y_true": y, "y_baselines": base},
               dict(listnet_transformation=lt, return_strategy=rs, add_ideal_ranking_to_mat=ideal, weighing_scheme=scheme),
               atol=1e-6, f64_slack=True)


@pytest.mark.parametrize("name", G.names("lambda_mask_full/"))
def test_lambda_mask_full_golden(name):
    from ltr_b200.losses.lambdaL import lambdaMask
    ins = G.inputs(name)
    out = lambdaMask(torch.from_numpy(ins["y_pred"]).cuda(), torch.from_numpy(ins["y_true"]).cuda(), **G.kwargs(name))
    np.testing.assert_allclose(out.cpu().numpy(), G.out(name), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("scheme", [None, "ndcgLoss1_scheme", "ndcgLoss2PP_scheme"])
def test_lambda_mask_masked_vs_oracle(scheme):
    from ltr_b200.losses.lambdaL import lambdaMask
    from oracle import ltr_oracle as O
    s, y = make_slates(5, 30, seed=44, min_len=12, pad_anywhere=True)
    for kw in (dict(), dict(k=7), dict(reduction_log="natural", sigma=2.0)):
        got = lambdaMask(s.cuda(), y.cuda(), weighing_scheme=scheme, **kw).cpu().numpy()
        ref = O.lambda_mask(s, y, weighing_scheme=scheme, **kw).numpy()
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6)


def test_risk_misc_semantics():
    from ltr_b200.losses.riskLosses import riskLosses as RL
    g = torch.Generator().manual_seed(3)
    s = torch.randn(8, 16, generator=g).cuda()
    y = torch.randint(0, 5, (8, 16), generator=g).float().cuda()
    assert RL.geoRiskListnetLoss(s, y, None, return_strategy=7, add_ideal_ranking_to_mat=2) is None   # riskLosses.py:60
    with pytest.raises(UnboundLocalError):
        RL.geoRiskListnetLoss(s, y, None, listnet_transformation=4)
    out = RL.geoRiskListnetLoss(s, y, None, add_ideal_ranking_to_mat=2, negative=-1)
    assert out.shape == (1,)


# ---------------------------------------------------------------------------------------------------
# risk functions, sorting, NDCG evaluation
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", G.names("georisk_fn/") + G.names("zrisk_fn/"))
def test_risk_functions_golden(name):
    from ltr_b200.losses.riskLosses.riskFunctions import geoRisk, zRisk
    fn = geoRisk if G.cases[name]["fn"] == "geoRisk" else zRisk
    mat = torch.from_numpy(G.inputs(name)["mat"]).cuda()
    out = fn(mat, **G.kwargs(name))
    assert tuple(out.shape) == tuple(G.out(name).shape)
    assert_loss_close(out.cpu().numpy(), G.out(name), rtol=2e-6)


def test_georisk_known_answer_and_gradient():
    """reference tests/georiskTorchTest.py:5-12 -> 0.31438308416523303; gradient w.r.t. the whole matrix vs the oracle."""
    from ltr_b200.losses.riskLosses.riskFunctions import geoRisk
    from ltr_b200.utils.metrics import getGeoRiskDefault
    from oracle import ltr_oracle as O
    known = torch.from_numpy(G.inputs("georisk_fn/known5x8/i0")["mat"])
    assert abs(float(geoRisk(known.cuda(), 3)) - 0.31438308416523303) < 1e-6
    np.testing.assert_allclose(getGeoRiskDefault(known.numpy(), 3), G.out("georisk_default/known5x8"), rtol=1e-6)
    for i in (0, 3, -1):
        m = known.clone().cuda().requires_grad_(True)
        geoRisk(m, 3, i=i).sum().backward()
        mc = known.clone().requires_grad_(True)
        O.geo_risk(mc, 3, i=i).sum().backward()
        assert_grad_close(m.grad.cpu().numpy(), mc.grad.numpy())


@pytest.mark.parametrize("name", G.names("sort/"))
def test_sort_golden(name):
    """torch.sort(descending=True) tie order (std::sort over (key, index)): permutations bit-exact."""
    from ltr_b200 import ops
    keys = torch.from_numpy(G.inputs(name)["keys"])
    vals, perm = ops.sort_desc(keys.cuda(), ops.TIES_STDSORT)
    ref_idx = G.out(name)
    assert np.array_equal(perm.cpu().numpy().astype(np.int64), ref_idx.astype(np.int64))
    assert np.array_equal(vals.cpu().numpy(), np.take_along_axis(keys.numpy(), ref_idx.astype(np.int64), axis=1))


@pytest.mark.parametrize("n", [17, 31, 64, 100, 128, 200, 256, 500, 1024, 2000])
def test_sort_vs_torch_cpu(n):
    """Against torch's CPU sort on this box: heavy ties, -inf runs, NaNs, distinct keys; plus the other tie rules."""
    from ltr_b200 import ops
    g = torch.Generator().manual_seed(n)
    keys = torch.randint(0, 3, (24, n), generator=g).float()
    keys[1] = torch.randn(n, generator=g)
    keys[2, n // 3:] = float("-inf")
    keys[3, ::7] = float("nan")
    keys[4] = 1.0
    keys[5] = torch.randint(0, n // 2, (n,), generator=g).float()
    ref_vals, ref_idx = keys.sort(descending=True, dim=-1)
    vals, perm = ops.sort_desc(keys.cuda(), ops.TIES_STDSORT)
    assert np.array_equal(perm.cpu().numpy(), ref_idx.numpy().astype(np.int32))
    assert np.array_equal(vals.cpu().numpy(), ref_vals.numpy(), equal_nan=True)
    _, stable = ops.sort_desc(keys.cuda(), ops.TIES_STABLE)
    ref_stable = keys.sort(descending=True, dim=-1, stable=True).indices
    assert np.array_equal(stable.cpu().numpy(), ref_stable.numpy().astype(np.int32))
    clean = torch.nan_to_num(keys, nan=5.0, neginf=-9.0)
    _, rev = ops.sort_desc(clean.cuda(), ops.TIES_REVERSED)
    ref_rev = np.argsort(clean.numpy(), axis=1, kind="stable")[:, ::-1]
    assert np.array_equal(rev.cpu().numpy(), ref_rev.astype(np.int32))


@pytest.mark.parametrize("name", G.names("torchndcg/"))
def test_torch_ndcg_golden(name):
    from ltr_b200.utils.metrics import torchNdcg
    ins = G.inputs(name)
    kw = G.kwargs(name)
    out = torchNdcg(torch.from_numpy(ins["ys_true"]).cuda(), torch.from_numpy(ins["ys_pred"]).cuda(), **kw)
    if kw.get("return_type") == "list":
        assert isinstance(out, list)
        out = torch.stack([o.reshape(()) for o in out])
    np.testing.assert_allclose(out.cpu().numpy(), G.out(name), rtol=2e-6, atol=1e-7, equal_nan=True)


@pytest.mark.parametrize("name", G.names("mndcg/"))
def test_mndcg_golden(name):
    from ltr_b200.utils.metrics import mNdcg
    ins = G.inputs(name)
    out = mNdcg(ins["true"], ins["pred"], **G.kwargs(name))
    assert isinstance(out, list) and isinstance(out[0], float)
    np.testing.assert_allclose(np.asarray(out), G.out(name), rtol=1e-12, atol=1e-15)


def test_ndcg_sweep_full_size():
    """BASELINE config 5 evaluation shape (8192 queries x 1024 docs would take the oracle minutes; the sweep is
    checked on a 64-query sample at n = 1024 and by properties at full size)."""
    from ltr_b200.utils.metrics import ndcg_sweep
    from oracle import ltr_oracle as O
    s, y = make_slates(64, 1024, seed=77)
    got = ndcg_sweep(y.cuda(), s.cuda(), ks=(1, 5, 10)).cpu().numpy()
    for j, k in enumerate((1, 5, 10)):
        ref = O.torch_ndcg(y, s, k=k, return_type="tensor").numpy()
        np.testing.assert_allclose(np.nan_to_num(got[:, j]), ref, rtol=2e-6, atol=1e-7)
    Q, n = 8192, 128
    s, y = make_slates(Q, n, seed=78)
    sd, yd = s.cuda(), y.cuda()
    sweep = ndcg_sweep(yd, sd, ks=(1, 5, 10, None))
    assert float(ndcg_sweep(yd, yd + 1e-3 * sd, ks=(10,)).min()) > 0.999999     # ranking by the labels is ideal
    assert float(sweep.max()) <= 1.0 + 1e-6 and float(sweep.min()) >= 0.0
    perm = torch.argsort(torch.rand(Q, n, device="cuda"), dim=1)                 # distinct scores: order-free
    again = ndcg_sweep(yd.gather(1, perm), sd.gather(1, perm), ks=(1, 5, 10, None))
    assert torch.allclose(sweep, again, rtol=1e-6, atol=1e-7)


def test_validation_epilogue_matches_driver_block():
    """main_batch_execution.py:173-200: mNdcg for the model and every baseline system, then geoRisk of the matrix."""
    from ltr_b200.utils.validation import ndcg_matrix, validation_scores
    from oracle import ltr_oracle as O
    g = torch.Generator().manual_seed(55)
    Q, n, S = 200, 60, 3
    y = torch.randint(0, 5, (Q, n), generator=g).float()
    pred = torch.randn(Q, n, 1, generator=g)
    base = torch.randn(Q, n, S, generator=g)
    for k in (5, 10):
        cols = [O.m_ndcg(y.numpy(), pred[:, :, 0].numpy(), k=k)]
        cols += [O.m_ndcg(y.numpy(), base[:, :, i].numpy(), k=k) for i in range(S)]
        ref_mat = torch.stack([torch.tensor(c) for c in cols]).t()
        ref_geo = float(O.geo_risk(ref_mat, 5))
        got_mat = ndcg_matrix(y.cuda(), pred.cuda(), base.cuda(), k=k).cpu().numpy()
        np.testing.assert_allclose(got_mat, np.asarray(cols, dtype=np.float64).T, rtol=1e-12, atol=1e-15)
        mean, geo = validation_scores(y.cuda(), pred.cuda(), base.cuda(), k=k, alpha=5)
        assert abs(mean - float(np.mean(cols[0]))) < 1e-12
        assert abs(geo - ref_geo) <= 2e-6 * abs(ref_geo)


# ---- NDCG@k sweep: every fast-path kernel against the oracle and against the full-sort path ---------------------
@pytest.mark.parametrize("Q,n,ks,kw", [
    (257, 128, (1, 5, 10), {}),                       # label-in-key kernel (rows of one chunk)
    (257, 128, (10,), dict(ties=True)),               # every query marked -> std::sort replay
    (100, 64, (1, 3, 10), {}),                        # fewer than 11 quads with >= 11 docs
    (100, 40, (5, 10), {}),                           # L = -inf: every doc is a candidate
    (50, 8, (1, 5, 10), {}),                          # k > n clips to n
    (33, 4, (1, 2), {}),
    (64, 256, (1, 5, 10), {}),                        # two warps per tile, two chunks
    (64, 1000, (1, 5, 10), {}),                       # four warps per tile, last chunk partial
    (20, 2048, (10,), {}),
    (64, 512, (1, 5, 10), dict(ties=True)),
    (48, 1024, (1, 5, 10), dict(ties=True)),          # tie-only marks: warp-cooperative std::sort replay, ideal order from level counts
    (40, 1000, (10,), dict(ties=True)),
    (30, 300, (1, 5, 10), dict(ties=True)),
    (64, 128, (1, 5, 20), {}),                        # k > 10: warp-per-query tournament kernel
])
def test_ndcg_sweep_kernels_vs_oracle(Q, n, ks, kw):
    from ltr_b200.utils.metrics import ndcg_sweep
    from oracle import ltr_oracle as O
    s, y = make_slates(Q, n, seed=1000 + n, **kw)
    got = ndcg_sweep(y.cuda(), s.cuda(), ks=ks).cpu().numpy()
    for j, k in enumerate(ks):
        ref = O.torch_ndcg(y, s, k=k, return_type="tensor").numpy()
        np.testing.assert_allclose(np.nan_to_num(got[:, j]), ref, rtol=2e-6, atol=1e-7, err_msg=f"k={k}")


def test_ndcg_sweep_special_rows():
    """NaN / inf scores, fractional and large labels, all-zero labels: such queries leave the fast path (marked) and
    must still match the oracle; an exact score tie inside the top 10 next to distinct rows as well."""
    from ltr_b200.utils.metrics import ndcg_sweep
    from oracle import ltr_oracle as O
    s, y = make_slates(96, 128, seed=4242)
    s[1, 5] = float("inf")
    s[2, 9] = float("-inf")
    s[3, 17] = s[3].max()                               # exact tie at rank 0 / 1
    y[4, 3] = 2.5
    y[5, 11] = 7.0
    y[6] = 0.0
    y[7, 0] = 31.0
    s[8, :] = 0.25                                      # all scores equal
    got = ndcg_sweep(y.cuda(), s.cuda(), ks=(1, 5, 10)).cpu().numpy()
    for j, k in enumerate((1, 5, 10)):
        ref = O.torch_ndcg(y, s, k=k, return_type="tensor").numpy()
        np.testing.assert_allclose(np.nan_to_num(got[:, j]), ref, rtol=2e-6, atol=1e-7, err_msg=f"k={k}")


def test_ndcg_sweep_near_ties_and_crowded_quads():
    """Label-in-key kernel (ndcg_keys.cuh): scores that differ only in their last three mantissa bits inside the top
    ranks (the thread redoes such a query with exact compares), three and four of the best docs in ONE quad (second /
    third / rebuilt fourth key), -0 against +0, and rows whose best scores are negative or denormal."""
    from ltr_b200.utils.metrics import ndcg_sweep
    from oracle import ltr_oracle as O
    Q, n = 160, 128
    s, y = make_slates(Q, n, seed=777)
    g = torch.Generator().manual_seed(778)
    for q in range(0, 64):                              # near ties: top docs one to seven ulps apart, labels all differ
        top = float(s[q].max()) + 1.0
        pos = torch.randperm(n, generator=g)[:6]
        vals = torch.tensor([top], dtype=torch.float32).repeat(6)
        bits = vals.view(torch.int32) + torch.randperm(8, generator=g)[:6].to(torch.int32)
        s[q, pos] = bits.view(torch.float32)
        y[q, pos] = torch.tensor([0., 4., 1., 3., 2., 4.])
    for q in range(64, 96):                             # the best docs crowd into one or two quads
        j = int(torch.randint(0, n // 4, (1,), generator=g))
        s[q, 4 * j:4 * j + 4] = float(s[q].max()) + torch.tensor([4., 2., 3., 1.])
        j2 = (j + 5) % (n // 4)
        s[q, 4 * j2:4 * j2 + 3] = float(s[q].max()) + torch.tensor([.5, .25, .75])
    s[96, :] = -s[96].abs() - 1.0                       # all negative
    s[97, :] = s[97] * 1e-41                            # denormals
    s[98, :8] = torch.tensor([0., -0., 0., -0., 1e-45, -1e-45, 0., 0.])
    s[98, 8:] = -1.0
    s[99, :] = s[99].abs() * 1e30                       # large magnitudes
    got = ndcg_sweep(y.cuda(), s.cuda(), ks=(1, 5, 10)).cpu().numpy()
    for j, k in enumerate((1, 5, 10)):
        ref = O.torch_ndcg(y, s, k=k, return_type="tensor").numpy()
        np.testing.assert_allclose(np.nan_to_num(got[:, j]), ref, rtol=2e-6, atol=1e-7, err_msg=f"k={k}")
    # fp64 arithmetic of mNdcg on the same rows (linear and exponential gains)
    from ltr_b200.utils.metrics import mNdcg
    for gains in ("linear", "exponential"):
        ref = np.asarray(O.m_ndcg(y.numpy(), s.numpy(), k=10, gains=gains))
        out = np.asarray(mNdcg(y.numpy(), s.numpy(), k=10, gains=gains))
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12, err_msg=gains)


@pytest.mark.parametrize("n", [512, 640, 1024])
def test_ndcg_sweep_long_rows_special_cases(n):
    """Long rows (four warps per tile of 32 queries, chunks of 128 docs): non-finite scores, odd labels and ties inside the
    first 11 ranks are marked for the full sort, ties at ranks 10 / 11 and beyond are not and must not matter; sorted and
    reverse-sorted rows (every doc / no doc enters the top list); fp64 arithmetic of mNdcg on the same rows."""
    from ltr_b200.utils.metrics import mNdcg, ndcg_sweep
    from oracle import ltr_oracle as O
    Q = 300
    s, y = make_slates(Q, n, seed=5000 + n)
    s[1, 5] = float("inf")
    s[2, n - 3] = float("-inf")
    s[3, 17] = float("nan")
    top = s.topk(13, dim=1)
    s[4, top.indices[4, 1]] = top.values[4, 0]          # tie at ranks 0 / 1
    s[5, top.indices[5, 10]] = top.values[5, 9]         # tie at ranks 9 / 10: decides the top 10
    s[6, top.indices[6, 11]] = top.values[6, 10]        # tie at ranks 10 / 11: immaterial
    s[7, top.indices[7, 12]] = top.values[7, 11]        # tie at ranks 11 / 12
    y[8, 3] = 2.5
    y[9, n - 1] = 7.0
    y[10] = 0.0
    y[11, 0] = -1.0                                     # a padded doc
    s[12, :] = 0.25                                     # all scores equal
    s[13] = torch.arange(n, 0, -1).float()              # already sorted: every doc of the first step enters, then none
    s[14] = torch.arange(n).float()                     # ascending: every doc enters the list
    s[15, : n // 2] = -0.0
    s[15, n // 2:] = 0.0                                # -0 == +0: one big tie
    got = ndcg_sweep(y.cuda(), s.cuda(), ks=(1, 5, 10)).cpu().numpy()
    for j, k in enumerate((1, 5, 10)):
        ref = O.torch_ndcg(y, s, k=k, return_type="tensor").numpy()
        np.testing.assert_allclose(np.nan_to_num(got[:, j]), np.nan_to_num(ref), rtol=2e-6, atol=1e-7, err_msg=f"k={k}")
    keep = [i for i in range(Q) if i not in (1, 2, 3)]                      # (finite scores for the fp64 oracle loop)
    for gains in ("linear", "exponential"):
        ref = np.asarray(O.m_ndcg(y[keep].numpy(), s[keep].numpy(), k=10, gains=gains))
        out = np.asarray(mNdcg(y[keep].numpy(), s[keep].numpy(), k=10, gains=gains))
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12, err_msg=gains)


@pytest.mark.parametrize("Q,n", [(3000, 1024), (777, 256), (100, 2048), (64, 1536), (50, 640)])
def test_ndcg_chunked_long_rows_equal_the_tile_path(Q, n):
    """`ltr_ndcg_at_k_ws` (label-in-key kernel per 128-doc chunk + merge) against `ltr_ndcg_at_k` (tile kernel) on the same
    rows: the same sums in the same order, so the same bits; both against an oracle sample; scratch argument checks."""
    import ctypes
    from ltr_b200 import _lib, ops, _runtime as rt
    from oracle import ltr_oracle as O
    lib = _lib.load()
    s, y = make_slates(Q, n, seed=7000 + n)
    # (the key kernel validates labels and scores per tile of 32 chunks: the special rows sit in different tiles)
    s[0, 5] = float("nan")
    y[10, 7] = 2.5
    top = s.topk(3, dim=1)
    s[20, top.indices[20, 1]] = top.values[20, 0]                 # exact tie at ranks 0 / 1, possibly across chunks
    s[30, n - 1] = s[30].max() + 1.0                              # the best doc is the last one of the last chunk
    sd, yd = s.cuda(), y.cuda()
    ks = (1, 5, 10)
    karr = (ctypes.c_int * 3)(*ks)
    d32, _ = ops._disc_tables(n, sd.device)
    out_a, out_b = torch.zeros(Q, 3, device="cuda"), torch.zeros(Q, 3, device="cuda")
    fl_a, fl_b = torch.zeros(Q + 1, dtype=torch.int32, device="cuda"), torch.zeros(Q + 1, dtype=torch.int32, device="cuda")
    need = int(lib.ltr_ndcg_at_k_scratch_bytes(Q, n))
    assert need == 13 * 4 * Q * (n // 128)
    _lib.check(lib.ltr_ndcg_at_k(yd.data_ptr(), sd.data_ptr(), Q, n, karr, 3, ops.TIES_STDSORT, d32.data_ptr(), out_a.data_ptr(),
                                 fl_a.data_ptr(), rt.stream_ptr()), "ltr_ndcg_at_k")
    scratch = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    _lib.check(lib.ltr_ndcg_at_k_ws(yd.data_ptr(), sd.data_ptr(), Q, n, karr, 3, ops.TIES_STDSORT, d32.data_ptr(), out_b.data_ptr(),
                                    fl_b.data_ptr(), scratch.data_ptr(), need, rt.stream_ptr()), "ltr_ndcg_at_k_ws")
    assert torch.equal(torch.nan_to_num(out_a), torch.nan_to_num(out_b))
    if need:
        assert int(fl_b[1 + 0]) == 1 and int(fl_b[1 + 10]) == 1 and int(fl_b[1 + 20]) == 2 and int(fl_b[1 + 30]) == 0
        assert lib.ltr_ndcg_at_k_ws(yd.data_ptr(), sd.data_ptr(), Q, n, karr, 3, ops.TIES_STDSORT, d32.data_ptr(), out_b.data_ptr(),
                                    fl_b.data_ptr(), scratch.data_ptr(), need - 1, rt.stream_ptr()) != 0      # scratch too small
    m = min(Q, 64)
    for j, k in enumerate(ks):
        ref = O.torch_ndcg(y[:m], s[:m], k=k, return_type="tensor").numpy()
        np.testing.assert_allclose(np.nan_to_num(out_b[:m, j].cpu().numpy()), np.nan_to_num(ref), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("Q,n", [(1000, 128), (600, 512), (300, 1000)])
def test_ndcg_eval_from_host_arrays(Q, n):
    """`ltr_ndcg_at_k_host` through the NdcgEvalHost plan (main.py:145 keeps validation data on the CPU): pinned labels / scores
    in, pinned NDCG out, slots reused; equal to the device sweep bit for bit (rows of whole chunks take the chunked path in both)."""
    from ltr_b200.host import NdcgEvalHost
    from ltr_b200.utils.metrics import ndcg_sweep
    plan = NdcgEvalHost(Q, n, ks=(1, 5, 10), depth=2)
    batches = []
    for i in range(5):                                            # more submissions than slots
        s, y = make_slates(Q, n, seed=8000 + 10 * n + i)
        batches.append((s.pin_memory(), y.pin_memory()))
    for i, (s, y) in enumerate(batches):
        slot, out = plan.submit(y, s)
        plan.wait(slot)
        want = ndcg_sweep(y.cuda(), s.cuda(), ks=(1, 5, 10)).cpu()
        assert torch.equal(torch.nan_to_num(out), torch.nan_to_num(want)), i


def test_ndcg_sweep_large_eval_set_matches_small_batches():
    """More tiles than the label-in-key kernel's resident grid (148 x 4 CTAs x 4 warps): tiles wrap around the grid
    (tile t -> CTA t % grid); the result must equal, bit for bit, the same queries evaluated in slices, and an oracle
    sample (the 2^20-query case with an oracle sample is tests/test_round2_gpu.py)."""
    from ltr_b200.utils.metrics import ndcg_sweep
    Q, n = 90000, 128
    s, y = make_slates(Q, n, seed=99)
    sd, yd = s.cuda(), y.cuda()
    whole = ndcg_sweep(yd, sd, ks=(1, 5, 10))
    parts = torch.cat([ndcg_sweep(yd[i:i + 30000], sd[i:i + 30000], ks=(1, 5, 10)) for i in range(0, Q, 30000)])
    assert torch.equal(torch.nan_to_num(whole), torch.nan_to_num(parts))
    from oracle import ltr_oracle as O
    idx = torch.arange(0, Q, 353)
    ref = np.array([float(v) for v in O.torch_ndcg(y[idx], s[idx], k=10, return_type="list")], dtype=np.float32)
    got = whole[idx.cuda(), 2].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)])
