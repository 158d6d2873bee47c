"""Generate golden vectors by running the UNMODIFIED reference on seeded inputs.

Run in the build container only (the reference cannot travel to the GPU box):

    cd /tmp && PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 \
        python /root/repo/tests/golden/make_golden.py

Writes `tests/golden/golden_v1.npz` (inputs + reference outputs, a few hundred KB) and
`tests/golden/golden_v1.json` (the manifest: which reference callable, which kwargs).
Nothing in `tests/`, `bench.py` or `smoke()` imports the reference at run time; they read
these two files.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")

from losses.approxNDCG import approxNDCGLoss            # noqa: E402  (reference, via PYTHONPATH)
from losses.exactNDCG import ndcgLoss                   # noqa: E402
from losses.lambdaL import lambdaLoss, lambdaMask       # noqa: E402
from losses.listnet import listnetLoss                  # noqa: E402
from losses.orderScore import orderScoreLoss            # noqa: E402
from losses.ordinal import ordinalLoss                  # noqa: E402
from losses.riskLosses import riskLosses as RL          # noqa: E402
from losses.riskLosses.riskFunctions import geoRisk, zRisk  # noqa: E402
from utils.metrics import getGeoRiskDefault, mNdcg, torchNdcg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 2020          # main_batch_execution.py:21

arrays = {}
manifest = []


def slate_batch(g, B, n, min_len=None, ties=False, pad_anywhere=False, all_zero_row=False, all_pad_row=False):
    """labels U{0..4} as fp32, scores N(0,1) (or U{0..5}/3 for the tie-stress set), -1 padding."""
    y = torch.randint(0, 5, (B, n), generator=g).float()
    if ties:
        s = torch.randint(0, 6, (B, n), generator=g).float() / 3.0
    else:
        s = torch.randn(B, n, generator=g)
    if min_len is not None:
        for b in range(B):
            L = int(torch.randint(min_len, n + 1, (1,), generator=g))
            if pad_anywhere:
                drop = torch.randperm(n, generator=g)[: n - L]
                y[b, drop] = -1.0
            else:
                y[b, L:] = -1.0
    if all_zero_row:
        y[0] = torch.where(y[0] >= 0, torch.zeros_like(y[0]), y[0])
    if all_pad_row and B > 1:
        y[1] = -1.0
    return s, y


def record(name, fn_name, inputs, kwargs, value, grads):
    entry = {"name": name, "fn": fn_name, "kwargs": kwargs, "inputs": list(inputs.keys()),
             "grads": list(grads.keys())}
    for k, v in inputs.items():
        arrays[f"{name}/in/{k}"] = np.asarray(v)
    arrays[f"{name}/out"] = np.asarray(value)
    for k, v in grads.items():
        arrays[f"{name}/grad/{k}"] = np.asarray(v)
    manifest.append(entry)


def run_loss(name, fn_name, fn, tensors, order, kwargs, grad_wrt=("y_pred",)):
    """tensors: dict name->tensor; order: positional order of names for the reference call."""
    leaf = {k: v.clone().requires_grad_(k in grad_wrt) for k, v in tensors.items()}
    out = fn(*[leaf[k] for k in order], **kwargs)
    grads = {}
    if out.requires_grad:
        out.sum().backward()
        for k in grad_wrt:
            if leaf[k].grad is not None:
                grads[k] = leaf[k].grad.detach().numpy()
    record(name, fn_name, {k: v.numpy() for k, v in tensors.items()}, kwargs, out.detach().numpy(), grads)


def main():
    g = torch.Generator().manual_seed(SEED)

    # ---------------- approxNDCG -------------------------------------------------------
    shapes = [
        ("b6n24_pad", dict(B=6, n=24, min_len=10)),
        ("b4n40_nopad", dict(B=4, n=40)),
        ("b5n33_padany", dict(B=5, n=33, min_len=8, pad_anywhere=True)),
        ("b4n20_ties", dict(B=4, n=20, min_len=9, ties=True)),
        ("b3n16_zero_allpad", dict(B=3, n=16, min_len=6, all_zero_row=True, all_pad_row=True)),
        ("b1n12", dict(B=1, n=12, min_len=5)),
        ("b2n130", dict(B=2, n=130, min_len=70)),
    ]
    for tag, kw in shapes:
        s, y = slate_batch(g, **kw)
        for alpha in (1.0, 10.0):
            run_loss(f"approx/{tag}/a{int(alpha)}", "approxNDCGLoss", approxNDCGLoss,
                     {"y_pred": s, "y_true": y}, ("y_pred", "y_true"), {"alpha": alpha})
    s, y = slate_batch(g, B=4, n=24, min_len=10)
    run_loss("approx/big_scores/a10", "approxNDCGLoss", approxNDCGLoss,
             {"y_pred": s * 30.0, "y_true": y}, ("y_pred", "y_true"), {"alpha": 10.0})
    run_loss("approx/eps1e-3/a10", "approxNDCGLoss", approxNDCGLoss,
             {"y_pred": s, "y_true": y}, ("y_pred", "y_true"), {"alpha": 10.0, "eps": 1e-3})

    # ---------------- LambdaLoss -------------------------------------------------------
    schemes = [None, "ndcgLoss1_scheme", "ndcgLoss2_scheme", "lamdbaRank_scheme", "ndcgLoss2PP_scheme",
               "rankNet_scheme", "rankNetWeightedByGTDiff_scheme", "rankNetWeightedByGTDiffPowed_scheme"]
    s, y = slate_batch(g, B=5, n=28, min_len=12)
    for sch in schemes:
        run_loss(f"lambda/b5n28/{sch}", "lambdaLoss", lambdaLoss, {"y_pred": s, "y_true": y},
                 ("y_pred", "y_true"), {"weighing_scheme": sch})
    for sch in ("ndcgLoss2PP_scheme", "ndcgLoss1_scheme", None):
        run_loss(f"lambda/b5n28_k5/{sch}", "lambdaLoss", lambdaLoss, {"y_pred": s, "y_true": y},
                 ("y_pred", "y_true"), {"weighing_scheme": sch, "k": 5})
        run_loss(f"lambda/b5n28_mean_nat/{sch}", "lambdaLoss", lambdaLoss, {"y_pred": s, "y_true": y},
                 ("y_pred", "y_true"), {"weighing_scheme": sch, "reduction": "mean",
                                        "reduction_log": "natural", "sigma": 2.0, "mu": 3.0})
    s, y = slate_batch(g, B=3, n=18, min_len=7, all_zero_row=True, all_pad_row=True)
    for sch in ("ndcgLoss2PP_scheme", "ndcgLoss1_scheme", "lamdbaRank_scheme"):
        run_loss(f"lambda/b3n18_zero_allpad/{sch}", "lambdaLoss", lambdaLoss, {"y_pred": s, "y_true": y},
                 ("y_pred", "y_true"), {"weighing_scheme": sch})
    s, y = slate_batch(g, B=4, n=40, min_len=20, ties=True)      # n > 16: std::sort tie order is exercised
    for sch in ("ndcgLoss2PP_scheme", "ndcgLoss1_scheme", "lamdbaRank_scheme", "ndcgLoss2_scheme", None):
        run_loss(f"lambda/b4n40_ties/{sch}", "lambdaLoss", lambdaLoss, {"y_pred": s, "y_true": y},
                 ("y_pred", "y_true"), {"weighing_scheme": sch})
    s, y = slate_batch(g, B=3, n=24, min_len=10)
    run_loss("lambda/big_scores/ndcgLoss2PP_scheme", "lambdaLoss", lambdaLoss,
             {"y_pred": s * 40.0, "y_true": y}, ("y_pred", "y_true"), {"weighing_scheme": "ndcgLoss2PP_scheme"})
    s, y = slate_batch(g, B=2, n=20)
    for sch in ("ndcgLoss2PP_scheme", "lamdbaRank_scheme"):
        out = lambdaMask(s, y, weighing_scheme=sch, return_losses=True)
        record(f"lambda_mask_full/b2n20/{sch}", "lambdaMask", {"y_pred": s.numpy(), "y_true": y.numpy()},
               {"weighing_scheme": sch, "return_losses": True}, out.numpy(), {})

    # ---------------- ListNet ----------------------------------------------------------
    for tag, (B, n) in {"b6n24": (6, 24), "b1n10": (1, 10), "b3n128": (3, 128)}.items():
        s, y = slate_batch(g, B=B, n=n)
        run_loss(f"listnet/{tag}", "listnetLoss", listnetLoss, {"y_true": y, "y_pred": s},
                 ("y_true", "y_pred"), {})
        run_loss(f"listnet/{tag}/sigmoid", "listnetLoss", listnetLoss, {"y_true": y, "y_pred": s},
                 ("y_true", "y_pred"), {"apply_sigmoid": True})

    # ---------------- orderScore / exact NDCG / ordinal --------------------------------
    for tag, (B, n) in {"b4n12": (4, 12), "b3n40": (3, 40)}.items():
        s, y = slate_batch(g, B=B, n=n)
        run_loss(f"orderscore/{tag}", "orderScoreLoss", orderScoreLoss, {"y_true": y, "y_pred": s},
                 ("y_true", "y_pred"), {})
        run_loss(f"exactndcg/{tag}", "ndcgLoss", ndcgLoss, {"y_true": y, "y_pred": s},
                 ("y_true", "y_pred"), {})
    for tag, (B, n, L) in {"b4n12l4": (4, 12, 4), "b2n30l5": (2, 30, 5)}.items():
        _, y = slate_batch(g, B=B, n=n)
        p = torch.rand(B, n, L, generator=g) * 0.98 + 0.01
        run_loss(f"ordinal/{tag}", "ordinalLoss", ordinalLoss, {"y_pred": p, "y_true": y},
                 ("y_pred", "y_true"), {"n": L})

    # ---------------- risk functions ---------------------------------------------------
    known = torch.tensor([[0.0500, 0.1500, 0.3000, 0.4500, 0.5500, 0.4000, 0.3500, 0.3000],
                          [0.2500, 0.2000, 0.3000, 0.3000, 0.3000, 0.3000, 0.3000, 0.2500],
                          [0.2500, 0.2500, 0.2500, 0.2500, 0.4000, 0.1500, 0.4000, 0.1500],
                          [0.4000, 0.2000, 0.4500, 0.2000, 0.4500, 0.2000, 0.2542, 0.2629],
                          [0.2802, 0.2975, 0.3061, 0.2918, 0.2994, 0.3147, 0.3301, 0.3378]])
    for col in range(8):
        record(f"georisk_fn/known5x8/i{col}", "geoRisk", {"mat": known.numpy()}, {"alpha": 3, "i": col},
               geoRisk(known, 3, i=col).numpy(), {})
        record(f"zrisk_fn/known5x8/i{col}", "zRisk", {"mat": known.numpy()}, {"alpha": 3, "i": col},
               zRisk(known, 3, i=col).numpy(), {})
    record("georisk_default/known5x8", "getGeoRiskDefault", {"mat": known.numpy()}, {"alpha": 3},
           getGeoRiskDefault(known.numpy().astype(np.float64), 3), {})
    m = torch.rand(40, 4, generator=g)
    mm = m.clone().requires_grad_(True)
    out = geoRisk(mm, 5, requires_grad=True)
    out.sum().backward()
    record("georisk_fn/rand40x4/i0", "geoRisk", {"mat": m.numpy()}, {"alpha": 5, "i": 0}, out.detach().numpy(),
           {"mat": mm.grad.numpy()})

    # ---------------- risk losses ------------------------------------------------------
    B, n, S = 6, 16, 3
    s, y = slate_batch(g, B=B, n=n)
    base = torch.randn(B, n, S, generator=g)
    for fname in ("geoRiskListnetLoss", "zRiskListnetLoss"):
        for lt in (1, 2, 3):
            for rs in (1, 2, 3):
                for ideal in (1, 2):
                    run_loss(f"risk/{fname}/lt{lt}_rs{rs}_id{ideal}", fname, getattr(RL, fname),
                             {"y_pred": s, "y_true": y, "y_baselines": base},
                             ("y_pred", "y_true", "y_baselines"),
                             {"alpha": 5, "listnet_transformation": lt, "return_strategy": rs,
                              "add_ideal_ranking_to_mat": ideal})
    run_loss("risk/geoRiskListnetLoss/nobase_lt1_rs1_id2", "geoRiskListnetLoss", RL.geoRiskListnetLoss,
             {"y_pred": s, "y_true": y}, ("y_pred", "y_true"),
             {"alpha": 5, "listnet_transformation": 1, "return_strategy": 1, "add_ideal_ranking_to_mat": 2})
    run_loss("risk/geoRiskListnetLoss/neg", "geoRiskListnetLoss", RL.geoRiskListnetLoss,
             {"y_pred": s, "y_true": y, "y_baselines": base}, ("y_pred", "y_true", "y_baselines"),
             {"alpha": 2, "negative": -1})
    for fname in ("geoRiskLambdaLoss", "zRiskLambdaLoss"):
        for lt in (1, 2):
            for rs in (1, 2, 3):
                for ideal in (1, 2):
                    run_loss(f"risk/{fname}/lt{lt}_rs{rs}_id{ideal}", fname, getattr(RL, fname),
                             {"y_pred": s, "y_true": y, "y_baselines": base},
                             ("y_pred", "y_true", "y_baselines"),
                             {"alpha": 5, "listnet_transformation": lt, "return_strategy": rs,
                              "add_ideal_ranking_to_mat": ideal})
        run_loss(f"risk/{fname}/lambdaRank_lt1", fname, getattr(RL, fname),
                 {"y_pred": s, "y_true": y, "y_baselines": base}, ("y_pred", "y_true", "y_baselines"),
                 {"alpha": 5, "weighing_scheme": "lamdbaRank_scheme"})
    base1 = base.mean(dim=2)     # caller averages systems for t-risk (main_batch_execution.py:160)
    for fname in ("tRiskListnetLoss", "tRiskLambdaLoss"):
        for lt in (1, 2, 3):
            run_loss(f"risk/{fname}/lt{lt}", fname, getattr(RL, fname),
                     {"y_pred": s, "y_true": y, "y_baselines": base1}, ("y_pred", "y_true", "y_baselines"),
                     {"alpha": 5, "listnet_transformation": lt})

    # ---------------- evaluation metrics -----------------------------------------------
    for tag, kw in {"q8n20": dict(B=8, n=20), "q6n40_ties": dict(B=6, n=40, ties=True),
                    "q4n16_zero": dict(B=4, n=16, all_zero_row=True)}.items():
        s, y = slate_batch(g, **kw)
        for k in (1, 5, 10, None, 100):
            lst = torchNdcg(y, s, k=k, return_type="list")
            record(f"torchndcg/{tag}/k{k}/list", "torchNdcg", {"ys_true": y.numpy(), "ys_pred": s.numpy()},
                   {"k": k, "return_type": "list"}, np.array([float(v) for v in lst], dtype=np.float32), {})
            ten = torchNdcg(y, s, k=k, return_type="tensor")
            record(f"torchndcg/{tag}/k{k}/tensor", "torchNdcg", {"ys_true": y.numpy(), "ys_pred": s.numpy()},
                   {"k": k, "return_type": "tensor"}, ten.numpy(), {})
        for k in (1, 5, 10, 100):
            for gains in ("linear", "exponential"):
                for nr in (True, False):
                    v = mNdcg(y.numpy(), s.numpy(), k=k, no_relevant=nr, gains=gains)
                    record(f"mndcg/{tag}/k{k}_{gains}_{int(nr)}", "mNdcg",
                           {"true": y.numpy(), "pred": s.numpy()},
                           {"k": k, "no_relevant": nr, "gains": gains}, np.array(v, dtype=np.float64), {})
    s, y = slate_batch(g, B=4, n=20)
    lst = torchNdcg(y, s[:, :, None], k=5, return_type="tensor")     # [Q,n,1] predictions (main.py:145)
    record("torchndcg/q4n20_3d/k5/tensor", "torchNdcg", {"ys_true": y.numpy(), "ys_pred": s[:, :, None].numpy()},
           {"k": 5, "return_type": "tensor"}, lst.numpy(), {})

    # ---------------- torch.sort tie order (std::sort introsort) ------------
    for n in (8, 16, 17, 33, 64, 100, 128, 257, 1024):
        keys = (torch.randint(0, 4, (6, n), generator=g).float())
        keys[1, n // 2:] = float("-inf")
        keys[2] = torch.randn(n, generator=g)
        vals, idx = keys.sort(descending=True, dim=-1)
        record(f"sort/n{n}", "torch.sort", {"keys": keys.numpy()}, {"descending": True},
               idx.numpy().astype(np.int32), {})

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **arrays)
    with open(os.path.join(HERE, "golden_v1.json"), "w") as f:
        json.dump({"seed": SEED, "torch": torch.__version__, "numpy": np.__version__, "cases": manifest}, f, indent=0)
    print(f"wrote {len(manifest)} cases, {sum(a.nbytes for a in arrays.values()) / 1e6:.2f} MB raw")


if __name__ == "__main__":
    sys.exit(main())
