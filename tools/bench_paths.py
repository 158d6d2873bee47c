"""Secondary measurements of the hot path (every kernel family besides the headline approxNDCG line):
ListNet fwd+bwd, NDCG@k evaluation sweep, LambdaLoss, GeoRisk-ListNet, approxNDCG at n = 1024.

    python tools/bench_paths.py [--only listnet,ndcg,...] [--iters N] [--profile]

Each entry is timed with CUDA events on the launching stream after warm-up, over a ring of distinct
batches larger than the 126 MB L2 where the kernel is HBM-bound, and reports algorithmic bytes / time
against the measured HBM peak (MEASURED_PEAKS.json).  `--profile` runs 3 iterations per kernel only (for ncu).
bench.py imports `run_all` for the `secondary` block of its JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]

import torch  # noqa: E402


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


def _ok(rc):
    """A bench step that fails is not a fast step: every C-ABI call's status is checked."""
    if rc != 0:
        from ltr_b200 import _lib
        raise RuntimeError(f"C-ABI call failed with status {rc}: {_lib.load().ltr_last_error().decode()}")


def _time(fn, iters, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(warm + i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3          # seconds per call


def _slates(B, n, seed, pad=True, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    if pad:
        lens = torch.randint(n // 2, n + 1, (B,), generator=g)
        y[torch.arange(n)[None, :] >= lens[:, None]] = -1.0
    s = torch.randn(B, n, generator=g)
    return s.to(dev), y.to(dev)


def _roof(alg_bytes, sec):
    peak, src = _hbm_peak()
    gbs = alg_bytes / sec / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": src}


_SFU_PEAK = [None]          # SFU-op/s measured by the caller (bench.py: ltr_probe_sfu) or here on first use


def _sfu_peak():
    if _SFU_PEAK[0] is None:
        import ctypes
        from ltr_b200 import _lib, _runtime as rt
        lib = _lib.load()
        out = torch.zeros(1, device="cuda")
        ops_count = ctypes.c_longlong(0)
        best = float("inf")
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lib.ltr_probe_sfu(out.data_ptr(), 4096, ctypes.byref(ops_count), rt.stream_ptr())
            b.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, a.elapsed_time(b))
        _SFU_PEAK[0] = ops_count.value / (best * 1e-3)
    return _SFU_PEAK[0]


def _roof_sfu(sfu_ops, sec, **extra):
    peak = _sfu_peak()
    return dict({"bound": "sfu", "achieved": sfu_ops / sec / 1e12, "peak": peak / 1e12, "unit": "TSFU-op/s",
                 "frac": sfu_ops / sec / peak, "peak_source": "pure ex2.approx loop (ltr_probe_sfu), measured in this run"}, **extra)


def bench_listnet(iters=200, B=65536, n=128):
    """BASELINE configs[0] loss: ListNet fwd+bwd, no padding.  12 n bytes per slate."""
    from ltr_b200 import _lib, _runtime as rt
    lib = _lib.load()
    ring = 3                                             # 3 x 96 MB of y/s/grad > L2
    data = [_slates(B, n, 100 + i, pad=False) for i in range(ring)]
    grads = [torch.empty(B, n, device="cuda") for _ in range(ring)]
    loss = torch.empty(1, device="cuda")
    ws = rt.workspace(torch.device("cuda", torch.cuda.current_device()))

    def step(i):
        s, y = data[i % ring]
        _ok(lib.ltr_listnet_fwdbwd(y.data_ptr(), s.data_ptr(), B, n, 0, None, loss.data_ptr(), grads[i % ring].data_ptr(),
                               ws.data_ptr(), rt.stream_ptr()))

    sec = _time(step, iters)
    return {"metric": "ListNet fwd+bwd slates/s", "value": B / sec, "unit": "slates/s", "us_per_call": sec * 1e6,
            "config": {"workload": f"{B} slates x {n} docs, no padding, ring of {ring} batches"},
            "roofline": dict(_roof(12.0 * n * B, sec), kernel="listnet_vec_kernel<1>", algorithmic="12 n bytes per slate")}


def bench_ndcg(iters=100, Q=65536, n=128, ks=(1, 5, 10), ring=3):
    """NDCG@{1,5,10} evaluation sweep through the C-ABI (preallocated outputs): one read of the query, no sort unless
    tied scores reach the first k + 1 ranks.  8 n + 4 |ks| bytes per query."""
    import ctypes
    from ltr_b200 import _lib, ops, _runtime as rt
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    data = [_slates(Q, n, 200 + i, pad=False) for i in range(ring)]
    d32, _ = ops._disc_tables(n, dev)
    out = torch.empty(Q, len(ks), device=dev)
    marks = [torch.empty(Q + 1, dtype=torch.int32, device=dev) for _ in range(ring)]
    karr = (ctypes.c_int * len(ks))(*ks)

    need = int(lib.ltr_ndcg_at_k_scratch_bytes(Q, n)) if not os.environ.get("LTR_NDCG_NO_CHUNKS") else 0
    scratch = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)

    def step(i):
        s, y = data[i % ring]
        if need:
            _ok(lib.ltr_ndcg_at_k_ws(y.data_ptr(), s.data_ptr(), Q, n, karr, len(ks), ops.TIES_STDSORT, d32.data_ptr(), out.data_ptr(),
                                     marks[i % ring].data_ptr(), scratch.data_ptr(), need, rt.stream_ptr()))
        else:
            _ok(lib.ltr_ndcg_at_k(y.data_ptr(), s.data_ptr(), Q, n, karr, len(ks), ops.TIES_STDSORT, d32.data_ptr(), out.data_ptr(),
                                  marks[i % ring].data_ptr(), rt.stream_ptr()))

    sec = _time(step, iters)
    marked = [int((m[1:] != 0).sum()) for m in marks]
    kern = "ndcg_key_kernel<false> (sorts the queries it marks itself)" if n <= 128 and max(ks) <= 10 else \
        ("ndcg_key_kernel<false> per 128-doc chunk + ndcg_chunk_merge_kernel + ndcg_kernel<false> (marked queries only)" if need else
         "ndcg_tpq_wide_kernel<false, 4> + ndcg_kernel<false> (marked queries only)")
    return {"metric": f"NDCG@{{{','.join(map(str, ks))}}} eval queries/s (n={n})", "value": Q / sec, "unit": "queries/s",
            "us_per_call": sec * 1e6,
            "config": {"workload": f"{Q} queries x {n} docs, k sweep {list(ks)}, torchNdcg arithmetic, ring of {ring} batches",
                       "queries_with_tied_top_scores_per_batch": marked},
            "roofline": dict(_roof((8.0 * n + 4.0 * len(ks)) * Q, sec), kernel=kern,
                             algorithmic="8 n + 4 |ks| bytes per query")}


def bench_ndcg_large(iters=20):
    """The same sweep on an evaluation set of 2^20 queries (tile granularity and launch gaps amortised)."""
    return bench_ndcg(iters, Q=1 << 20, n=128, ring=2)


def bench_lambda(iters=50, B=1024, n=256):
    """BASELINE configs[2]: LambdaLoss ndcgLoss2PP, 256-doc slates, 1024 slates per GPU."""
    from ltr_b200 import _lib, _runtime as rt
    lib = _lib.load()
    g = torch.Generator().manual_seed(300)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    lens = torch.randint(n // 2, n + 1, (B,), generator=g)
    y[torch.arange(n)[None, :] >= lens[:, None]] = -1.0
    s = torch.randn(B, n, generator=g)
    yv = y.clone()
    pairs = 0
    for lv in range(5):                                  # ordered pairs with label_i > label_j, both valid
        cnt_hi = (yv > lv).sum(1)
        cnt_eq = (yv == lv).sum(1)
        pairs += int((cnt_hi * cnt_eq).sum())
    s, y = s.cuda(), y.cuda()
    sums = torch.empty(2, B, device="cuda")
    out = torch.empty(2, device="cuda")
    grad = torch.empty(B, n, device="cuda")
    ws = rt.workspace(torch.device("cuda", torch.cuda.current_device()))

    def step(i):
        _ok(lib.ltr_lambda_fwdbwd(s.data_ptr(), y.data_ptr(), B, n, 1e-10, -1.0, 4, -1, 1.0, 10.0, 0, 0, sums[0].data_ptr(),
                              sums[1].data_ptr(), out[0:1].data_ptr(), out[1:2].data_ptr(), grad.data_ptr(),
                              ws.data_ptr(), rt.stream_ptr()))

    sec = _time(step, iters)
    return {"metric": "LambdaLoss ndcgLoss2PP fwd+bwd slates/s (n=256)", "value": B / sec, "unit": "slates/s",
            "us_per_call": sec * 1e6,
            "config": {"workload": f"{B} slates x {n} docs, valid len U{{128..256}}, sigma=1, mu=10, sum, binary"},
            "roofline": _roof_sfu(3.0 * pairs, sec,
                                  algorithmic=f"3 SFU ops per masked ordered pair, {pairs} pairs per launch (SURVEY 8d)",
                                  kernel="lambda_kernel<4>")}


def bench_georisk(iters=200, B=8192, n=128, S=2):
    """BASELINE configs[3] on one GPU: geoRiskListnetLoss lt=1, rs=1, alpha=5, fwd+bwd = three launches (rows + phase-1
    state, z-risk sums, loss + backward rows).  Timed twice: through the drop-in callable's autograd Function (what a
    training step pays, Python included) and as the bare C-ABI call replayed from a CUDA graph (the kernels alone)."""
    from ltr_b200 import _lib, _runtime as rt, ops
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(400)
    s = torch.randn(B, n, generator=g).cuda().requires_grad_(True)
    y = torch.randint(0, 5, (B, n), generator=g).float().cuda()
    base = torch.randn(B, n, S, generator=g).cuda()

    def step(i):
        ops.risk_listnet(s, y, base, ops.KIND_GEO, 5.0, 1, 1, 1.0, False)

    sec = _time(step, iters)
    # bare C-ABI call, captured: 20 calls per graph
    mat = torch.empty(B, 1 + S, device=dev)
    rows = torch.empty(B, 3, dtype=torch.float64, device=dev)
    loss, grad = torch.empty(1, device=dev), torch.empty(B, n, device=dev)
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    sd = s.detach()
    with torch.cuda.stream(side):
        xchg, ws, st = rt.risk_xchg(dev), rt.workspace(dev), rt.stream_ptr()

        def raw(phases=7):
            _lib.check(lib.ltr_risk_listnet_fwdbwd(sd.data_ptr(), y.data_ptr(), base.data_ptr(), B, n, S, 1, 0, ops.KIND_GEO, 1, 1,
                                                   0, 5.0, 1.0, mat.data_ptr(), rows.data_ptr(), loss.data_ptr(), grad.data_ptr(),
                                                   phases, xchg.data_ptr(), 0, 1, None, ws.data_ptr(), st), "ltr_risk_listnet_fwdbwd")
        raw()
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(20):
                raw()
        graph3 = torch.cuda.CUDAGraph()                       # the three-launch form (what all-gather sharding uses)
        with torch.cuda.graph(graph3, stream=side):
            for _ in range(20):
                raw(1), raw(2), raw(4)
    sec_graph = _time(lambda i: graph.replay(), max(5, iters // 20)) / 20
    sec_graph3 = _time(lambda i: graph3.replay(), max(5, iters // 20)) / 20
    alg = (4.0 * n * (2 + S) + 4.0 * n) * B + 8.0 * n * B      # rows read + grad write + backward re-read of s, y
    return {"metric": "geoRiskListnetLoss fwd+bwd queries/s (n=128, S=2)", "value": B / sec_graph, "unit": "queries/s",
            "us_per_call": sec_graph * 1e6, "us_per_call_three_launches": sec_graph3 * 1e6,
            "us_per_call_python_callable": sec * 1e6,
            "config": {"workload": f"{B} queries x {n} docs x {S} baseline systems, lt=1, rs=1; value / us_per_call: the C-ABI "
                                   "call (ONE cooperative launch) replayed from a CUDA graph of 20 calls; "
                                   "us_per_call_three_launches: the same call phase by phase; us_per_call_python_callable: the "
                                   "drop-in callable's autograd Function, one call per step"},
            "roofline": dict(_roof(alg, sec_graph), kernel="risk_coop_kernel<4,2> (one cooperative launch, two grid rendezvous)",
                             algorithmic="4 n (2+S) read + 4 n written + 8 n re-read per query")}


def bench_approx_1024(iters=50, B=64, n=1024):
    """BASELINE configs[4] loss: approxNDCG on 1024-doc slates (one thread-block cluster per slate)."""
    from ltr_b200 import _lib, _runtime as rt
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(500)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    lens = torch.randint(n // 2, n + 1, (B,), generator=g)
    y[torch.arange(n)[None, :] >= lens[:, None]] = -1.0
    s = torch.randn(B, n, generator=g)
    pairs = int((lens * (lens - 1)).sum())
    s, y = s.cuda(), y.cuda()
    grad, slate, loss = torch.empty_like(s), torch.empty(B, device="cuda"), torch.empty(1, device="cuda")
    tab, ws = rt.disc_prefix(dev), rt.workspace(dev)
    nscratch = lib.ltr_approx_ndcg_scratch_bytes(B, n)
    scratch = torch.empty(max(nscratch, 1), dtype=torch.uint8, device="cuda")

    def step(i):
        _ok(lib.ltr_approx_ndcg_fwdbwd(s.data_ptr(), y.data_ptr(), B, n, 10.0, 1e-10, -1.0, tab.data_ptr(), tab.numel(),
                                   slate.data_ptr(), loss.data_ptr(), grad.data_ptr(), scratch.data_ptr(), nscratch,
                                   ws.data_ptr(), rt.stream_ptr()))

    sec = _time(step, iters)
    return {"metric": "approxNDCG fwd+bwd slates/s (n=1024)", "value": B / sec, "unit": "slates/s", "us_per_call": sec * 1e6,
            "config": {"workload": f"{B} slates x {n} docs, valid len U{{512..1024}}, alpha=10"},
            "roofline": _roof_sfu(2.0 * pairs, sec, algorithmic=f"2 SFU ops per ordered valid pair, {pairs} pairs per launch",
                                  kernel="approx_ndcg_cluster_kernel (one cluster of 8 CTAs per slate, partial sums over "
                                         "distributed shared memory)")}


def bench_approx_large(iters=30, B=65536, n=128):
    """The headline kernel on a 16x larger batch (launch ramp and tail amortised): same slate distribution."""
    from ltr_b200 import _lib, _runtime as rt
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    ring = 2
    data = []
    pairs = []
    for i in range(ring):
        g = torch.Generator().manual_seed(600 + i)
        y = torch.randint(0, 5, (B, n), generator=g).float()
        lens = torch.randint(n // 2, n + 1, (B,), generator=g)
        y[torch.arange(n)[None, :] >= lens[:, None]] = -1.0
        s = torch.randn(B, n, generator=g)
        pairs.append(int((lens * (lens - 1)).sum()))
        data.append((s.cuda(), y.cuda()))
    grad = [torch.empty(B, n, device="cuda") for _ in range(ring)]
    slate, loss = torch.empty(B, device="cuda"), torch.empty(1, device="cuda")
    tab, ws = rt.disc_prefix(dev), rt.workspace(dev)

    def step(i):
        s, y = data[i % ring]
        _ok(lib.ltr_approx_ndcg_fwdbwd(s.data_ptr(), y.data_ptr(), B, n, 10.0, 1e-10, -1.0, tab.data_ptr(), tab.numel(),
                                   slate.data_ptr(), loss.data_ptr(), grad[i % ring].data_ptr(), None, 0, ws.data_ptr(),
                                   rt.stream_ptr()))

    sec = _time(step, iters)
    return {"metric": "approxNDCG fwd+bwd slates/s (n=128, 65536 slates per launch)", "value": B / sec, "unit": "slates/s",
            "us_per_call": sec * 1e6, "config": {"workload": f"{B} slates x {n} docs, valid len U{{64..128}}, alpha=10"},
            "roofline": _roof_sfu(2.0 * (sum(pairs) / ring), sec,
                                  algorithmic="2 SFU ops per ordered valid pair (the kernel issues one MUFU per UNORDERED pair "
                                              "and pass: frac can exceed 1)", kernel="approx_ndcg_warp_kernel")}


def bench_ranker_step(iters=30, B=4096, n=128, F=136):
    """BASELINE configs[1] as a TRAINING STEP: TripleLayerNet forward + approxNDCG (alpha=10) + backward to all six
    parameter gradients.  fused: `ltr_triple_approx_ndcg_fwdbwd` (one TMA + tcgen05 kernel, X read once);
    torch+K1: stock torch.nn modules (cuBLAS, fp32 or tf32 matmuls) with the fused approxNDCG loss kernel; both through
    autograd, CUDA events over `iters` steps on 3 distinct feature batches (3 x 285 MB > L2)."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_approx_ndcg
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    torch.manual_seed(0)
    net = TripleLayerNet(F).cuda()
    ring = 3
    data = []
    for i in range(ring):
        g = torch.Generator().manual_seed(700 + i)
        X = torch.rand(B, n, F, generator=g).cuda()
        _, y = _slates(B, n, 710 + i)
        data.append((X, y))

    def fused(i):
        X, y = data[i % ring]
        net.zero_grad(set_to_none=True)
        fused_triple_approx_ndcg(net, X, y, alpha=10.0).backward()

    def stock(i):
        X, y = data[i % ring]
        net.zero_grad(set_to_none=True)
        approxNDCGLoss(net(X, None, None).squeeze(-1), y, alpha=10.0).backward()

    sec = _time(fused, iters)
    # the same step through the C-ABI alone (what a training loop without Python in it pays): fold, fused kernel, reduce, unfold
    from ltr_b200 import _lib, _runtime as rt
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    prm = [t.detach() for t in (net.l1.weight, net.l1.bias, net.l2.weight, net.l2.bias, net.l3.weight, net.l3.bias)]
    outs = [torch.empty_like(t) for t in prm]
    scores, slate, loss = torch.empty(B, n, device=dev), torch.empty(B, device=dev), torch.empty(1, device=dev)
    nscr = lib.ltr_triple_step_scratch_bytes(B, n, F)
    scr = torch.empty(nscr, dtype=torch.uint8, device=dev)
    tab, ws = rt.disc_prefix(dev), rt.workspace(dev)

    def cabi(i):
        X, y = data[i % ring]
        _ok(lib.ltr_triple_approx_ndcg_step(X.data_ptr(), y.data_ptr(), *[t.data_ptr() for t in prm], B, n, F, 10.0, 1e-10, -1.0,
                                            tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), loss.data_ptr(),
                                            *[t.data_ptr() for t in outs], scr.data_ptr(), nscr, ws.data_ptr(), rt.stream_ptr()))

    sec_abi = _time(cabi, iters)

    def cabi_listnet(i):                                    # the same step with the ListNet loss (losses/listnet.py:5-16)
        X, y = data[i % ring]
        _ok(lib.ltr_triple_listnet_step(X.data_ptr(), y.data_ptr(), *[t.data_ptr() for t in prm], B, n, F, tab.data_ptr(), tab.numel(),
                                        scores.data_ptr(), slate.data_ptr(), loss.data_ptr(), *[t.data_ptr() for t in outs],
                                        scr.data_ptr(), nscr, ws.data_ptr(), rt.stream_ptr()))

    sec_listnet = _time(cabi_listnet, iters)
    torch.backends.cuda.matmul.allow_tf32 = False
    sec_fp32 = _time(stock, max(5, iters // 3))
    torch.backends.cuda.matmul.allow_tf32 = True
    sec_tf32 = _time(stock, max(5, iters // 3))
    torch.backends.cuda.matmul.allow_tf32 = False
    alg = (4.0 * F * n + 8.0 * n) * B                       # features once, labels in, scores out
    return {"metric": "TripleLayerNet + approxNDCG training step slates/s (n=128, F=136)", "value": B / sec_abi, "unit": "slates/s",
            "us_per_call": sec_abi * 1e6, "us_per_step_through_autograd": sec * 1e6, "us_per_call_listnet_variant": sec_listnet * 1e6,
            "us_per_step_torch_modules_fp32_plus_K1": sec_fp32 * 1e6, "us_per_step_torch_modules_tf32_plus_K1": sec_tf32 * 1e6,
            "speedup_vs_torch_fp32": sec_fp32 / sec_abi, "speedup_vs_torch_tf32": sec_tf32 / sec_abi,
            "speedup_through_autograd_vs_torch_fp32": sec_fp32 / sec, "speedup_through_autograd_vs_torch_tf32": sec_tf32 / sec,
            "note": "us_per_call = ltr_triple_approx_ndcg_step through the C-ABI (4 launches: fold, fused kernel, reduce, unfold); through "
                    "torch.autograd the same step is bound by ~200 us of Python / autograd-engine time per step, the torch-module steps "
                    "are GPU-bound",
            "config": {"workload": f"{B} slates x {n} docs x {F} features, valid len U{{64..128}}, alpha=10; forward + backward "
                                   "to l1/l2/l3 weight and bias gradients, through autograd, zero_grad included"},
            "roofline": dict(_roof(alg, sec_abi), kernel="ranker_fold_kernel + ranker_fused_kernel<3> (TMA + tcgen05 tf32, 1 CTA per SM) + "
                                                         "ranker_fused_reduce_kernel + ranker_unfold_kernel",
                             algorithmic="4 F n + 8 n bytes per slate (features read once, labels, scores)",
                             second_bound="every feature byte crosses L2 -> SM twice (forward copy, gradient copy): 2 x 285 MB at the "
                                          "11.5 TB/s an L2-resident TMA stream reaches (tools/probe_tma.py) = 50 us")}


def bench_ndcg_1024(iters=30):
    """BASELINE configs[4] evaluation sweep: 8192 queries x 1024 docs."""
    return bench_ndcg(iters, Q=8192, n=1024)


ALL = {"ranker_step": bench_ranker_step, "listnet": bench_listnet, "ndcg": bench_ndcg, "ndcg_large": bench_ndcg_large, "ndcg1024": bench_ndcg_1024, "lambda": bench_lambda,
       "georisk": bench_georisk, "approx1024": bench_approx_1024, "approx_large_batch": bench_approx_large}


def run_all(only=None, iters=None, sfu_peak=None):
    if sfu_peak is not None:
        _SFU_PEAK[0] = sfu_peak
    out = {}
    for name, fn in ALL.items():
        if only and name not in only:
            continue
        out[name] = fn(iters) if iters else fn()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    res = run_all([x for x in a.only.split(",") if x] or None, 3 if a.profile else (a.iters or None))
    for k, v in res.items():
        print(json.dumps({k: v}))
