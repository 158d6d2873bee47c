#!/usr/bin/env python
"""Headline benchmark: fused approxNDCG forward+backward, slates/s (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (loss + d loss/d y_pred) over one batch of 4096 synthetic
MSLR-shaped slates of 128 docs (labels U{0..4}, valid length U{64..128}, -1 padding, scores
N(0,1), alpha = 10), per GPU.  Slates are independent, so N GPUs each take their own batch
(weak scaling, no data-path collective).  Rank 0 prints ONE JSON line.

  value        slates/s, inputs resident in HBM, CUDA events around K back-to-back launches of the
               C-ABI entry `ltr_approx_ndcg_fwdbwd`, max over ranks.
  e2e          same metric through the host-buffer C-ABI call (`ltr_approx_ndcg_fwdbwd_host`):
               pinned host y_pred/y_true -> (PCIe, read by the kernel itself) -> kernel -> (PCIe) loss and
               gradient in pinned host memory, every step, all inside the timed region.
  roofline     SFU (MUFU) roofline of the pairwise kernel: algorithmic 2 SFU ops per ordered valid
               pair / measured kernel time, against the SFU peak measured on this GPU by
               a pure ex2.approx loop in the same run; `roofline_hbm` gives the HBM view.
  cpu_baseline the reference's CPU algorithm on all host threads, on a bounded sample of the same workload: the UNMODIFIED
               reference files when they travelled with the tree (baseline/_ref/, git-ignored, copied byte for byte by
               oracle/install_reference.py in the build container: kind "reference"), else the oracle port
               (oracle/ltr_oracle.py, pinned to the reference by tests/golden: kind "port").
  ndcg_eval    the second half of BASELINE.json's metric, NDCG@{1,5,10} evaluation queries/s: device-resident value,
               roofline, e2e through `ltr_ndcg_at_k_host` (pinned host labels / scores in, pinned NDCG out) and the
               CPU port of torchNdcg on all host cores.  The driver's parser keeps `config`, `roofline`, `e2e` and
               `cpu_baseline`: the block is `config.ndcg_eval`, with its roofline / e2e / cpu numbers repeated as
               `roofline.ndcg_eval`, `e2e.ndcg_eval`, `cpu_baseline.ndcg_eval`.
  ranker_ddp_step (N > 1) `config.ranker_ddp_step`: the fused ranker training step with its gradients averaged over the ranks
                  through NVLink peer memory (ltr_peer_allreduce_avg) and through NCCL, us per step and whole-job slates/s.
  georisk_sharded (N > 1) `config.georisk_sharded`: BASELINE configs[3], 8192 queries over the N ranks, S in {2, 4}:
               us per fwd+bwd step with the in-kernel NVLink exchange and with one all-gather per round, the latency
               of one such all-gather, and a parity flag against the unsharded call on rank 0.

`--impl reference` times that CPU arm alone.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")
for _p in (ROOT, PKG_ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "approxNDCG fwd+bwd slates/sec (n=128)"
UNIT = "slates/s"
B_PER_GPU = 4096
N_DOCS = 128
ALPHA = 10.0
SEED = 2020
RING = 48          # distinct batches cycled through: 48 x (2+2+2 MB) = 288 MB > 126 MB L2


def make_batch(seed: int, B: int = B_PER_GPU, n: int = N_DOCS):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, 5, (B, n), generator=g).float()
    lens = torch.randint(n // 2, n + 1, (B,), generator=g)
    y[torch.arange(n)[None, :] >= lens[:, None]] = -1.0
    s = torch.randn(B, n, generator=g)
    return s, y, lens


# ----------------------------------------------------------------------------------------------
# clocks: NVML polling thread (the timed region is milliseconds long; nvidia-smi -lms is too slow)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
        0x10: "sync_boost",
    }

    def __init__(self, index: int):
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._h = None
        try:
            import pynvml
            self._nv = pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-safe: match by PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self._h = h
                        break
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._h is not None:
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._h is not None and self._t.is_alive():
            self._t.join()
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        reasons = [name for bit, name in self.REASONS.items() if self.reason_bits & bit]
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own code when its files travelled with the tree (baseline/_ref/, put there unmodified by
# oracle/install_reference.py in the build container: kind "reference"), else the oracle port (kind "port")
# ----------------------------------------------------------------------------------------------
_REF_MODULES = {}


def reference_module(rel):
    """baseline/_ref/<rel> loaded by path under a private name (the reference's package names `losses` / `utils` are the
    drop-in packages' names as well), or None when the file is not there."""
    if rel not in _REF_MODULES:
        import importlib.util
        path = os.path.join(ROOT, "baseline", "_ref", rel)
        mod = None
        if os.path.exists(path):
            try:
                spec = importlib.util.spec_from_file_location("_ltr_ref_" + rel.replace("/", "_")[:-3], path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
            except Exception:
                mod = None
        _REF_MODULES[rel] = mod
    return _REF_MODULES[rel]


def cpu_kind():
    return "reference" if reference_module("losses/approxNDCG.py") is not None else "port"


def cpu_what():
    return ("the UNMODIFIED reference losses/approxNDCG.py from baseline/_ref" if cpu_kind() == "reference"
            else "oracle/ltr_oracle.py port of losses/approxNDCG.py")


def cpu_step(s, y):
    sc = s.clone().requires_grad_(True)
    ref = reference_module("losses/approxNDCG.py")
    if ref is not None:
        loss = ref.approxNDCGLoss(sc, y, alpha=ALPHA)
    else:
        from oracle import ltr_oracle as O
        loss = O.approx_ndcg_loss(sc, y, alpha=ALPHA)
    loss.backward()
    return float(loss)


def cpu_throughput(sample_slates: int, reps: int, seed: int = SEED):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s, y, _ = make_batch(seed, B=sample_slates)
    cpu_step(s[:8], y[:8])
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(s, y)
        best = min(best, time.perf_counter() - t0)
    return sample_slates / best, cores


def _ndcg_cpu_worker(job):
    """One host process: the oracle port of torchNdcg (utils/metrics.py:83-104) at k = 10 over its chunk of queries."""
    seed, q, n = job
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, 5, (q, n), generator=g).float()
    s = torch.randn(q, n, generator=g)
    ref = reference_module("utils/metrics.py")
    t0 = time.perf_counter()
    if ref is not None:
        out = ref.torchNdcg(y, s, k=10, return_type="tensor")
    else:
        from oracle import ltr_oracle as O
        out = O.torch_ndcg(y, s, k=10, return_type="tensor")
    return time.perf_counter() - t0, float(out.sum())


def ndcg_cpu_throughput(queries_per_core: int = 768, n: int = N_DOCS):
    """torchNdcg@10 on every host core at once (the reference's evaluator is a Python loop per query: one process per
    core, fork -- call this BEFORE the CUDA context exists).  Returns (queries/s, cores, sample description)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    jobs = [(SEED + 7 * i, queries_per_core, n) for i in range(cores)]
    have_ref = reference_module("utils/metrics.py") is not None             # (loaded once in the parent, forked along)
    if not have_ref:
        from oracle import ltr_oracle  # noqa: F401
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_ndcg_cpu_worker, [(SEED, 8, n)] * cores)                # workers up
        t0 = time.perf_counter()
        pool.map(_ndcg_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    total = queries_per_core * cores
    what = ("the UNMODIFIED reference utils/metrics.py:83-104 torchNdcg from baseline/_ref" if have_ref
            else "oracle/ltr_oracle.py port of utils/metrics.py:83-104 torchNdcg")
    return total / wall, cores, (f"{total} queries x {n} docs, NDCG@10 only (the GPU sweep computes @1, @5 and @10), "
                                 f"{cores} processes x 1 thread ({what})")


WORKLOAD = ("approxNDCG alpha=10 fused fwd+bwd, 4096 slates x 128 docs per GPU per step (BASELINE configs[1]), "
            "labels U{0..4}, valid len U{64..128} (-1 padding), scores N(0,1)")


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # size the per-step sample so that (steps + warmup) steps end within ~2 minutes
    probe_rate, _ = cpu_throughput(256, 1)
    budget_s = 120.0
    sample = int(max(16, min(B_PER_GPU, budget_s * probe_rate / max(1, args.steps + args.warmup))))
    s, y, _ = make_batch(SEED, B=sample)
    for _ in range(args.warmup):
        cpu_step(s, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(s, y)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "host CPU",
                   "sample": f"{sample} slates x {N_DOCS} docs per step (bounded sample of the 4096 x 128 batch)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                         "sample": f"{sample} slates x {N_DOCS} docs per step, {args.steps} steps, torch CPU fp32, "
                                   f"{cores} threads ({cpu_what()})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def _bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run (and first-touch the pinned e2e buffers) on the CPUs NVML lists as local to this GPU, so
    that the kernel's PCIe reads of host memory do not cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda._get_nvml_device_index(local_rank)
                                              if hasattr(torch.cuda, "_get_nvml_device_index") else local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:                                   # no NVML / restricted container: keep the default placement
        print(f"bench.py: NUMA binding skipped ({type(e).__name__}: {e})", file=sys.stderr)


def run_b200_arm(args, rank, local_rank, world):
    import ctypes

    import torch.distributed as dist
    from ltr_b200 import _lib, _runtime as rt
    from ltr_b200.host import ApproxNDCGHost, NdcgEvalHost

    # ---- CPU baselines first (rank 0 of a single-GPU run): the NDCG one forks worker processes, which must happen
    #      before this process owns a CUDA context
    cpu = {}
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, sample = ndcg_cpu_throughput()
        cpu["ndcg"] = {"value": rate, "unit": "queries/s", "cores": cores, "sample": sample,
                       "kind": "reference" if reference_module("utils/metrics.py") is not None else "port"}
        rate, cores = cpu_throughput(1024, 3)
        cpu["approx"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                         "sample": f"1024 slates x 128 docs fwd+bwd, best of 3, torch CPU fp32 ({cpu_what()})"}

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        _bind_to_gpu_numa_node(local_rank)
    lib = _lib.load()
    B, n = B_PER_GPU, N_DOCS

    # ---- synthetic inputs: RING distinct batches per rank, generated on the host, resident in HBM ----
    host_batches = [make_batch(SEED + 1000 * rank + i) for i in range(RING)]
    pairs_per_batch = [int((l * (l - 1)).sum()) for _, _, l in host_batches]
    d_s = [s.to(dev) for s, _, _ in host_batches]
    d_y = [y.to(dev) for _, y, _ in host_batches]
    d_grad = [torch.empty(B, n, device=dev) for _ in range(RING)]
    d_slate = torch.empty(B, device=dev)
    d_loss = torch.empty(1, device=dev)
    tab = rt.disc_prefix(dev)
    ws = rt.workspace(dev)
    stream = rt.stream_ptr()

    def launch(i):
        k = i % RING
        rc = lib.ltr_approx_ndcg_fwdbwd(d_s[k].data_ptr(), d_y[k].data_ptr(), B, n, ALPHA, 1e-10, -1.0,
                                        tab.data_ptr(), tab.numel(), d_slate.data_ptr(), d_loss.data_ptr(),
                                        d_grad[k].data_ptr(), None, 0, ws.data_ptr(), stream)
        if rc:
            _lib.check(rc, "ltr_approx_ndcg_fwdbwd")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    # ---- device-resident throughput (`value`) --------------------------------------------------------
    # The kernel (~20 us) is shorter than a Python->ctypes->cudaLaunchKernel round trip, so the timed launches are
    # replayed from a CUDA graph of min(steps, RING) launches of the same C-ABI entry over distinct batches, captured on
    # the launching stream (the driver's 20-step run is ONE replay of a 20-launch graph); a remainder is launched directly.
    sampler = ClockSampler(local_rank)
    for i in range(args.warmup):
        launch(i)
    barrier()
    graph, n_replays, g_len = None, 0, min(args.steps, RING)
    if g_len >= 2 and not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            stream = rt.stream_ptr()
            ws = rt.workspace(dev)
            launch(0)                                   # warm this stream's workspace outside the capture
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for i in range(g_len):
                    launch(args.warmup + i)
        stream = rt.stream_ptr()
        ws = rt.workspace(dev)
        n_replays = args.steps // g_len
        for _ in range(3):
            graph.replay()                              # untimed replays (graph upload, clocks up)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n_replays):
        graph.replay()
    for i in range(n_replays * g_len, args.steps):
        launch(args.warmup + i)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    loss_value = float(d_loss.cpu())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the host C-ABI, copies inside the timed region ---------------------
    plan = ApproxNDCGHost(B, n, alpha=ALPHA, device=dev, depth=3)
    h_s = [s.pin_memory() for s, _, _ in host_batches[:8]]
    h_y = [y.pin_memory() for _, y, _ in host_batches[:8]]
    e2e_steps = max(1, min(args.steps, 2000))
    for i in range(min(args.warmup, 20) + 3):
        plan.submit(h_s[i % 8], h_y[i % 8])
    plan.wait_all()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        plan.submit(h_s[i % 8], h_y[i % 8])
    plan.wait_all()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    h2d = 2 * B * n * 4
    d2h = B * n * 4 + 4
    del plan, h_s, h_y

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json (burst: the kernel is timed alone)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    # ---- NDCG@{1,5,10} evaluation (the second half of BASELINE.json's metric), every rank its own queries ------------
    ndcg = bench_ndcg_eval(lib, rt, dev, rank, world, barrier, max_over_ranks, NdcgEvalHost, hbm_peak, hbm_src,
                           steps=max(10, min(args.steps, 100)))

    # ---- GeoRisk, query-sharded (BASELINE configs[3]) -----------------------------------------------------------------
    georisk = bench_georisk_sharded(dev, rank, world, barrier, max_over_ranks) if world > 1 else None

    # ---- the fused ranker step under data parallelism: gradients averaged over the ranks (peer memory / NCCL) ----------
    ddp_step = bench_ranker_ddp_step(dev, rank, world, barrier, max_over_ranks) if world > 1 else None

    if rank != 0:
        return

    # ---- roofline denominators measured here: SFU (ex2 loop), FP32 (FFMA loop) ------------------------
    probe_out = torch.zeros(1, device=dev)
    ops_count = ctypes.c_longlong(0)

    def probe(fn, iters):
        fn(probe_out.data_ptr(), iters, ctypes.byref(ops_count), stream)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(probe_out.data_ptr(), iters, ctypes.byref(ops_count), stream)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return ops_count.value / (best * 1e-3)

    sfu_peak = probe(lib.ltr_probe_sfu, 4096)
    fma_peak = probe(lib.ltr_probe_fma, 4096)

    # kernel time per launch = ms / steps (one launch per step, nothing else on the stream)
    t_launch = ms * 1e-3 / args.steps
    used = [(args.warmup + i) % RING for i in range(args.steps)]
    mean_pairs = sum(pairs_per_batch[k] for k in used) / len(used)
    sfu_ops = 2.0 * mean_pairs                       # 2 SFU ops per ordered valid pair
    achieved = sfu_ops / t_launch
    alg_bytes = 12.0 * n * B                          # read s, y; write grad
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("approx_ndcg_kernel_dram_bytes_per_launch")
    except Exception:
        pass

    config = {
        "workload": WORKLOAD,
        "global_batch": world * B, "slate_len": n, "parallelism": f"query-sharded x{world}, no data-path collective",
        "l2_policy": f"inputs larger than L2: ring of {RING} distinct batches ({RING * 3 * B * n * 4 // 2**20} MB of s/y/grad) cycled",
        "loss_check": loss_value,
        "launch_mode": (f"CUDA graph of {g_len} launches replayed {n_replays}x + {args.steps - n_replays * g_len} direct"
                        if graph is not None else "direct launches"),
        "ndcg_eval": ndcg,
    }
    if georisk is not None:
        config["georisk_sharded"] = georisk
    if ddp_step is not None:
        config["ranker_ddp_step"] = ddp_step
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "path": "ltr_approx_ndcg_fwdbwd_host: pinned host y_pred/y_true read by the kernel over PCIe, gradient and loss "
                        "written back to pinned host memory by the kernel (no staging copies), 3 streams in flight",
                "ndcg_eval": ndcg["e2e"]},
        "gpu_launches": args.steps,
        "clocks": clocks,
        "roofline": {
            "bound": "sfu", "bound_note": "pairwise kernel: MUFU (SFU) pipe roofline as BASELINE.json north_star asks; "
                                           "the HBM view of the same kernel is in roofline_hbm",
            "achieved": achieved / 1e12, "peak": sfu_peak / 1e12, "unit": "TSFU-op/s",
            "frac": achieved / sfu_peak, "traffic": traffic,
            "kernel": "approx_ndcg_sym_kernel", "kernel_us": t_launch * 1e6,
            "algorithmic": f"2 SFU ops per ordered valid pair, {mean_pairs:.0f} pairs per launch (SURVEY 8d); the kernel "
                           "itself issues 1 MUFU.RCP per ordered pair (sigma_ji = 1 - sigma_ij), so frac may exceed 0.5 "
                           "of the pipe's busy time and is capped at 2.0",
            "peak_source": "measured in this run: pure ex2.approx loop (ltr_probe_sfu)",
            "peak_nominal": 148 * 16 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12,
            "fp32_peak_measured_tflane_ops": fma_peak / 1e12,
            "ndcg_eval": ndcg["roofline"],
        },
        "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / t_launch / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg_bytes / t_launch / 1e9 / hbm_peak, "peak_source": hbm_src,
                         "note": "pairwise kernel is SFU/issue bound; HBM view shown for completeness"},
    }
    if world == 1 and not args.no_secondary:
        # the other kernel families of the path (ListNet, NDCG@k sweeps, LambdaLoss, GeoRisk, approxNDCG n=1024):
        # measured by tools/bench_paths.py with CUDA events; reported beside the headline, never mixed into it.
        # Their rooflines (achieved / peak / frac each) are repeated under roofline.others, which the driver keeps.
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_paths
            line["secondary"] = bench_paths.run_all(sfu_peak=sfu_peak)
            line["roofline"]["others"] = {k: dict(v["roofline"], us_per_call=v["us_per_call"], value=v["value"], unit_value=v["unit"])
                                          for k, v in line["secondary"].items() if isinstance(v, dict) and "roofline" in v}
        except Exception as exc:                      # the headline line must still be printed
            line["secondary"] = {"error": repr(exc)}
    if "approx" in cpu:
        line["cpu_baseline"] = dict(cpu["approx"], ndcg_eval=cpu["ndcg"])
        ndcg["cpu_baseline"] = cpu["ndcg"]
    print(json.dumps(line), flush=True)


def bench_ndcg_eval(lib, rt, dev, rank, world, barrier, max_over_ranks, NdcgEvalHost, hbm_peak, hbm_src, steps):
    """NDCG@{1,5,10} sweep, torchNdcg arithmetic (utils/metrics.py:83-104), 65,536 queries x 128 docs per GPU per step.
    value: queries resident in HBM, CUDA events over a CUDA graph of `steps` calls of `ltr_ndcg_at_k` cycling 3 distinct batches
    (201 MB > L2); e2e: pinned host labels / scores -> `ltr_ndcg_at_k_host` -> pinned NDCG, 3 streams in flight."""
    import ctypes
    from ltr_b200 import ops
    Q, n, ks, ring = 65536, N_DOCS, (1, 5, 10), 3
    host = []
    for i in range(ring):
        g = torch.Generator().manual_seed(SEED + 500 + 10 * rank + i)
        host.append((torch.randint(0, 5, (Q, n), generator=g).float(), torch.randn(Q, n, generator=g)))
    d = [(y.to(dev), s.to(dev)) for y, s in host]
    d32, _ = ops._disc_tables(n, dev)
    out = torch.empty(Q, len(ks), device=dev)
    marks = [torch.empty(Q + 1, dtype=torch.int32, device=dev) for _ in range(ring)]
    karr = (ctypes.c_int * len(ks))(*ks)
    stream = rt.stream_ptr()

    def step(i):
        y, s = d[i % ring]
        rc = lib.ltr_ndcg_at_k(y.data_ptr(), s.data_ptr(), Q, n, karr, len(ks), ops.TIES_STDSORT, d32.data_ptr(),
                               out.data_ptr(), marks[i % ring].data_ptr(), stream)
        if rc:
            from ltr_b200 import _lib
            _lib.check(rc, "ltr_ndcg_at_k")

    for i in range(5):
        step(i)
    barrier()
    # like the headline: the ~20 us kernel is shorter than a Python -> ctypes -> launch round trip (and with several ranks on
    # one host the launching threads compete for its cores), so the timed calls are ONE replay of a CUDA graph of `steps` calls
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        stream = rt.stream_ptr()
        step(0)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for i in range(steps):
                step(5 + i)
    stream = rt.stream_ptr()
    for _ in range(2):
        graph.replay()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    sec = ms * 1e-3 / steps
    check = float(torch.nan_to_num(out).sum().cpu())
    # e2e
    plan = NdcgEvalHost(Q, n, ks=ks, device=dev, depth=3)
    pinned = [(y.pin_memory(), s.pin_memory()) for y, s in host]
    e_steps = max(6, min(steps, 30))
    for i in range(3):
        plan.submit(*pinned[i % ring])
    plan.wait_all()
    barrier()
    t0 = time.perf_counter()
    for i in range(e_steps):
        plan.submit(*pinned[i % ring])
    plan.wait_all()
    e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    del plan, pinned
    alg = (8.0 * n + 4.0 * len(ks)) * Q
    gbs = alg / sec / 1e9
    return {
        "metric": "NDCG@{1,5,10} eval queries/s (n=128)", "value": world * Q / sec, "unit": "queries/s", "n_gpus": world,
        "us_per_call": sec * 1e6, "steps": steps,
        "workload": f"{Q} queries x {n} docs per GPU per step, k sweep {list(ks)}, torchNdcg arithmetic, labels U{{0..4}}, "
                    f"scores N(0,1), ring of {ring} batches ({ring * 2 * Q * n * 4 // 2**20} MB > L2); query-sharded x{world}",
        "check_sum_of_ndcg": check,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                     "peak_source": hbm_src, "kernel": "ndcg_key_kernel<false>", "kernel_us": sec * 1e6,
                     "algorithmic": "8 n + 4 |ks| bytes per query"},
        "e2e": {"value": world * Q * e_steps / (e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": 2 * Q * n * 4,
                "d2h_bytes_per_step": Q * len(ks) * 4, "steps": e_steps,
                "path": "ltr_ndcg_at_k_host: pinned host labels/scores -> cudaMemcpyAsync -> sweep kernel -> cudaMemcpyAsync "
                        "-> pinned [Q, 3] NDCG, 3 streams in flight"},
    }


def bench_ranker_ddp_step(dev, rank, world, barrier, max_over_ranks, B=B_PER_GPU, n=N_DOCS, F=136, iters=30):
    """BASELINE configs[1] as a data-parallel TRAINING step (the reference's DDP loop, main_batch_execution.py:120-171): every
    rank runs the fused TripleLayerNet + approxNDCG step on its own `B` slates, then the 13,441 gradient floats are averaged over
    the ranks -- through NVLink peer memory (`ltr_peer_allreduce_avg`, one launch) or by one NCCL all-reduce.  us per step
    (CUDA events, max over ranks), whole-job slates/s, and the check that both averages agree."""
    import torch.distributed as dist
    from ltr_b200 import _lib, _runtime as rt, sharding
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    lib = _lib.load()
    torch.manual_seed(SEED)                                               # the same weights on every rank
    net = TripleLayerNet(F).to(dev)
    ring = 3
    data = []
    for i in range(ring):
        g = torch.Generator().manual_seed(SEED + 1000 * rank + i)
        _, y, _ = make_batch(SEED + 7 * rank + i, B, n)
        data.append((torch.rand(B, n, F, generator=g).to(dev), y.to(dev)))
    prm = [t.detach() for t in (net.l1.weight, net.l1.bias, net.l2.weight, net.l2.bias, net.l3.weight, net.l3.bias)]
    sizes = [t.numel() for t in prm]
    flat = torch.zeros(sum(sizes), device=dev)
    offs = [sum(sizes[:i]) for i in range(len(sizes))]
    scores, slate, loss = torch.empty(B, n, device=dev), torch.empty(B, device=dev), torch.empty(1, device=dev)
    nscr = lib.ltr_triple_step_scratch_bytes(B, n, F)
    scr = torch.empty(nscr, dtype=torch.uint8, device=dev)
    tab, ws = rt.disc_prefix(dev), rt.workspace(dev)
    base = flat.data_ptr()
    ex = sharding.grad_peer_exchange(dist.group.WORLD, dev)

    def step(i, how):
        X, y = data[i % ring]
        rc = lib.ltr_triple_approx_ndcg_step(X.data_ptr(), y.data_ptr(), *[t.data_ptr() for t in prm], B, n, F, 10.0, 1e-10, -1.0,
                                             tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), loss.data_ptr(),
                                             *[base + 4 * o for o in offs], scr.data_ptr(), nscr, ws.data_ptr(), rt.stream_ptr())
        if rc != 0:
            raise RuntimeError("ltr_triple_approx_ndcg_step failed: " + lib.ltr_last_error().decode())
        if how == "p2p":
            ex.average([flat])
        elif how == "nccl":
            dist.all_reduce(flat)
            flat.mul_(1.0 / world)

    def timed(how):
        for i in range(5):
            step(i, how)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            step(5 + i, how)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / iters * 1e3            # us

    out = {"workload": f"TripleLayerNet({F}) + approxNDCG alpha=10 fused training step, {B} slates x {n} docs x {F} features per rank, "
                       f"gradients (13,441 floats) averaged over {world} ranks", "iters": iters,
           "peer_memory_available": ex is not None}
    out["step_only_us"] = timed("none")
    out["nccl_allreduce_us_per_step"] = timed("nccl")
    out["nccl_slates_per_s"] = world * B / (out["nccl_allreduce_us_per_step"] * 1e-6)
    if ex is not None:
        out["peer_memory_us_per_step"] = timed("p2p")
        out["peer_memory_slates_per_s"] = world * B / (out["peer_memory_us_per_step"] * 1e-6)
        step(0, "nccl")
        ref = flat.clone()
        step(0, "p2p")
        out["peer_vs_nccl_max_rel_diff"] = float((flat - ref).abs().max() / ref.abs().max())
        out["peer_status"] = ex.status()
        flag = torch.tensor([1 if (out["peer_vs_nccl_max_rel_diff"] < 1e-6 and out["peer_status"] == 0) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["peer_parity_ok"] = bool(int(flag))
    return out


def bench_georisk_sharded(dev, rank, world, barrier, max_over_ranks, Q_total=8192, n=N_DOCS, iters=200):
    """BASELINE configs[3]: geoRiskListnetLoss (lt=1, rs=1, alpha=5) fwd+bwd, `Q_total` queries sharded over the ranks,
    S baseline systems.  us per step (CUDA events, max over ranks) for the two exchange mechanisms, the latency of one
    all-gather of a phase state, and parity of the sharded loss / gradient with the unsharded call."""
    import torch.distributed as dist
    from ltr_b200 import ops, sharding
    group = dist.group.WORLD
    lo, hi = sharding.shard_bounds(Q_total, rank, world)
    out = {"workload": f"geoRiskListnetLoss lt=1 rs=1 alpha=5 fwd+bwd, {Q_total} queries x {n} docs over {world} ranks "
                       f"({hi - lo} per rank), through the drop-in callable's autograd Function", "iters": iters}

    def timed(fn, reps):
        for _ in range(10):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps * 1e3            # us

    for S in (2, 4):
        g = torch.Generator().manual_seed(SEED + 40 + S)                 # every rank draws the whole batch, keeps its rows
        s = torch.randn(Q_total, n, generator=g)
        y = torch.randint(0, 5, (Q_total, n), generator=g).float()
        base = torch.randn(Q_total, n, S, generator=g)
        sl = s[lo:hi].to(dev).requires_grad_(True)
        yl, bl = y[lo:hi].to(dev), base[lo:hi].to(dev)
        res = {}
        for mode in ("p2p", "allgather"):
            os.environ["LTR_RISK_P2P"] = "1" if mode == "p2p" else "0"
            used_p2p = sharding.risk_peer_exchange(group, dev) is not None
            res[mode + "_us"] = timed(lambda: ops.risk_listnet(sl, yl, bl, ops.KIND_GEO, 5.0, 1, 1, 1.0, False, False, group), iters)
            if mode == "p2p":
                res["p2p_available"] = used_p2p
            sl.grad = None
            loss = ops.risk_listnet(sl, yl, bl, ops.KIND_GEO, 5.0, 1, 1, 1.0, False, False, group)
            loss.sum().backward()
            res[mode + "_loss"] = float(loss)
            if rank == 0:                                                 # the unsharded call on the whole batch
                sf = s.to(dev).requires_grad_(True)
                lf = ops.risk_listnet(sf, y.to(dev), base.to(dev), ops.KIND_GEO, 5.0, 1, 1, 1.0, False, False, None)
                lf.sum().backward()
                gref = sf.grad[lo:hi]
                res[mode + "_loss_rel_err"] = abs(float(loss) - float(lf)) / abs(float(lf))
                res[mode + "_grad_rel_err"] = float((sl.grad - gref).abs().max() / gref.abs().max())
                ok = res[mode + "_loss_rel_err"] < 1e-5 and res[mode + "_grad_rel_err"] < 1e-4
            else:
                ok = True
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            res[mode + "_parity_ok"] = bool(int(flag))
        os.environ["LTR_RISK_P2P"] = "1"
        out[f"S{S}"] = res
    blk = torch.zeros(1024, dtype=torch.float64, device=dev)
    out["allgather_16_doubles_us"] = timed(lambda: sharding.allgather_risk_round(blk, 0, group), iters)
    one = torch.zeros(1, device=dev)
    out["allreduce_1_float_us"] = timed(lambda: dist.all_reduce(one), iters)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch every timed step from Python (no CUDA graph)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary kernel-family measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baselines (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29541"),
               os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup)]
        raise SystemExit(subprocess.call(cmd))

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / logs off stdout: rank 0 prints ONE JSON line
        # NCCL's version banner is a C-level printf that ignores NCCL_DEBUG_FILE: file descriptor 1 goes to stderr for
        # native code, Python's own stdout (the JSON line) keeps the original descriptor
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200_arm(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
