"""The committed bench lines of the current round (profiles/r02_bench_n*.json, written by `bench.py` on the GPU box) carry every
key of the measurement contract: metric / config / e2e / gpu_launches / clocks, `roofline` with its measured peak and per-launch
traffic, `cpu_baseline` at N = 1 with its kind and sample, the NDCG-eval half of the metric, and at N > 1 the sharded GeoRisk and
data-parallel ranker blocks with their in-run parity flags.  A format regression in bench.py shows up here before a GPU run."""
import json
import os

import pytest

from conftest import ROOT

PROFILES = os.path.join(ROOT, "profiles")


def _line(n):
    path = os.path.join(PROFILES, f"r02_bench_n{n}.json")
    if not os.path.exists(path):
        pytest.skip(f"{path} not committed")
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_bench_line_contract(n):
    d = _line(n)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in d, key
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                                  # BASELINE.md holds no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] >= d["steps"] > 0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0
    ev = d["config"]["ndcg_eval"]
    assert ev["value"] > 0 and 0 < ev["roofline"]["frac"] <= 1.0 and ev["e2e"]["value"] > 0
    # value = units of all ranks / max-over-ranks time
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]


def test_single_gpu_line_has_the_cpu_arm_and_every_kernel_family():
    d = _line(1)
    cb = d["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in cb, key
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0
    assert cb["ndcg_eval"]["value"] > 0
    fams = d["secondary"]
    for fam in ("ranker_step", "listnet", "ndcg", "ndcg_large", "ndcg1024", "lambda", "georisk", "approx1024"):
        assert fam in fams and 0 < fams[fam]["roofline"]["frac"], fam
    ref = os.path.join(PROFILES, "r02_bench_reference_arm.json")
    with open(ref) as f:
        a = json.loads(f.read().strip().splitlines()[-1])
    assert a["impl"] == "reference" and a["metric"] == d["metric"] and a["unit"] == d["unit"]
    assert a["e2e"]["h2d_bytes_per_step"] == 0 and a["e2e"]["d2h_bytes_per_step"] == 0 and a["e2e"]["value"] == a["value"]
    assert a["cpu_baseline"]["kind"] in ("reference", "port")


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines_carry_parity_flags(n):
    d = _line(n)
    g = d["config"]["georisk_sharded"]
    for S in ("S2", "S4"):
        assert g[S]["p2p_parity_ok"] is True and g[S]["allgather_parity_ok"] is True
        assert g[S]["p2p_us"] > 0 and g[S]["allgather_us"] > 0
    r = d["config"]["ranker_ddp_step"]
    assert r["peer_memory_available"] is True and r["peer_parity_ok"] is True and r["peer_status"] == 0
    assert r["peer_memory_us_per_step"] < r["nccl_allreduce_us_per_step"]
    assert r["step_only_us"] < r["peer_memory_us_per_step"]
