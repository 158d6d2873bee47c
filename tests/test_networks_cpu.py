"""The generated selection networks of the NDCG@k kernels (csrc/ndcg_select32.inc, csrc/ndcg_top11_32.inc) are checked
on the CPU: replaying their compare-exchange lists on random inputs (with ties and -inf) must give the 11th largest /
the 11 largest values in order, and the committed files must be what tools/gen_select32.py generates."""
import os
import random
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200", "csrc")


def _ces(name, macro):
    text = open(os.path.join(CSRC, name)).read()
    return [(int(a), int(b)) for a, b in re.findall(macro + r"\((\d+), (\d+)\);", text)]


def _run(v, ces):
    v = list(v)
    for a, b in ces:
        if v[a] < v[b]:
            v[a], v[b] = v[b], v[a]
    return v


def _inputs(rng):
    kind = rng.randrange(3)
    if kind == 0:
        return [rng.random() for _ in range(32)]
    if kind == 1:
        return [rng.choice([0.0, 0.25, 0.5, 1.0]) for _ in range(32)]           # heavy ties
    return [rng.random() if rng.random() < 0.4 else float("-inf") for _ in range(32)]


def test_select11_network():
    ces = _ces("ndcg_select32.inc", "TPQ_CE")
    assert ces and all(0 <= a < b < 32 for a, b in ces)
    rng = random.Random(1)
    for _ in range(3000):
        v = _inputs(rng)
        assert _run(v, ces)[10] == sorted(v, reverse=True)[10]


def test_top11_network():
    ces = _ces("ndcg_top11_32.inc", "TPQ_CEP")
    assert ces and all(0 <= a < b < 32 for a, b in ces)
    rng = random.Random(2)
    for _ in range(3000):
        v = _inputs(rng)
        assert _run(v, ces)[:11] == sorted(v, reverse=True)[:11]


def test_generated_files_are_current(tmp_path):
    committed = {n: open(os.path.join(CSRC, n)).read() for n in ("ndcg_select32.inc", "ndcg_top11_32.inc")}
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_select32.py"), str(tmp_path)], check=True,
                   capture_output=True)
    generated = {n: open(os.path.join(str(tmp_path), n)).read() for n in committed}
    assert committed == generated
