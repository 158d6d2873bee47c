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


# ---- the label-in-key scheme of csrc/ndcg_keys.cuh, restated in numpy -------------------------------------------
def _keys(scores, labels):
    """key = (bits(score + 0.0) & ~7) | label, read back as fp32 (key_of in ndcg_keys.cuh)."""
    import numpy as np
    s = (np.asarray(scores, dtype=np.float32) + np.float32(0.0)).view(np.uint32)
    t = (np.asarray(labels, dtype=np.float32) + np.float32(8388608.0)).view(np.uint32)      # label in the low mantissa bits
    return ((s & np.uint32(0xFFFFFFF8)) | (t & np.uint32(7))).view(np.float32)


def test_label_in_key_order_matches_score_order():
    """Keys whose upper 29 bits differ compare exactly like the scores (any signs, denormals, huge values); keys whose
    upper 29 bits agree are what the kernel redoes with exact compares; the label is read back from the low 3 bits."""
    import numpy as np
    rng = np.random.default_rng(5)
    n = 200000
    bits = rng.integers(0, 2**32, size=2 * n, dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32)
    vals = vals[np.isfinite(vals)]
    close = np.float32(1.5) + (rng.integers(0, 64, size=4096).astype(np.float32) * np.float32(2.0 ** -23))
    small = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, -1e-38, 3.4e38, -3.4e38], dtype=np.float32)
    pool = np.concatenate([vals, close, -close, small, rng.standard_normal(50000).astype(np.float32)])
    a = rng.choice(pool, size=n)
    b = rng.choice(pool, size=n)
    # a quarter of the pairs a few ulps apart (both sides of the 8-ulp key granularity), some exactly tied
    m = n // 4
    nb = (a[:m].view(np.uint32).astype(np.int64) + rng.integers(-12, 13, size=m)).clip(0, 2**32 - 1).astype(np.uint32).view(np.float32)
    b[:m] = np.where(np.isfinite(nb), nb, a[:m])
    la = rng.integers(0, 5, size=n)
    lb = rng.integers(0, 5, size=n)
    ka, kb = _keys(a, la), _keys(b, lb)
    assert np.isfinite(ka).all() and np.isfinite(kb).all()                   # finite scores never become NaN / inf keys
    assert ((ka.view(np.uint32) & 7) == la).all() and ((kb.view(np.uint32) & 7) == lb).all()
    near = ((ka.view(np.uint32) ^ kb.view(np.uint32)) < 8)
    assert near.sum() > 100 and (~near).sum() > 100
    far = ~near
    assert ((ka[far] > kb[far]) == (a[far] > b[far])).all()
    assert ((ka[far] < kb[far]) == (a[far] < b[far])).all()
    # an exact tie (also -0 against +0) is always "near": it can never be ordered silently by the label bits
    tie = a == b
    assert near[tie].all()
    z = _keys(np.array([0.0, -0.0], dtype=np.float32), [3, 1])
    assert ((z.view(np.uint32)[0] ^ z.view(np.uint32)[1]) < 8)


def test_quad_histogram_table():
    """The 625-entry table of the kernel: quad code l0 + 5 l1 + 25 l2 + 125 l3 -> one count byte per label level 1..4,
    and the fp32 fma chain that forms the code is exact."""
    import numpy as np
    lut = np.zeros(625, dtype=np.uint32)
    for t in range(125):                                                     # as the kernel's 125 threads fill it
        l0, l1, l2 = t % 5, (t // 5) % 5, t // 25
        h = sum((1 << (8 * (l - 1))) if l else 0 for l in (l0, l1, l2))
        lut[t] = h
        for l3 in range(1, 5):
            lut[t + 125 * l3] = h + (1 << (8 * (l3 - 1)))
    rng = np.random.default_rng(6)
    for _ in range(2000):
        y = rng.integers(0, 5, size=4).astype(np.float32)
        code = np.float32(y[0] + np.float32(8388608.0))
        for c, w in ((1, 5.0), (2, 25.0), (3, 125.0)):
            code = np.float32(np.float32(y[c]) * np.float32(w) + code)
        idx = int(code.view(np.uint32)) & 1023
        assert idx == int(y[0] + 5 * y[1] + 25 * y[2] + 125 * y[3])
        want = sum(int((y == lv).sum()) << (8 * (lv - 1)) for lv in range(1, 5))
        assert int(lut[idx]) == want
