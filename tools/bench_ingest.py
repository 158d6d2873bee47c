"""Slate ingestion timed on the host: the library's multi-threaded svmlight parser (`load_slates`: mmap, scan,
parse straight into padded [Q, n, F] / [Q, n] arrays) against the reference's procedure restated (utils/dataset.py:33-65:
sklearn's load_svmlight_file + a Python loop that groups rows by change of qid + padding), on a synthetic MSLR-WEB10K-shaped
file (136 dense features per line, 1-based indices).  CPU only; prints one JSON line.

    python tools/bench_ingest.py [queries] [threads]
"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import numpy as np
from ltr_b200.utils import dataset as D

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 600
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
F = 136
rng = np.random.default_rng(1)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "Norm.train.txt")
lines = 0
with open(path, "w") as f:
    for q in range(Q):
        docs = int(rng.integers(40, 201))
        X = rng.random((docs, F)).astype(np.float32)
        y = rng.integers(0, 5, docs)
        for d in range(docs):
            f.write(f"{y[d]} qid:{q + 1} " + " ".join(f"{i + 1}:{X[d, i]:.6f}" for i in range(F)) + "\n")
        lines += docs
size = os.path.getsize(path)


def reference():
    from sklearn.datasets import load_svmlight_file
    X, y, qid = load_svmlight_file(path, query_id=True)
    X = X.toarray().astype(np.float32)
    feats, labels, cur_f, cur_l = [], [], [], []
    for i in range(qid.size):
        if i and qid[i] != qid[i - 1]:
            feats.append(cur_f), labels.append(cur_l)
            cur_f, cur_l = [], []
        cur_f.append(X[i]), cur_l.append(y[i])
    feats.append(cur_f), labels.append(cur_l)
    n = max(map(len, feats))
    Xp = np.zeros((len(feats), n, F), np.float32)
    yp = np.full((len(feats), n), -1.0, np.float32)
    for q, (a, b) in enumerate(zip(feats, labels)):
        Xp[q, :len(a)] = np.stack(a)
        yp[q, :len(b)] = b
    return Xp, yp


def timed(fn, reps):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best, out


t_ours, (Xo, yo, _, _) = timed(lambda: D.load_slates(path, pin=False, **({"threads": threads} if threads else {})), 3)
t_ref, (Xr, yr) = timed(reference, 1)
same = bool(np.array_equal(Xo.numpy(), Xr) and np.array_equal(yo.numpy(), yr))
print(json.dumps({"file_MB": size / 1e6, "lines": lines, "queries": Q, "features": F, "host_threads": threads or os.cpu_count(),
                  "load_slates_s": t_ours, "load_slates_MBps": size / 1e6 / t_ours, "reference_procedure_s": t_ref,
                  "reference_MBps": size / 1e6 / t_ref, "speedup": t_ref / t_ours, "bit_identical_to_reference_procedure": same}))
os.remove(path)
