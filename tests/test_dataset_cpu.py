"""Slate ingestion on the CPU: the library's svmlight / baseline-CSV parsers against the reference's own
procedure restated here -- sklearn's load_svmlight_file + grouping by CHANGE of qid (utils/dataset.py:33-101) -- on a
synthetic MSLR-shaped file (136 features, 1-based indices, ragged queries, comments, a qid that re-appears later)."""
import os

import numpy as np
import pytest
import torch

from ltr_b200.utils import dataset as D


def _write_collection(tmp_path, rng, n_queries=23, F=136, ragged=True, zero_based=False):
    qids, lines, base_lines = [], [], []
    for q in range(n_queries):
        qid = 100 + (q % 19)                              # qid 100..103 re-appear at the end: still separate queries
        docs = int(rng.integers(3, 40)) if ragged else 12
        for _ in range(docs):
            label = int(rng.integers(0, 5))
            idx = np.sort(rng.choice(F, size=int(rng.integers(1, F // 3)), replace=False)) + (0 if zero_based else 1)
            vals = rng.random(idx.size).astype(np.float32)
            feats = " ".join(f"{i}:{v:.6g}" for i, v in zip(idx, vals))
            lines.append(f"{label} qid:{qid} {feats} #docid = {rng.integers(1e6)}")
            base_lines.append(f"{qid},{rng.normal():.5f},{rng.normal():.5f}")
            qids.append(qid)
    lines.insert(5, "# a comment line")
    lines.insert(9, "")
    path = os.path.join(tmp_path, "Norm.train.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    bpath = os.path.join(tmp_path, "baseline.Norm.train.txt")
    with open(bpath, "w") as f:
        f.write("\n".join(base_lines) + "\n")
    return path, bpath


def _reference_procedure(path):
    """dataset.py:33-65, restated: load_svmlight_file(query_id=True), new query where the qid changes."""
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
    return feats, labels


@pytest.mark.parametrize("zero_based", [False, True])
def test_svmlight_matches_reference_procedure(tmp_path, zero_based):
    rng = np.random.default_rng(5 + zero_based)
    path, _ = _write_collection(str(tmp_path), rng, zero_based=zero_based)
    feats, labels = _reference_procedure(path)
    info = D.scan_svmlight(path)
    assert info["queries"] == len(feats) == 23 and info["max_docs"] == max(map(len, feats))
    X, y, qids, lens = D.load_slates(path, pin=False)
    assert X.shape == (23, info["max_docs"], feats[0][0].shape[0]) and y.shape == X.shape[:2]
    for q in range(23):
        m = len(feats[q])
        assert int(lens[q]) == m
        assert np.array_equal(X[q, :m].numpy(), np.stack(feats[q]))           # bit-exact: both parse with strtod -> fp32
        assert np.array_equal(y[q, :m].numpy(), np.asarray(labels[q], dtype=np.float32))
        assert float(y[q, m:].abs().min() if m < y.shape[1] else 1.0) == 1.0 and bool((y[q, m:] == -1).all())
        assert float(X[q, m:].abs().sum()) == 0.0
    assert int(qids[0]) == 100 and int(qids[19]) == 100                          # the same qid later = another query


def test_get_data_drop_in_and_sharding(tmp_path):
    rng = np.random.default_rng(9)
    path, bpath = _write_collection(str(tmp_path), rng)
    info = D.svmDataset("web10k")
    assert info.num_features == 136
    info.train_data_path, info.baseline_train_data_path = path, bpath
    feats, labels = D.get_data(info, "train")
    rf, rl = _reference_procedure(path)
    assert len(feats) == len(rf)
    for a, b, c, d in zip(feats, rf, labels, rl):
        assert len(a) == len(b) and all(np.array_equal(x, z) for x, z in zip(a, b)) and list(map(float, c)) == list(map(float, d))
    base = D.get_baseline_data(info, "train")
    assert [len(b) for b in base] == [len(f) for f in feats] and base[0][0].dtype == np.float32 and base[0][0].shape == (2,)
    with open(bpath) as f:
        first = f.readline().strip().split(",")
    assert np.allclose(base[0][0], np.asarray(first[1:], dtype=np.float32))
    # query-sharded: rank r of 3 holds its contiguous range, padded to a common n; together they are the whole file
    whole = D.load_slates(path, n_docs=48, pin=False)
    parts = [D.load_slates(path, n_docs=48, rank=r, world=3, pin=False) for r in range(3)]
    for k in range(4):
        assert torch.equal(torch.cat([p[k] for p in parts]), whole[k])
    sb = [D.load_baselines(bpath, n_docs=48, rank=r, world=3, pin=False)[0] for r in range(3)]
    assert torch.equal(torch.cat(sb), D.load_baselines(bpath, n_docs=48, pin=False)[0])
    # truncation keeps the first n docs and reports the true length
    Xs, ys, _, ls = D.load_slates(path, n_docs=5, pin=False)
    assert Xs.shape[1] == 5 and torch.equal(Xs, whole[0][:, :5]) and torch.equal(ls, whole[3])
    with pytest.raises(Exception):
        D.load_slates(os.path.join(str(tmp_path), "missing.txt"))


def test_file_of_whole_pages_and_no_trailing_newline(tmp_path):
    """A file whose size is a multiple of the page size (no zero tail behind a mapping) and without a final newline:
    strtod on the last value must not run off the end."""
    line = "2 qid:7 1:0.5 3:0.25"
    page = os.sysconf("SC_PAGESIZE")
    body = (line + "\n") * 300
    filler = (-(len(body) + len(line)) - 2) % page            # a comment line "#xxx\n" in front makes the size k * page
    text = "#" + "x" * filler + "\n" + body + line
    assert len(text) % page == 0
    path = os.path.join(str(tmp_path), "pages.txt")
    with open(path, "w") as f:
        f.write(text)
    X, y, qids, lens = D.load_slates(path, pin=False)
    assert X.shape == (1, 301, 3) and int(qids[0]) == 7 and int(lens[0]) == 301
    assert float(X[0, 0, 0]) == 0.5 and float(X[0, 300, 2]) == 0.25 and bool((y[0] == 2).all())


def test_number_forms_parse_like_strtod(tmp_path):
    """The parsers' fast decimal path (at most 15 significant digits, |exponent| <= 22: one exact IEEE operation) and its strtod
    fallback return what Python's float() -- correctly rounded, like strtod -- returns, for every number form a file may hold."""
    toks = ["0", "-0", "+3.5", ".5", "5.", "0.1", "0.30000000000000004", "1e-5", "1E5", "1e22", "1e23", "1e-22", "1e-23",
            "123456789012345", "1234567890123456", "123456789012345678901234567890", "0.000000000000000000001234",
            "4.9e-324", "1.7976931348623157e308", "2.2250738585072014e-308", "9007199254740993", "0.1e1", "00012.5000",
            "3.4028235e38", "1.17549435e-38", "7.006492321624085e-46", "1e400", "-1e-400", "16777217", "0.3333333333333333",
            "8.5e-1", "12345.678901234567", "1e+3", "2e-0", "inf", "-inf", "nan"]
    F = len(toks)
    line = "3 qid:1 " + " ".join(f"{i + 1}:{t}" for i, t in enumerate(toks))
    path = os.path.join(str(tmp_path), "numbers.txt")
    with open(path, "w") as f:
        f.write(line + "\n" + "1.5e0 qid:1 1:1\n")
    X, y, qids, lens = D.load_slates(path, n_features=F, index_base=1, pin=False)
    with np.errstate(over="ignore"):
        want = np.asarray([float(t) for t in toks], dtype=np.float64).astype(np.float32)
    got = X[0, 0].numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), [(t, g, w) for t, g, w in zip(toks, got, want) if not (g == w or (g != g and w != w))]
    assert float(y[0, 0]) == 3.0 and float(y[0, 1]) == 1.5
    bpath = os.path.join(str(tmp_path), "baseline.numbers.txt")
    with open(bpath, "w") as f:
        f.write("1, 0.25,-1e-3,7\n1,1e22, .5,-0\n")
    S = D.load_baselines(bpath, pin=False)
    S = S[0] if isinstance(S, (tuple, list)) else S
    assert np.array_equal(S[0, :2].numpy(), np.asarray([[0.25, -1e-3, 7.0], [1e22, 0.5, -0.0]], dtype=np.float32))


@pytest.mark.parametrize("trailing_newline", [True, False])
def test_parallel_line_split_equals_the_serial_one(tmp_path, trailing_newline):
    """Files above 1 MB are split into lines by several threads (a line belongs to the byte range its first byte lies in):
    the same slates as with one thread for 2..13 threads, on lines of very different lengths with comments, blank lines and
    CR LF endings in between."""
    rng = np.random.default_rng(11 + trailing_newline)
    rows = []
    for q in range(260):
        for _ in range(int(rng.integers(1, 30))):
            k = int(rng.integers(1, 136))
            idx = np.sort(rng.choice(136, size=k, replace=False)) + 1
            feats = " ".join(f"{i}:{rng.random():.6f}" for i in idx)
            rows.append(f"{int(rng.integers(0, 5))} qid:{q} {feats}" + (" # doc" if rng.random() < 0.1 else "") + ("\r" if rng.random() < 0.05 else ""))
        if rng.random() < 0.2:
            rows.append("# only a comment " + "x" * int(rng.integers(0, 3000)))
        if rng.random() < 0.2:
            rows.append("")
    text = "\n".join(rows) + ("\n" if trailing_newline else "")
    assert len(text) > (1 << 20)
    path = os.path.join(str(tmp_path), "big.txt")
    with open(path, "w", newline="") as f:
        f.write(text)
    one = D.load_slates(path, pin=False, threads=1)
    info1 = D.scan_svmlight(path, threads=1)
    for th in (2, 3, 8, 13):
        assert D.scan_svmlight(path, threads=th) == info1
        many = D.load_slates(path, pin=False, threads=th)
        for a, b in zip(one, many):
            assert torch.equal(a, b), th
    feats, labels = _reference_procedure(path)
    assert len(feats) == one[0].shape[0] and [len(x) for x in feats] == one[3].tolist()
