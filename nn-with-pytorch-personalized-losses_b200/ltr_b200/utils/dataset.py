"""Slate ingestion -- drop-in for reference utils/dataset.py:33-101 (get_data, get_baseline_data) plus the tensor
form the loss kernels consume.  svmlight / LETOR text and the reference's baseline-score CSV are parsed by the
library's host threads (`ltr_svmlight_load`, `ltr_baseline_csv_load`: mmap + multi-threaded parse straight into
pinned buffers) into query-major PADDED arrays:

    X [Q, n, F] fp32 (zero rows beyond a query's docs), y [Q, n] fp32 (-1 padding), baselines [Q, n, S] fp32

Queries are the runs of consecutive lines with the same qid, exactly as the reference groups them
(dataset.py:52-65).  `rank` / `world` give every rank its contiguous query range: only those lines
are materialised.  The reference requires equal-length queries (`torch.tensor(X_train)`, main_batch_execution.py:79)
and never pads; here shorter queries are padded with the loss functions' padded_value_indicator.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from .. import _lib
from ..sharding import shard_bounds

PADDED_Y_VALUE = -1.0


class svmDataset:
    """Paths and shape facts of a LETOR-style collection (dataset.py:9-30): same attribute names as the reference."""

    def __init__(self, dataset_name):
        self.train_data_path = self.test_data_path = self.vali_data_path = None
        self.baseline_train_data_path = self.baseline_test_data_path = self.baseline_vali_data_path = None
        self.docs_per_query = 1000
        self.queries_on_test = 10
        self.queries_on_train = 10
        if "2003_td" in dataset_name:
            self.num_features = 64
            self.normalized_num_docs = True
        elif "web10k" in dataset_name:
            self.num_features = 136
            self.normalized_num_docs = True


def scan_svmlight(path, threads=0):
    """{lines, queries, max_docs, min_index, max_index} of an svmlight file (one pass, no arrays)."""
    counts = (ctypes.c_longlong * 5)()
    _lib.check(_lib.load().ltr_svmlight_scan(os.fsencode(path), int(threads), counts), "ltr_svmlight_scan")
    return dict(lines=counts[0], queries=counts[1], max_docs=counts[2], min_index=counts[3], max_index=counts[4])


def _host_buffer(shape, dtype, pin):
    t = torch.empty(shape, dtype=dtype)
    if pin and torch.cuda.is_available():
        t = t.pin_memory()
    return t


def load_slates(path, n_docs=None, n_features=None, index_base=-1, pad_value=PADDED_Y_VALUE, rank=0, world=1, pin=True,
                threads=0):
    """svmlight file -> (X [Q_r, n, F], y [Q_r, n], qids [Q_r] int64, lens [Q_r] int32) host tensors (pinned when a GPU
    is present) holding this rank's queries.  n_docs defaults to the longest query of the file (longer queries are
    truncated to n_docs, `lens` keeps their true length); n_features to the largest feature index (sklearn's
    inference); index_base -1 = sklearn's "auto"."""
    info = scan_svmlight(path, threads)
    if info["queries"] == 0:
        raise ValueError(f"{path}: no data lines")
    zero_based = info["min_index"] == 0 if index_base < 0 else index_base == 0
    n = int(n_docs) if n_docs else int(info["max_docs"])
    F = int(n_features) if n_features else int(info["max_index"] + (1 if zero_based else 0))
    lo, hi = shard_bounds(int(info["queries"]), rank, world)
    Q = hi - lo
    X = _host_buffer((Q, n, F), torch.float32, pin)
    y = _host_buffer((Q, n), torch.float32, pin)
    qids = torch.empty(Q, dtype=torch.int64)
    lens = torch.empty(Q, dtype=torch.int32)
    if Q:
        _lib.check(_lib.load().ltr_svmlight_load(os.fsencode(path), lo, Q, n, F, 0 if zero_based else 1, float(pad_value),
                                                 X.data_ptr(), y.data_ptr(), qids.data_ptr(), lens.data_ptr(), int(threads)),
                   "ltr_svmlight_load")
    return X, y, qids, lens


def load_baselines(path, n_docs=None, rank=0, world=1, pin=True):
    """baseline CSV ("qid,score_1,...,score_S" per doc) -> (scores [Q_r, n, S], lens [Q_r]) for this rank's queries."""
    counts = (ctypes.c_longlong * 4)()
    lib = _lib.load()
    _lib.check(lib.ltr_baseline_csv_scan(os.fsencode(path), counts), "ltr_baseline_csv_scan")
    if counts[1] == 0:
        raise ValueError(f"{path}: no data lines")
    n = int(n_docs) if n_docs else int(counts[2])
    S = int(counts[3])
    lo, hi = shard_bounds(int(counts[1]), rank, world)
    Q = hi - lo
    scores = _host_buffer((Q, n, S), torch.float32, pin)
    lens = torch.empty(Q, dtype=torch.int32)
    if Q:
        _lib.check(lib.ltr_baseline_csv_load(os.fsencode(path), lo, Q, n, S, scores.data_ptr(), lens.data_ptr()),
                   "ltr_baseline_csv_load")
    return scores, lens


def _path_of(info_dataset, type_file, baseline=False):
    prefix = "baseline_" if baseline else ""
    if type_file not in ("train", "test", "vali"):
        raise UnboundLocalError("local variable 'data' referenced before assignment")      # dataset.py:37-42
    return getattr(info_dataset, f"{prefix}{type_file}_data_path")


def get_data(info_dataset, type_file="train"):
    """Reference signature and return value (dataset.py:33-65): (features_docs_by_query, labels_by_query) as per-query
    lists -- features as float32 vectors of the file's inferred width, labels as floats -- unpadded."""
    X, y, _, lens = load_slates(_path_of(info_dataset, type_file), pin=False)
    Xn, yn = X.numpy(), y.numpy().astype(np.float64)
    feats = [[Xn[q, d] for d in range(int(lens[q]))] for q in range(X.shape[0])]
    labels = [[yn[q, d] for d in range(int(lens[q]))] for q in range(X.shape[0])]
    return feats, labels


def get_baseline_data(info_dataset, type_file="train"):
    """Reference signature and return value (dataset.py:137-149): per query, the list of per-doc float32 score vectors."""
    scores, lens = load_baselines(_path_of(info_dataset, type_file, baseline=True), pin=False)
    sn = scores.numpy()
    return [[sn[q, d] for d in range(int(lens[q]))] for q in range(scores.shape[0])]
