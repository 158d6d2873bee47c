"""bench.py's CPU arm: when the unmodified reference files are beside the tree (baseline/_ref/, copied by
oracle/install_reference.py in the build container) the arm times THEM (kind "reference"), else the oracle port; either way
the two agree on the bench workload, so the baseline number means the same thing on both kinds of box."""
import importlib.util
import os
import sys

import pytest
import torch

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("ltr_bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_reference_copy_is_byte_identical_and_agrees_with_the_port():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import install_reference
    finally:
        sys.path.remove(os.path.join(ROOT, "oracle"))
    if not install_reference.install():
        if not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "losses", "approxNDCG.py")):
            pytest.skip("no reference checkout and no baseline/_ref copy on this box: the arm runs the port")
    b = _bench()
    assert b.cpu_kind() == "reference"
    ref = b.reference_module("losses/approxNDCG.py")
    from oracle import ltr_oracle as O
    s, y, _ = b.make_batch(b.SEED, B=16)
    a = s.clone().requires_grad_(True)
    c = s.clone().requires_grad_(True)
    la = ref.approxNDCGLoss(a, y, alpha=b.ALPHA)
    lc = O.approx_ndcg_loss(c, y, alpha=b.ALPHA)
    la.backward()
    lc.backward()
    assert torch.equal(la.detach(), lc.detach()) and torch.equal(a.grad, c.grad)       # the port restates it op for op
    m = b.reference_module("utils/metrics.py")
    g = torch.Generator().manual_seed(1)
    yt, sp = torch.randint(0, 5, (9, 40), generator=g).float(), torch.randn(9, 40, generator=g)
    assert torch.equal(m.torchNdcg(yt, sp, k=10, return_type="tensor"), O.torch_ndcg(yt, sp, k=10, return_type="tensor"))
