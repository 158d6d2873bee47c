"""pytest wiring: marker registration, import paths, golden-vector loader."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")
for p in (ROOT, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_terminal_summary(terminalreporter):
    """How often the fp64 slack of assert_loss_close / assert_grad_close was needed in this session."""
    try:
        from ltr_testing import SLACK_STATS as st
    except Exception:
        return
    if st["loss_checks_with_ref64"] or st["grad_checks_with_ref64"]:
        terminalreporter.write_line(
            "fp64-slack report: loss checks %d, of which %d would fail the plain 1e-5 gate; gradient checks %d, of which %d "
            "(%d of %d elements) would fail the plain 1e-4 gate" % (
                st["loss_checks_with_ref64"], st["loss_checks_that_needed_it"], st["grad_checks_with_ref64"],
                st["grad_checks_that_needed_it"], st["grad_elements_that_needed_it"], st["grad_elements_with_ref64"]))


class Golden:
    """Golden vectors produced by tests/golden/make_golden.py and make_golden_v2.py from the unmodified reference."""

    def __init__(self):
        here = os.path.join(ROOT, "tests", "golden")
        self.cases, self.arrays = {}, {}
        for version in ("golden_v1", "golden_v2"):        # v2: round-2 additions (make_golden_v2.py), v1 untouched
            with open(os.path.join(here, version + ".json")) as f:
                meta = json.load(f)
            npz = np.load(os.path.join(here, version + ".npz"))
            self.arrays.update({k: npz[k] for k in npz.files})
            self.cases.update({c["name"]: c for c in meta["cases"]})

    def names(self, prefix):
        return [n for n in self.cases if n.startswith(prefix)]

    def inputs(self, name):
        return {k: self.arrays[f"{name}/in/{k}"] for k in self.cases[name]["inputs"]}

    def out(self, name):
        return self.arrays[f"{name}/out"]

    def grad(self, name, key):
        return self.arrays[f"{name}/grad/{key}"]

    def has_grad(self, name, key):
        return key in self.cases[name]["grads"]

    def kwargs(self, name):
        return dict(self.cases[name]["kwargs"])


_GOLDEN = None


def golden():
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = Golden()
    return _GOLDEN


@pytest.fixture(scope="session")
def gold():
    return golden()
