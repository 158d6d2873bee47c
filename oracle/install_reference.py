"""Puts the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored; it travels to the GPU box with the tree
like the built .so does) so that bench.py's CPU arm (`cpu_baseline`, `--impl reference`) can time the reference's own code
instead of the oracle port:

    python oracle/install_reference.py            # in the build container, where /root/reference exists

Byte-for-byte copies of losses/ (approxNDCG.py, lambdaL.py, listnet.py, riskLosses/, ...) and utils/metrics.py; nothing is
committed, nothing under /root/reference is touched.  `__graft_entry__.build()` runs this when /root/reference is present.
SURVEY 8(d): "CPU baseline beside it: import the reference from baseline/_ref/"; the reference has no setup.py, so the
`pip install --target baseline/_ref` of the base contract does not apply -- a copy of the files is the same thing.
Test infrastructure only: nothing in the package imports baseline/_ref."""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LTR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "losses")):
        return False
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(os.path.join(SRC, "losses"), os.path.join(DST, "losses"), dirs_exist_ok=True,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DST, "utils"), exist_ok=True)
    shutil.copy2(os.path.join(SRC, "utils", "metrics.py"), os.path.join(DST, "utils", "metrics.py"))
    for rel in ("losses/approxNDCG.py", "losses/lambdaL.py", "losses/listnet.py", "utils/metrics.py"):
        assert filecmp.cmp(os.path.join(SRC, rel), os.path.join(DST, rel), shallow=False), rel
    if verbose:
        print("reference files copied to", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if install(verbose=True) else 1)
