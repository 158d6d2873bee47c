"""Build the sm_100a CUDA library in-tree with nvcc (no torch headers, no JIT cache).

    python -m ltr_b200.build [--force] [--verbose]

Produces ltr_b200/lib/libltr_b200.so from csrc/*.cu.  nvcc cross-compiles without a GPU, so
this runs in the build container; the resulting .so travels to the GPU box with the tree.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libltr_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    root = os.path.dirname(os.path.dirname(HERE))
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
                  + glob.glob(os.path.join(root, "include", "*.h")))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _sources() + _headers())


def _compile_one(nvcc, src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    newest_dep = max(os.path.getmtime(f) for f in [src] + _headers())
    if os.path.exists(obj) and os.path.getmtime(obj) > newest_dep:
        return obj, ""
    cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A/B experiments: compile every source with extra -D flags into lib/variants/<name>/libltr_b200.so."""
    nvcc = _nvcc()
    out_dir = os.path.join(LIB_DIR, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for src in _sources():
        obj = os.path.join(out_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        objs.append(obj)
    lib = os.path.join(out_dir, "libltr_b200.so")
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJ_DIR, "*.o")):
            os.remove(f)
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    if verbose:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    objs = [o for o, _ in results]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
