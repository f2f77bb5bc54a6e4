"""In-tree build of libsdfr.so for sm_100a (nvcc cross-compiles without a GPU).

The .so lives next to this file so that it travels to the GPU box with the
repository snapshot and shows up as an in-tree native library when loaded.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SDFR_BUILD_TAG=<tag> builds a variant (with SDFR_NVCC_FLAGS) next to the product library: libsdfr_<tag>.so, loaded
# with SDFR_LIB for A/B measurements
_TAG = os.environ.get("SDFR_BUILD_TAG", "")
OBJ_DIR = os.path.join(HERE, "build", _TAG) if _TAG else os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, f"libsdfr_{_TAG}.so" if _TAG else "libsdfr.so")
SOURCES = ["api.cu", "mlp_ffma.cu", "mlp_tc.cu", "surface.cu", "splat.cu", "circle.cu", "loss.cu", "refine.cu", "trace.cu", "pose.cu", "rotate_iou.cu", "np_random.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; libsdfr.so cannot be built")


def _deps(src: str):
    deps = [os.path.join(CSRC, src)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "include", "sdfr.h"))
    return deps


def _stale(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every CUDA source and links libsdfr.so; returns its path."""
    nvcc = _nvcc()
    extra = os.environ.get("SDFR_NVCC_FLAGS", "").split()   # e.g. -DSDFR_TC_PROFILE for the in-kernel timers
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, _deps(src)):
            jobs.append([nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + p.stdout + p.stderr)
        return p

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2)) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB_PATH, objs):
        run([nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
