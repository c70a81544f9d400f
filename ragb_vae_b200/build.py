"""Builds librgbavae.so (the sm_100a kernel library) in-tree with nvcc.

    python ragb_vae_b200/build.py [--force] [-v]

The library is compiled for sm_100a only (``-gencode arch=compute_100a,code=sm_100a``); nvcc
cross-compiles without a GPU.  Objects go to ``ragb_vae_b200/csrc/build/``, the shared library to
``ragb_vae_b200/librgbavae.so`` (git-ignored, travels with the repo snapshot to the GPU box).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "librgbavae.so")
SOURCES = ["rv_common.cu", "rv_elementwise.cu", "rv_reduce.cu", "rv_conv_direct.cu", "rv_conv_tc.cu", "rv_conv_halo.cu", "rv_conv_out.cu", "rv_attention.cu", "rv_plumbing.cu", "rv_train.cu", "rv_conv_wgrad.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; librgbavae has no prebuilt or CPU fallback")


def _deps(src: str):
    yield os.path.join(CSRC, src)
    yield os.path.join(CSRC, "rv_common.cuh")
    yield os.path.join(CSRC, "rv_tc_common.cuh")
    yield os.path.join(os.path.dirname(PKG), "include", "rgbavae.h")
    yield os.path.abspath(__file__)


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in sources:
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, _deps(s)):
            cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(run, jobs):
            if verbose and out.strip():
                print(out)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in sources]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
