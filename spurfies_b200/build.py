"""Builds spurfies_b200/csrc/libspurfies_b200.so in-tree with nvcc for sm_100a (no torch headers: seconds per file)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libspurfies_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# per-file flags: the scan / geometry kernels must round like the reference's separate fp32 ops
SOURCES = {
    "grid.cu": ["-fmad=false"],
    "render.cu": ["-fmad=false"],
    "mlp_f32.cu": [],
    "mlp_tc.cu": [],
    "mlp_tc2.cu": [],
    "optim.cu": [],
    "mesh.cu": ["-fmad=false"],
    "ingest.cu": ["-fmad=false"],
    "local_loss.cu": [],
}


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libspurfies_b200.so cannot be built (there is no CPU fallback)")


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, s) for s in srcs] + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(CSRC), "..", "include", "spurfies_b200.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    nvcc = _nvcc()
    objs = []
    for s in srcs:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        src = os.path.join(CSRC, s)
        if force or not os.path.exists(obj) or any(os.path.getmtime(obj) < os.path.getmtime(d) for d in deps if not d.endswith(".cu") or d == src):
            dbg = ["-DSPF_TIMELINE"] if os.environ.get("SPF_TIMELINE") == "1" else []
            cmd = [nvcc, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", *SOURCES[s], *dbg, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
