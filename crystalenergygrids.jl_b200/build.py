"""Build recipe of libceg_cuda.so: plain nvcc, sm_100a only, in-tree output (the .so travels to the
GPU box with the repo snapshot; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libceg_cuda.so")
SOURCES = ["ceg_cuda.cu"]
HEADERS = ["common.cuh", "grid_build.cuh", "interp.cuh", "ewald.cuh", "fastmath.cuh", "math_coeffs.cuh", "guest.cuh", "latency.cuh"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=true", "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.access(c, os.X_OK):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "ceg_cuda.h"),
                                                                os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    extra = os.environ.get("CEG_NVCC_DEFS", "").split()     # experiment switches, e.g. -DCEG_EXP_TABLE=1
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
