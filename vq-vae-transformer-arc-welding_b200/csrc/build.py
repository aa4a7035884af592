"""Builds libvqb200.so (in-tree, next to the package) with nvcc for sm_100a only.

    python vq-vae-transformer-arc-welding_b200/csrc/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libvqb200.so")
SOURCES = ["vq_api.cu", "vq_fwd_fma.cu", "vq_fwd_tc.cu", "vq_fwd_tcs.cu", "vq_bwd.cu", "vq_hostpipe.cu", "tok_linear.cu", "vq_pack.cu", "vq_dedupe.cu", "enc_chain.cu"]
HEADERS = ["vq_common.cuh", "vq_ptx.cuh", os.path.join("..", "..", "include", "vqb200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--fmad=true",            # explicit fmaf/__f*_rn everywhere the rounding is part of the contract
    "-Xptxas", "-v",
]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def build(force: bool = False, verbose: bool = False, out: str = OUT, extra_flags=()) -> str:
    """out / extra_flags: experiment builds (tools/ab_build.py) next to the product library."""
    if out == OUT and not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + list(extra_flags) + ["-o", out] + [os.path.join(HERE, s) for s in SOURCES]
    host_cxx = "/usr/bin/g++"
    if os.path.exists(host_cxx):
        cmd[1:1] = ["-ccbin", host_cxx]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log" if out == OUT else os.path.basename(out) + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libvqb200.so")
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
