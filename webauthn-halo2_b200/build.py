"""Builds libzkw_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzkw_b200.so")


def build(verbose: bool = False, jobs: int | None = None) -> str:
    jobs = jobs or min(8, os.cpu_count() or 1)
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc build of libzkw_b200.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    if not os.path.exists(LIB):
        raise RuntimeError("build finished but libzkw_b200.so is missing")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
