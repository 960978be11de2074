"""Compiles tests/host/host_check.cu with nvcc as HOST code and runs it: the product's
__host__ __device__ field and curve formulas (field.cuh / curve.cuh, portable path) against the CPU
oracle.  The device (PTX carry-chain) path of the same functions is covered by the -m gpu parity tests."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_host_formulas(tmp_path, oracle):
    exe = str(tmp_path / "host_check")
    cmd = ["nvcc", "-O2", "-std=c++17", "-x", "cu", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "webauthn-halo2_b200", "csrc"),
           "-o", exe, os.path.join(ROOT, "tests", "host", "host_check.cu"), "-L", os.path.join(ROOT, "oracle"), "-lzkw_oracle",
           "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle"), "-Xcompiler", "-fopenmp"]
    subprocess.run(cmd, check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "HOST_CHECK OK" in res.stdout
