"""The C-ABI library loads without a GPU, exports every symbol include/zkw_b200.h declares, and refuses
to run without a device (no CPU fallback).  No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zkw_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(zkw_[a-z0-9_]+)\s*\(", hdr))
    return {n for n in names if not n.startswith("zkw_shape_")}  # static inline helpers


def test_library_exports_every_declared_symbol(zkw):
    lib = zkw.load_library()
    declared = _declared_symbols()
    assert declared, "header parse failed"
    missing = sorted(s for s in declared if not hasattr(lib, s))
    assert not missing, missing
    assert declared == set(zkw.EXPORTS), sorted(declared ^ set(zkw.EXPORTS))


def test_strerror(zkw):
    lib = zkw.load_library()
    assert lib.zkw_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.zkw_strerror(-1)


def test_no_device_means_error_not_fallback(zkw):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(zkw.ZkwError) as ei:
        zkw.Context(0)
    assert ei.value.status == -1


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under webauthn-halo2_b200/ may reference it."""
    pkg = os.path.join(ROOT, "webauthn-halo2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    s = line.strip()
                    uses = (re.match(r"(from|import)\s+\.*oracle\b", s) or re.match(r"from\s+\S*\boracle\b", s)
                            or (s.startswith("#include") and "oracle" in s) or "libzkw_oracle" in s or "zko_" in s
                            or re.search(r"import_module\([^)]*oracle", s))
                    assert not uses, (os.path.join(dirpath, f), line)


def test_shape_helpers(zkw):
    s = zkw.CircuitShape.from_config(19, 1, 1, 1)
    assert (s.k, s.ext_k, s.cs_degree, s.perm_columns, s.perm_sets, s.lookups) == (19, 21, 5, 2, 1, 1)
    s = zkw.CircuitShape.from_config(17, 4, 1, 1)
    assert (s.k, s.ext_k, s.cs_degree, s.perm_columns, s.perm_sets, s.lookups) == (17, 19, 4, 6, 3, 1)
    s = zkw.CircuitShape.from_config(11, 291, 53, 4)
    assert (s.ext_k, s.perm_columns, s.perm_sets, s.lookups) == (13, 348, 174, 53)


def test_header_is_plain_c_and_links_from_c(tmp_path, zkw):
    """gcc -std=c11 compiles a C caller against include/zkw_b200.h and links it with the library alone: what a
    cgo / Rust `cc` / JNI shim on the reference side would do."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    zkw.load_library()
    libdir = os.path.join(ROOT, "webauthn-halo2_b200")
    exe = str(tmp_path / "ffi_caller")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "host", "ffi_caller.c"), "-L", libdir, "-l:libzkw_b200.so", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "FFI_CALLER OK" in res.stdout
