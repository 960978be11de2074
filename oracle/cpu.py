"""ctypes binding of the C CPU oracle (oracle/libzkw_oracle.so).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.

Arrays are numpy uint64 with a trailing limb axis: Fr/Fq vectors (n, 4), affine points (n, 8),
Jacobian points (n, 12); everything in halo2curves' Montgomery in-memory form.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzkw_oracle.so")
_SRCS = ["bn254.c", "msm.c", "ntt.c", "quotient.c", "prover.c", "zkw_oracle.h", "bn254_internal.h", "field_impl.h"]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (Makefile in this directory) if missing or stale."""
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.exists(os.path.join(_HERE, s)) and os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SRCS)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "libzkw_oracle.so"], check=True)
    return _LIB_PATH


_lib = None
u64p = C.POINTER(C.c_uint64)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.zko_max_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(u64p)


def max_threads() -> int:
    return int(lib().zko_max_threads())


# ---- field helpers -------------------------------------------------------------------------
def int_to_limbs(x: int) -> np.ndarray:
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def limbs_to_int(a) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(a))


def fr_to_mont(vals) -> np.ndarray:
    """list of python ints (canonical) -> (n,4) Montgomery"""
    a = np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    lib().zko_fr_vec_to_mont(_p(out), _p(a), C.c_size_t(a.shape[0]))
    return out


def fr_from_mont(a: np.ndarray) -> list[int]:
    a = np.ascontiguousarray(a.reshape(-1, 4))
    out = np.empty_like(a)
    lib().zko_fr_vec_from_mont(_p(out), _p(a), C.c_size_t(a.shape[0]))
    return [limbs_to_int(r) for r in out]


def fq_to_mont_one(v: int) -> np.ndarray:
    a = int_to_limbs(v)
    out = np.empty(4, dtype=np.uint64)
    lib().zko_fq_to_mont(_p(out), _p(a))
    return out


def fq_from_mont_one(a) -> int:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
    out = np.empty(4, dtype=np.uint64)
    lib().zko_fq_from_mont(_p(out), _p(a))
    return limbs_to_int(out)


def fr_random(n: int, seed: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().zko_fr_random(_p(out), C.c_size_t(n), C.c_uint64(seed))
    return out


def fr_mul(a, b) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64)
    lib().zko_fr_mul(_p(out), _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)))
    return out


# ---- G1 ---------------------------------------------------------------------------------------
def g1_generator() -> np.ndarray:
    out = np.empty(8, dtype=np.uint64)
    lib().zko_g1_generator(_p(out))
    return out


def g1_affine_to_ints(xy) -> tuple[int, int] | None:
    xy = np.asarray(xy, dtype=np.uint64)
    x, y = fq_from_mont_one(xy[:4]), fq_from_mont_one(xy[4:])
    return None if (x == 0 and y == 0) else (x, y)


def g1_ints_to_affine(pt) -> np.ndarray:
    if pt is None:
        return np.zeros(8, dtype=np.uint64)
    return np.concatenate([fq_to_mont_one(pt[0]), fq_to_mont_one(pt[1])])


def g1_to_affine(xyz: np.ndarray) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz.reshape(-1, 12))
    out = np.empty((xyz.shape[0], 8), dtype=np.uint64)
    lib().zko_g1_batch_to_affine(_p(out), _p(xyz), C.c_size_t(xyz.shape[0]))
    return out


def g1_is_on_curve(xy) -> bool:
    return bool(lib().zko_g1_is_on_curve(_p(np.ascontiguousarray(xy))))


def g1_fixed_base_mul(scalars: np.ndarray, threads: int = 0) -> np.ndarray:
    scalars = np.ascontiguousarray(scalars.reshape(-1, 4))
    out = np.empty((scalars.shape[0], 8), dtype=np.uint64)
    lib().zko_g1_fixed_base_mul(_p(out), _p(scalars), C.c_size_t(scalars.shape[0]), C.c_int(threads))
    return out


def srs_powers(n: int, tau_mont: np.ndarray, threads: int = 0) -> np.ndarray:
    out = np.empty((n, 8), dtype=np.uint64)
    lib().zko_srs_powers(_p(out), C.c_size_t(n), _p(np.ascontiguousarray(tau_mont)), C.c_int(threads))
    return out


# ---- MSM / NTT ----------------------------------------------------------------------------------
def best_multiexp(scalars: np.ndarray, bases: np.ndarray, threads: int = 0) -> np.ndarray:
    scalars = np.ascontiguousarray(scalars.reshape(-1, 4))
    bases = np.ascontiguousarray(bases.reshape(-1, 8))
    assert scalars.shape[0] == bases.shape[0]
    out = np.empty(12, dtype=np.uint64)
    lib().zko_best_multiexp(_p(out), _p(scalars), _p(bases), C.c_size_t(scalars.shape[0]), C.c_int(threads))
    return out


def msm_naive(scalars: np.ndarray, bases: np.ndarray) -> np.ndarray:
    scalars = np.ascontiguousarray(scalars.reshape(-1, 4))
    bases = np.ascontiguousarray(bases.reshape(-1, 8))
    out = np.empty(12, dtype=np.uint64)
    lib().zko_msm_naive(_p(out), _p(scalars), _p(bases), C.c_size_t(scalars.shape[0]))
    return out


def best_fft(a: np.ndarray, omega_mont: np.ndarray, threads: int = 0) -> np.ndarray:
    a = np.array(a.reshape(-1, 4), dtype=np.uint64, order="C", copy=True)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    lib().zko_best_fft(_p(a), C.c_uint(log_n), _p(np.ascontiguousarray(omega_mont)), C.c_int(threads))
    return a


class Domain(C.Structure):
    _fields_ = [
        ("k", C.c_uint), ("ext_k", C.c_uint), ("quotient_poly_degree", C.c_uint),
        ("omega", C.c_uint64 * 4), ("omega_inv", C.c_uint64 * 4),
        ("ext_omega", C.c_uint64 * 4), ("ext_omega_inv", C.c_uint64 * 4),
        ("g_coset", C.c_uint64 * 4), ("g_coset_inv", C.c_uint64 * 4),
        ("ifft_divisor", C.c_uint64 * 4), ("ext_ifft_divisor", C.c_uint64 * 4),
        ("t_evaluations", (C.c_uint64 * 4) * 16),
    ]

    @classmethod
    def new(cls, cs_degree: int, k: int) -> "Domain":
        d = cls()
        lib().zko_domain_new(C.byref(d), C.c_uint(cs_degree), C.c_uint(k))
        return d

    def arr(self, name: str) -> np.ndarray:
        return np.array(list(getattr(self, name)), dtype=np.uint64)

    def lagrange_to_coeff(self, a, threads=0):
        a = np.array(a.reshape(-1, 4), dtype=np.uint64, order="C", copy=True)
        assert a.shape[0] == 1 << self.k
        lib().zko_lagrange_to_coeff(C.byref(self), _p(a), C.c_int(threads))
        return a

    def coeff_to_lagrange(self, a, threads=0):
        a = np.array(a.reshape(-1, 4), dtype=np.uint64, order="C", copy=True)
        assert a.shape[0] == 1 << self.k
        lib().zko_coeff_to_lagrange(C.byref(self), _p(a), C.c_int(threads))
        return a

    def coeff_to_extended(self, a, threads=0):
        a = np.ascontiguousarray(a.reshape(-1, 4))
        assert a.shape[0] == 1 << self.k
        out = np.empty((1 << self.ext_k, 4), dtype=np.uint64)
        lib().zko_coeff_to_extended(C.byref(self), _p(a), _p(out), C.c_int(threads))
        return out

    def extended_to_coeff(self, a, threads=0):
        a = np.array(a.reshape(-1, 4), dtype=np.uint64, order="C", copy=True)
        assert a.shape[0] == 1 << self.ext_k
        lib().zko_extended_to_coeff(C.byref(self), _p(a), C.c_int(threads))
        return a

    def divide_by_vanishing_poly(self, a, threads=0):
        a = np.array(a.reshape(-1, 4), dtype=np.uint64, order="C", copy=True)
        lib().zko_divide_by_vanishing_poly(C.byref(self), _p(a), C.c_int(threads))
        return a


# ---- quotient ---------------------------------------------------------------------------------
class CircuitShape(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("k", "ext_k", "num_advice", "num_lookup_advice", "num_fixed", "blinding_factors", "cs_degree", "reserved")]


class QuotientInputs(C.Structure):
    _fields_ = [
        ("shape", CircuitShape),
        ("advice", C.POINTER(u64p)), ("constants", C.POINTER(u64p)), ("table", u64p),
        ("q_enable", C.POINTER(u64p)), ("q_lookup", u64p), ("sigma", C.POINTER(u64p)),
        ("perm_z", C.POINTER(u64p)), ("lookup_z", C.POINTER(u64p)), ("lookup_a", C.POINTER(u64p)),
        ("lookup_s", C.POINTER(u64p)), ("l0", u64p), ("l_last", u64p), ("l_active", u64p),
        ("y", C.c_uint64 * 4), ("beta", C.c_uint64 * 4), ("gamma", C.c_uint64 * 4), ("theta", C.c_uint64 * 4),
    ]


def make_shape(k: int, num_advice: int, num_lookup_advice: int, num_fixed: int = 1, blinding_factors: int = 6) -> CircuitShape:
    """Mirror of what ECDSACircuit::configure yields (ecdsa_p256.rs:94-115): cs_degree is 4 with a
    dedicated lookup advice column and 5 in selector mode (q_lookup * a); ext_k from
    EvaluationDomain::new."""
    selector_mode = num_lookup_advice == 0
    deg = 5 if selector_mode else 4
    ek = k
    while (1 << ek) < (1 << k) * (deg - 1):
        ek += 1
    return CircuitShape(k, ek, num_advice, num_lookup_advice, num_fixed, blinding_factors, deg, 0)


def _ptr_array(arrs):
    t = (u64p * max(len(arrs), 1))()
    for i, a in enumerate(arrs):
        t[i] = _p(a)
    return t


def build_quotient_inputs(shape: CircuitShape, cols: dict, challenges: dict, ptr_of=None):
    """cols: dict of name -> array or list of arrays (numpy (2^ext_k,4) or anything ptr_of maps to
    an address).  Returns (struct, keepalive)."""
    if ptr_of is None:
        ptr_of = _p
    keep = []

    def table(name):
        lst = cols.get(name) or []
        t = (u64p * max(len(lst), 1))()
        for i, a in enumerate(lst):
            t[i] = ptr_of(a)
        keep.append(t)
        return C.cast(t, C.POINTER(u64p))

    def single(name):
        a = cols.get(name)
        return ptr_of(a) if a is not None else C.cast(None, u64p)

    q = QuotientInputs()
    q.shape = shape
    q.advice = table("advice")
    q.constants = table("constants")
    q.table = single("table")
    q.q_enable = table("q_enable")
    q.q_lookup = single("q_lookup")
    q.sigma = table("sigma")
    q.perm_z = table("perm_z")
    q.lookup_z = table("lookup_z")
    q.lookup_a = table("lookup_a")
    q.lookup_s = table("lookup_s")
    q.l0 = single("l0")
    q.l_last = single("l_last")
    q.l_active = single("l_active")
    for name in ("y", "beta", "gamma", "theta"):
        v = np.asarray(challenges[name], dtype=np.uint64)
        getattr(q, name)[:] = [int(x) for x in v]
    keep.append(cols)
    return q, keep


def quotient_ecdsa(shape: CircuitShape, cols: dict, challenges: dict, threads: int = 0) -> np.ndarray:
    q, keep = build_quotient_inputs(shape, cols, challenges)
    out = np.empty((1 << shape.ext_k, 4), dtype=np.uint64)
    rc = lib().zko_quotient_ecdsa(C.byref(q), _p(out), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"zko_quotient_ecdsa failed: {rc}")
    del keep
    return out
