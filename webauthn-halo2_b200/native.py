"""ctypes binding of libzkw_b200.so — the C ABI of include/zkw_b200.h.

This is the host-side mirror of the seam a patched halo2_proofs would bind over Rust FFI (see
INTEGRATION.md): `Context.msm` <-> best_multiexp, `Context.ntt` <-> best_fft,
`lagrange_to_coeff / coeff_to_extended / extended_to_coeff` <-> the EvaluationDomain methods,
`Context.quotient` <-> evaluate_h + divide_by_vanishing_poly.  The reference reaches those through
create_proof / keygen at halo2-circuits/src/ecc/ecdsa_p256.rs:259-260, 366-373, 416-423, 555-562.

There is no CPU fallback: a missing library raises at import of the symbol table, and a missing
GPU raises ZkwError(ZKW_ERR_NO_DEVICE) from Context().

Array convention (numpy uint64, C-contiguous): Fr/Fq vectors (n, 4); affine points (n, 8);
Jacobian points (12,) — all in halo2curves' Montgomery in-memory form.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKW_B200_LIB") or os.path.join(HERE, "libzkw_b200.so")   # override: A/B builds of the same ABI

ZKW_OK = 0
ZKW_ERR_NO_DEVICE = -1
ZKW_ERR_CUDA = -2
ZKW_ERR_INVALID = -3
ZKW_ERR_OOM = -4
ZKW_ERR_STATE = -5
ZKW_ERR_UNSUPPORTED = -6
ZKW_ERR_SIGNATURE = -7

BASES_G, BASES_G_LAGRANGE, BASES_CALLER = 0, 1, 2

u64p = C.POINTER(C.c_uint64)

# every symbol include/zkw_b200.h declares (tests/test_abi.py checks the library exports them all)
EXPORTS = [
    "zkw_ctx_create", "zkw_ctx_destroy", "zkw_ctx_sync", "zkw_ctx_stream", "zkw_strerror", "zkw_last_cuda_error",
    "zkw_ctx_launch_count", "zkw_srs_load", "zkw_srs_load_dev", "zkw_msm_config", "zkw_msm_window_bits", "zkw_msm_bn254_g1",
    "zkw_msm_bn254_g1_dev", "zkw_msm_bn254_g1_dev_to_host", "zkw_g1_batch_normalize", "zkw_ntt_bn254_fr",
    "zkw_ntt_bn254_fr_dev", "zkw_lagrange_to_coeff", "zkw_lagrange_to_coeff_dev", "zkw_coeff_to_lagrange",
    "zkw_coeff_to_lagrange_dev", "zkw_coeff_to_extended", "zkw_coeff_to_extended_dev", "zkw_extended_to_coeff",
    "zkw_extended_to_coeff_dev", "zkw_quotient_ecdsa", "zkw_quotient_ecdsa_dev", "zkw_dev_alloc", "zkw_dev_free",
    "zkw_memcpy_h2d", "zkw_memcpy_d2h", "zkw_srs_setup", "zkw_srs_get", "zkw_g1_fixed_base_mul",
    "zkw_profile_enable", "zkw_profile_filter", "zkw_profile_reset", "zkw_profile_read", "zkw_profile_names",
    "zkw_keygen", "zkw_pk_destroy", "zkw_pk_info", "zkw_pk_vk", "zkw_create_proof", "zkw_create_proof_ex", "zkw_create_proof_seeded", "zkw_create_proof_overlapped", "zkw_fr_to_mont", "zkw_fr_from_mont",
    "zkw_synth_witness", "zkw_host_alloc", "zkw_host_free",
    "zkw_ecdsa_circuit_new", "zkw_ecdsa_circuit_free", "zkw_ecdsa_circuit_shape", "zkw_ecdsa_circuit_rows", "zkw_ecdsa_circuit_fixed",
    "zkw_ecdsa_circuit_permutation", "zkw_ecdsa_synthesize",
    "zkw_pk_write", "zkw_pk_read", "zkw_vk_write", "zkw_vk_read",
    "zkw_prover_create", "zkw_prover_destroy", "zkw_prover_ctx", "zkw_prover_pk", "zkw_prover_last_synthesis_ms", "zkw_prover_prove",
    "zkw_prove_batch", "zkw_g1_sum", "zkw_selftest_field",
]


class ZkwError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"{what}: status {status} ({_strerror(status)}){' — ' + detail if detail else ''}")


class CircuitShape(C.Structure):
    """zkw_circuit_shape: what ECDSACircuit::configure (ecdsa_p256.rs:94-115) yields for a config line."""
    _fields_ = [(n, C.c_uint32) for n in
                ("k", "ext_k", "num_advice", "num_lookup_advice", "num_fixed", "blinding_factors", "cs_degree", "reserved")]

    @classmethod
    def from_config(cls, degree: int, num_advice: int, num_lookup_advice: int, num_fixed: int = 1,
                    blinding_factors: int = 6) -> "CircuitShape":
        # halo2-lib's RangeConfig folds the lookup into the single gate column behind a q_lookup
        # selector when num_advice == 1 (lookup input q*a has degree 2 => cs degree 5)
        selector_mode = num_advice == 1
        lookup_cols = 0 if selector_mode else num_lookup_advice
        deg = 5 if selector_mode else 4
        ek = degree
        while (1 << ek) < (1 << degree) * (deg - 1):
            ek += 1
        return cls(degree, ek, num_advice, lookup_cols, num_fixed, blinding_factors, deg, 0)

    @property
    def perm_columns(self) -> int:
        return self.num_fixed + self.num_advice + self.num_lookup_advice

    @property
    def perm_sets(self) -> int:
        chunk = self.cs_degree - 2
        return (self.perm_columns + chunk - 1) // chunk

    @property
    def lookups(self) -> int:
        return self.num_lookup_advice or 1


class CircuitParamsC(C.Structure):
    """zkw_circuit_params: the reference's one-line JSON config (struct CircuitParams, ecdsa_p256.rs:53-63)."""
    _fields_ = [(n, C.c_uint32) for n in ("degree", "num_advice", "num_lookup_advice", "num_fixed", "lookup_bits", "limb_bits", "num_limbs")]


class QuotientInputs(C.Structure):
    _fields_ = [
        ("shape", CircuitShape),
        ("advice", C.POINTER(u64p)), ("constants", C.POINTER(u64p)), ("table", u64p),
        ("q_enable", C.POINTER(u64p)), ("q_lookup", u64p), ("sigma", C.POINTER(u64p)),
        ("perm_z", C.POINTER(u64p)), ("lookup_z", C.POINTER(u64p)), ("lookup_a", C.POINTER(u64p)),
        ("lookup_s", C.POINTER(u64p)), ("l0", u64p), ("l_last", u64p), ("l_active", u64p),
        ("y", C.c_uint64 * 4), ("beta", C.c_uint64 * 4), ("gamma", C.c_uint64 * 4), ("theta", C.c_uint64 * 4),
    ]


_lib = None


def load_library() -> C.CDLL:
    """Load libzkw_b200.so; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The B200 path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.zkw_strerror.restype = C.c_char_p
    lib.zkw_last_cuda_error.restype = C.c_char_p
    lib.zkw_last_cuda_error.argtypes = [C.c_void_p]
    lib.zkw_ctx_stream.restype = C.c_void_p
    lib.zkw_ctx_stream.argtypes = [C.c_void_p]
    lib.zkw_ctx_launch_count.restype = C.c_uint64
    lib.zkw_ctx_launch_count.argtypes = [C.c_void_p]
    lib.zkw_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.zkw_ctx_destroy.argtypes = [C.c_void_p]
    lib.zkw_ctx_destroy.restype = None
    lib.zkw_profile_read.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.zkw_profile_names.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.zkw_profile_filter.argtypes = [C.c_void_p, C.c_char_p]
    lib.zkw_selftest_field.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.POINTER(C.c_uint)]
    lib.zkw_synth_witness.argtypes = [C.POINTER(CircuitShape), C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(u64p), C.POINTER(C.c_size_t)]
    lib.zkw_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.zkw_host_free.argtypes = [C.c_void_p, C.c_void_p]
    u8p = C.c_char_p
    u32p = C.POINTER(C.c_uint32)
    lib.zkw_ecdsa_circuit_new.argtypes = [C.POINTER(CircuitParamsC), C.POINTER(C.c_void_p)]
    lib.zkw_ecdsa_circuit_free.argtypes = [C.c_void_p]
    lib.zkw_ecdsa_circuit_free.restype = None
    lib.zkw_ecdsa_circuit_shape.argtypes = [C.c_void_p, C.POINTER(CircuitShape)]
    lib.zkw_ecdsa_circuit_rows.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
    lib.zkw_ecdsa_circuit_fixed.argtypes = [C.c_void_p, C.POINTER(u64p)]
    lib.zkw_ecdsa_circuit_permutation.argtypes = [C.c_void_p, C.POINTER(u32p)]
    lib.zkw_ecdsa_synthesize.argtypes = [C.c_void_p, u8p, u8p, u8p, u8p, u8p, C.POINTER(u64p), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    lib.zkw_g1_sum.argtypes = [u64p, C.c_size_t, u64p]
    lib.zkw_pk_write.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    lib.zkw_pk_read.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.zkw_vk_write.argtypes = [C.c_void_p, C.c_char_p]
    lib.zkw_vk_read.argtypes = [C.c_char_p, C.POINTER(CircuitShape), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), u64p, C.c_size_t, u64p, C.c_size_t, u64p]
    lib.zkw_prover_create.argtypes = [C.c_int, C.POINTER(CircuitParamsC), u64p, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.zkw_prover_destroy.argtypes = [C.c_void_p]
    lib.zkw_prover_destroy.restype = None
    lib.zkw_prover_ctx.argtypes = [C.c_void_p]
    lib.zkw_prover_ctx.restype = C.c_void_p
    lib.zkw_prover_pk.argtypes = [C.c_void_p]
    lib.zkw_prover_pk.restype = C.c_void_p
    lib.zkw_prover_last_synthesis_ms.argtypes = [C.c_void_p]
    lib.zkw_prover_last_synthesis_ms.restype = C.c_double
    lib.zkw_prover_prove.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_uint, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.zkw_prove_batch.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_uint, C.POINTER(C.c_uint8),
                                    C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    _lib = lib
    return lib


def witness_rows(shape: "CircuitShape") -> list[int]:
    """cells per advice column that zkw_synth_witness fills (gate columns, then lookup-advice columns)."""
    u = (1 << shape.k) - (shape.blinding_factors + 1)
    return [4 * (u // 4)] * shape.num_advice + [u] * shape.num_lookup_advice


def synth_witness(shape: "CircuitShape", lookup_bits: int, assertion: bytes, out: list[np.ndarray] | None = None) -> list[np.ndarray]:
    """zkw_synth_witness: the synthetic circuit's advice columns (canonical uint64 cells) for `assertion`,
    written into `out` (e.g. views of page-locked memory) or fresh arrays.  Host code, needs no GPU."""
    lib = load_library()
    rows = witness_rows(shape)
    if out is None:
        out = [np.empty(r, dtype=np.uint64) for r in rows]
    if len(out) != len(rows) or any(o.dtype != np.uint64 or o.shape != (r,) or not o.flags.c_contiguous for o, r in zip(out, rows)):
        raise ValueError("synth_witness: out must be C-contiguous uint64 arrays of witness_rows(shape) cells")
    ptrs = (u64p * len(out))(*[o.ctypes.data_as(u64p) for o in out])
    status = lib.zkw_synth_witness(C.byref(shape), C.c_uint32(lookup_bits), assertion, C.c_size_t(len(assertion)), ptrs, None)
    if status != ZKW_OK:
        raise ZkwError(status, "zkw_synth_witness")
    return out


def g1_sum(points_xyz: np.ndarray) -> np.ndarray:
    """zkw_g1_sum: (m, 12) Jacobian points -> their sum as (x, y, 1), or z = 0 for the identity.  Host code."""
    pts = _as_u64(points_xyz, 12)
    out = np.zeros(12, dtype=np.uint64)
    rc = load_library().zkw_g1_sum(_p(pts), C.c_size_t(pts.shape[0]), _p(out))
    if rc != ZKW_OK:
        raise ZkwError(rc, "zkw_g1_sum")
    return out


def _strerror(status: int) -> str:
    try:
        return load_library().zkw_strerror(status).decode()
    except Exception:  # pragma: no cover
        return "?"


def _as_u64(a: np.ndarray, cols: int | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def _addr(x) -> C.c_void_p:
    """device address from an int, a ctypes pointer, or anything with .data_ptr() (torch tensors)."""
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, int):
        return C.c_void_p(x)
    return C.cast(x, C.c_void_p)


class Context:
    """One zkw_ctx: bound to one CUDA device, one stream.  One Context per GPU / per process."""

    def __init__(self, device: int = 0, _borrowed=None):
        self.lib = load_library()
        self.device = device
        self._host_allocs: list = []
        self._owned = _borrowed is None
        if _borrowed is not None:          # a context owned by a zkw_prover
            self.h = C.c_void_p(_borrowed)
            return
        h = C.c_void_p()
        rc = self.lib.zkw_ctx_create(device, C.byref(h))
        if rc != ZKW_OK:
            raise ZkwError(rc, "zkw_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            for p in self._host_allocs:
                self.lib.zkw_host_free(self.h, p)
            self._host_allocs = []
            if self._owned:
                self.lib.zkw_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != ZKW_OK:
            detail = (self.lib.zkw_last_cuda_error(self.h) or b"").decode() if rc in (ZKW_ERR_CUDA, ZKW_ERR_OOM) else ""
            raise ZkwError(rc, what, detail)

    # -- plumbing ---------------------------------------------------------------------------
    @property
    def stream(self) -> int:
        return int(self.lib.zkw_ctx_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.zkw_ctx_launch_count(self.h))

    def sync(self):
        self._check(self.lib.zkw_ctx_sync(self.h), "zkw_ctx_sync")

    def host_array(self, count: int) -> np.ndarray:
        """(count,) uint64 array in page-locked host memory (freed with the context)."""
        p = C.c_void_p()
        self._check(self.lib.zkw_host_alloc(self.h, C.c_size_t(8 * max(count, 1)), C.byref(p)), "zkw_host_alloc")
        self._host_allocs.append(p)
        return np.ctypeslib.as_array(C.cast(p, u64p), shape=(max(count, 1),))[:count]

    def msm_window_bits(self, n: int) -> int:
        """window width c of an n-point MSM; the MSM makes ceil(255 / c) mixed additions per point"""
        c = self.lib.zkw_msm_window_bits(self.h, C.c_size_t(n))
        if c <= 0:
            raise ZkwError(c, "zkw_msm_window_bits")
        return int(c)

    def msm_config(self, window_bits: int = 0, precompute: bool = True):
        self._check(self.lib.zkw_msm_config(self.h, window_bits, int(precompute)), "zkw_msm_config")

    # -- per-kernel device timing ------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        self._check(self.lib.zkw_profile_enable(self.h, int(on)), "zkw_profile_enable")

    def selftest_field(self, seed: int = 1, count: int = 1 << 16) -> tuple[int, int]:
        """(Fr mismatches, Fq mismatches) of the dedicated squaring against the general product on the device"""
        out = (C.c_uint * 2)()
        self._check(self.lib.zkw_selftest_field(self.h, C.c_uint64(seed), C.c_uint(count), out), "zkw_selftest_field")
        return int(out[0]), int(out[1])

    def profile_filter(self, kernel: str | None):
        """time only launches of `kernel` (None: all kernels)"""
        self._check(self.lib.zkw_profile_filter(self.h, kernel.encode() if kernel else None), "zkw_profile_filter")

    def profile_reset(self):
        self._check(self.lib.zkw_profile_reset(self.h), "zkw_profile_reset")

    def profile_read(self, kernel: str) -> tuple[float, int]:
        ms, cnt = C.c_double(0), C.c_uint64(0)
        self._check(self.lib.zkw_profile_read(self.h, kernel.encode(), C.byref(ms), C.byref(cnt)), "zkw_profile_read")
        return ms.value, cnt.value

    def profile_all(self) -> dict:
        buf = C.create_string_buffer(4096)
        self._check(self.lib.zkw_profile_names(self.h, buf, C.c_size_t(4096)), "zkw_profile_names")
        names = [x for x in buf.value.decode().split(",") if x]
        return {n: self.profile_read(n) for n in names}

    # -- SRS ----------------------------------------------------------------------------------
    def srs_setup(self, k: int, tau: np.ndarray):
        """gen_srs(k) analogue with an explicit tau (Montgomery form, (4,) uint64)."""
        t = _as_u64(tau).reshape(4)
        self._check(self.lib.zkw_srs_setup(self.h, C.c_uint(k), _p(t)), "zkw_srs_setup")

    def srs_get(self, which: int, n: int) -> np.ndarray:
        out = np.zeros((n, 8), dtype=np.uint64)
        self._check(self.lib.zkw_srs_get(self.h, which, _p(out), C.c_size_t(n)), "zkw_srs_get")
        return out

    def fixed_base_mul(self, scalars: np.ndarray) -> np.ndarray:
        s = _as_u64(scalars, 4)
        out = np.zeros((s.shape[0], 8), dtype=np.uint64)
        self._check(self.lib.zkw_g1_fixed_base_mul(self.h, _p(s), C.c_size_t(s.shape[0]), _p(out)), "zkw_g1_fixed_base_mul")
        return out

    def srs_load(self, g: np.ndarray, g_lagrange: np.ndarray | None = None):
        g = _as_u64(g, 8)
        gl = _as_u64(g_lagrange, 8) if g_lagrange is not None else None
        if gl is not None and gl.shape != g.shape:
            raise ValueError("g and g_lagrange must have the same length")
        self._check(self.lib.zkw_srs_load(self.h, _p(g), _p(gl) if gl is not None else None, C.c_size_t(g.shape[0])), "zkw_srs_load")

    def srs_load_dev(self, g_dev, g_lagrange_dev, n: int):
        self._check(self.lib.zkw_srs_load_dev(self.h, _addr(g_dev), _addr(g_lagrange_dev) if g_lagrange_dev is not None else None,
                                              C.c_size_t(n)), "zkw_srs_load_dev")

    # -- MSM: best_multiexp ---------------------------------------------------------------------
    def msm(self, scalars: np.ndarray, bases: np.ndarray | None = None, which: int | None = None) -> np.ndarray:
        """sum_i scalars[i] * bases[i] -> Jacobian (12,) with Z = 1 (or Z = 0 for the identity)."""
        s = _as_u64(scalars, 4)
        if which is None:
            which = BASES_CALLER if bases is not None else BASES_G
        out = np.zeros(12, dtype=np.uint64)
        if which == BASES_CALLER:
            b = _as_u64(bases, 8)
            if b.shape[0] != s.shape[0]:
                raise ValueError("scalars / bases length mismatch")
            rc = self.lib.zkw_msm_bn254_g1(self.h, which, _p(b), _p(s), C.c_size_t(s.shape[0]), _p(out))
        else:
            rc = self.lib.zkw_msm_bn254_g1(self.h, which, None, _p(s), C.c_size_t(s.shape[0]), _p(out))
        self._check(rc, "zkw_msm_bn254_g1")
        return out

    def msm_dev(self, scalars_dev, n: int, which: int = BASES_G, bases_dev=None) -> np.ndarray:
        out = np.zeros(12, dtype=np.uint64)
        rc = self.lib.zkw_msm_bn254_g1_dev_to_host(self.h, which, _addr(bases_dev) if bases_dev is not None else None,
                                                   _addr(scalars_dev), C.c_size_t(n), _p(out))
        self._check(rc, "zkw_msm_bn254_g1_dev_to_host")
        return out

    def g1_batch_normalize(self, xyz: np.ndarray) -> np.ndarray:
        x = _as_u64(xyz, 12)
        out = np.zeros((x.shape[0], 8), dtype=np.uint64)
        self._check(self.lib.zkw_g1_batch_normalize(self.h, _p(x), C.c_size_t(x.shape[0]), _p(out)), "zkw_g1_batch_normalize")
        return out

    # -- NTT: best_fft and the EvaluationDomain transforms ------------------------------------------
    def ntt(self, a: np.ndarray, omega: np.ndarray, scale: np.ndarray | None = None) -> np.ndarray:
        a = np.array(_as_u64(a, 4), copy=True)
        n = a.shape[0]
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise ZkwError(ZKW_ERR_INVALID, "ntt", "length must be a power of two")
        om = _as_u64(omega).reshape(4)
        sc = _as_u64(scale).reshape(4) if scale is not None else None
        self._check(self.lib.zkw_ntt_bn254_fr(self.h, _p(a), C.c_uint(log_n), _p(om), _p(sc) if sc is not None else None), "zkw_ntt_bn254_fr")
        return a

    def ntt_dev(self, a_dev, log_n: int, omega: np.ndarray, scale: np.ndarray | None = None):
        om = _as_u64(omega).reshape(4)
        sc = _as_u64(scale).reshape(4) if scale is not None else None
        self._check(self.lib.zkw_ntt_bn254_fr_dev(self.h, _addr(a_dev), C.c_uint(log_n), _p(om), _p(sc) if sc is not None else None),
                    "zkw_ntt_bn254_fr_dev")

    def lagrange_to_coeff(self, a: np.ndarray, inplace: bool = False) -> np.ndarray:
        """inplace=True transforms the caller's (C-contiguous uint64) array itself, as the C ABI does - no copy, and page-locked
        caller memory (Context.host_array) keeps its transfer speed."""
        a = _as_u64(a, 4) if inplace else np.array(_as_u64(a, 4), copy=True)
        k = a.shape[0].bit_length() - 1
        self._check(self.lib.zkw_lagrange_to_coeff(self.h, _p(a), C.c_uint(k)), "zkw_lagrange_to_coeff")
        return a

    def coeff_to_lagrange(self, a: np.ndarray) -> np.ndarray:
        a = np.array(_as_u64(a, 4), copy=True)
        k = a.shape[0].bit_length() - 1
        self._check(self.lib.zkw_coeff_to_lagrange(self.h, _p(a), C.c_uint(k)), "zkw_coeff_to_lagrange")
        return a

    def coeff_to_extended(self, a: np.ndarray, ext_k: int, out: np.ndarray | None = None) -> np.ndarray:
        a = _as_u64(a, 4)
        k = a.shape[0].bit_length() - 1
        if out is None:
            out = np.zeros((1 << ext_k, 4), dtype=np.uint64)
        elif out.shape != (1 << ext_k, 4) or out.dtype != np.uint64 or not out.flags.c_contiguous:
            raise ValueError("coeff_to_extended: out must be a C-contiguous (2^ext_k, 4) uint64 array")
        self._check(self.lib.zkw_coeff_to_extended(self.h, _p(a), C.c_uint(k), C.c_uint(ext_k), _p(out)), "zkw_coeff_to_extended")
        return out

    def extended_to_coeff(self, a: np.ndarray, inplace: bool = False) -> np.ndarray:
        a = _as_u64(a, 4) if inplace else np.array(_as_u64(a, 4), copy=True)
        ek = a.shape[0].bit_length() - 1
        self._check(self.lib.zkw_extended_to_coeff(self.h, _p(a), C.c_uint(ek)), "zkw_extended_to_coeff")
        return a

    def lagrange_to_coeff_dev(self, a_dev, k: int):
        self._check(self.lib.zkw_lagrange_to_coeff_dev(self.h, _addr(a_dev), C.c_uint(k)), "zkw_lagrange_to_coeff_dev")

    def coeff_to_lagrange_dev(self, a_dev, k: int):
        self._check(self.lib.zkw_coeff_to_lagrange_dev(self.h, _addr(a_dev), C.c_uint(k)), "zkw_coeff_to_lagrange_dev")

    def coeff_to_extended_dev(self, coeffs_dev, k: int, ext_k: int, out_dev):
        self._check(self.lib.zkw_coeff_to_extended_dev(self.h, _addr(coeffs_dev), C.c_uint(k), C.c_uint(ext_k), _addr(out_dev)),
                    "zkw_coeff_to_extended_dev")

    def extended_to_coeff_dev(self, a_dev, ext_k: int):
        self._check(self.lib.zkw_extended_to_coeff_dev(self.h, _addr(a_dev), C.c_uint(ext_k)), "zkw_extended_to_coeff_dev")

    # -- quotient: evaluate_h + divide_by_vanishing_poly --------------------------------------------
    @staticmethod
    def _quotient_struct(shape: CircuitShape, cols: dict, challenges: dict, ptr_of):
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
        for name in ("advice", "constants", "q_enable", "sigma", "perm_z", "lookup_z", "lookup_a", "lookup_s"):
            setattr(q, name, table(name))
        for name in ("table", "q_lookup", "l0", "l_last", "l_active"):
            setattr(q, name, single(name))
        for name in ("y", "beta", "gamma", "theta"):
            v = np.asarray(challenges[name], dtype=np.uint64).reshape(4)
            getattr(q, name)[:] = [int(x) for x in v]
        return q, keep

    def quotient(self, shape: CircuitShape, cols: dict, challenges: dict, out: np.ndarray | None = None) -> np.ndarray:
        """cols: name -> (2^ext_k, 4) array, or list of them for advice / constants / q_enable / sigma /
        perm_z / lookup_z / lookup_a / lookup_s.  Returns h on the extended coset (written into `out` if given)."""
        cols = {k: ([_as_u64(x, 4) for x in v] if isinstance(v, (list, tuple)) else (_as_u64(v, 4) if v is not None else None))
                for k, v in cols.items()}
        q, keep = self._quotient_struct(shape, cols, challenges, _p)
        if out is None:
            out = np.zeros((1 << shape.ext_k, 4), dtype=np.uint64)
        elif out.shape != (1 << shape.ext_k, 4) or out.dtype != np.uint64 or not out.flags.c_contiguous:
            raise ValueError("quotient: out must be a C-contiguous (2^ext_k, 4) uint64 array")
        self._check(self.lib.zkw_quotient_ecdsa(self.h, C.byref(q), _p(out)), "zkw_quotient_ecdsa")
        del keep
        return out

    def quotient_dev(self, shape: CircuitShape, cols_dev: dict, challenges: dict, h_dev):
        q, keep = self._quotient_struct(shape, cols_dev, challenges, lambda a: C.cast(_addr(a), u64p))
        self._check(self.lib.zkw_quotient_ecdsa_dev(self.h, C.byref(q), _addr(h_dev)), "zkw_quotient_ecdsa_dev")
        del keep
