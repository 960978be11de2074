"""Host-side mirror of the reference's prover API (halo2-circuits/src/ecc/ecdsa_p256.rs) over the device
prover of libzkw_b200.so:

    download_keys(degree, pk_path, vk_path)                       ecdsa_p256.rs:256-272
    generate_proof(pubkey_x, pubkey_y, r, s, msg_hash, pk_path, degree)      :379-427  (Blake2b transcript)
    generate_proof_evm(...same...)                                           :329-377  (EVM transcript)

Same argument meaning (five 32-byte little-endian canonical encodings + a proving-key path + degree) and
error behaviour (ValueError where the reference `.unwrap()`s a failed from_bytes / from_xy).  Unlike the
reference, which re-reads the SRS and proving key from disk on every request (ecdsa_p256.rs:338-343,
388-393), keys stay resident on the GPU in a per-process cache keyed by (degree, proving_key_path).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import native
from .circuit import CircuitParams, EcdsaCircuit, InvalidSignature, SyntheticEcdsaCircuit, to_limbs, validate_assertion  # noqa: F401

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)

TRANSCRIPT_BLAKE2B, TRANSCRIPT_EVM = 0, 1

# development tau of the resident SRS (upstream's gen_srs draws it from a seeded RNG; any fixed value
# gives a reproducible dev SRS — NOT a production ceremony)
DEV_TAU_CANONICAL = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class ProvingKey:
    def __init__(self, ctx: native.Context, handle, shape: native.CircuitShape, circuit, owned: bool = True):
        self.ctx, self.h, self.shape, self.circuit, self._owned = ctx, handle, shape, circuit, owned

    def write(self, path: str):
        """ProvingKey::to_bytes(RawBytes) to a file (ecdsa_p256.rs:261-265)."""
        _bind(self.ctx.lib)
        self.ctx._check(self.ctx.lib.zkw_pk_write(self.ctx.h, self.h, os.fsencode(path)), "zkw_pk_write")

    def write_vk(self, path: str):
        """VerifyingKey::to_bytes(RawBytes) to a file (ecdsa_p256.rs:266-270)."""
        _bind(self.ctx.lib)
        self.ctx._check(self.ctx.lib.zkw_vk_write(self.h, os.fsencode(path)), "zkw_vk_write")

    @classmethod
    def read(cls, ctx: native.Context, path: str, circuit=None) -> "ProvingKey":
        """ProvingKey::read(RawBytes) (ecdsa_p256.rs:339-343); the context must hold the SRS of the key's size."""
        _bind(ctx.lib)
        h = C.c_void_p()
        ctx._check(ctx.lib.zkw_pk_read(ctx.h, os.fsencode(path), C.byref(h)), "zkw_pk_read")
        shape, _, _, _, _ = read_vk(path)
        return cls(ctx, h, shape, circuit)

    def vk(self):
        """(fixed commitments (nfixed, 8), permutation commitments (nperm, 8), digest (4,)) — Montgomery."""
        lib = self.ctx.lib
        nf, npm = C.c_uint32(0), C.c_uint32(0)
        self.ctx._check(lib.zkw_pk_info(self.h, C.byref(nf), C.byref(npm)), "zkw_pk_info")
        fx = np.zeros((nf.value, 8), dtype=np.uint64)
        pm = np.zeros((npm.value, 8), dtype=np.uint64)
        dg = np.zeros(4, dtype=np.uint64)
        self.ctx._check(lib.zkw_pk_vk(self.h, fx.ctypes.data_as(u64p), pm.ctypes.data_as(u64p), dg.ctypes.data_as(u64p)), "zkw_pk_vk")
        return fx, pm, dg

    def close(self):
        if self.h:
            if self._owned:
                self.ctx.lib.zkw_pk_destroy(self.ctx.h, self.h)
            self.h = None


def read_vk(path: str):
    """VerifyingKey::read(RawBytes) (ecdsa_p256.rs:280-284), from a verifying-key or a proving-key file:
    (shape, fixed commitments (nfixed, 8), permutation commitments (nperm, 8), digest (4,)) — Montgomery — and the raw counts."""
    lib = native.load_library()
    shape = native.CircuitShape()
    nf, npm = C.c_uint32(0), C.c_uint32(0)
    dg = np.zeros(4, dtype=np.uint64)
    rc = lib.zkw_vk_read(os.fsencode(path), C.byref(shape), C.byref(nf), C.byref(npm), None, 0, None, 0, dg.ctypes.data_as(u64p))
    if rc != native.ZKW_OK:
        raise native.ZkwError(rc, "zkw_vk_read")
    fx = np.zeros((nf.value, 8), dtype=np.uint64)
    pm = np.zeros((npm.value, 8), dtype=np.uint64)
    rc = lib.zkw_vk_read(os.fsencode(path), None, None, None, fx.ctypes.data_as(u64p), nf.value, pm.ctypes.data_as(u64p), npm.value, None)
    if rc != native.ZKW_OK:
        raise native.ZkwError(rc, "zkw_vk_read")
    return shape, fx, pm, dg, (nf.value, npm.value)


def _bind(lib):
    lib.zkw_keygen.argtypes = [C.c_void_p, C.POINTER(native.CircuitShape), C.POINTER(u64p), C.POINTER(u32p), C.POINTER(C.c_void_p)]
    lib.zkw_pk_destroy.argtypes = [C.c_void_p, C.c_void_p]
    lib.zkw_pk_vk.restype = C.c_int
    lib.zkw_pk_destroy.restype = None
    lib.zkw_pk_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.zkw_pk_vk.argtypes = [C.c_void_p, u64p, u64p, u64p]
    lib.zkw_create_proof.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(u64p), C.POINTER(C.c_size_t), C.c_uint64, C.c_int,
                                     C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.zkw_create_proof_ex.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(u64p), C.POINTER(C.c_size_t), C.c_uint64, C.c_int, C.c_uint,
                                        C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.zkw_create_proof_seeded.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(u64p), C.POINTER(C.c_size_t), C.c_char_p, C.c_int, C.c_uint,
                                            C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.zkw_fr_to_mont.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
    lib.zkw_fr_from_mont.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]


def fr_to_mont(ctx: native.Context, canonical: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(canonical, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    _bind(ctx.lib)
    ctx._check(ctx.lib.zkw_fr_to_mont(ctx.h, a.ctypes.data_as(u64p), out.ctypes.data_as(u64p), C.c_size_t(a.shape[0])), "zkw_fr_to_mont")
    return out


def fr_from_mont(ctx: native.Context, mont: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(mont, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    _bind(ctx.lib)
    ctx._check(ctx.lib.zkw_fr_from_mont(ctx.h, a.ctypes.data_as(u64p), out.ctypes.data_as(u64p), C.c_size_t(a.shape[0])), "zkw_fr_from_mont")
    return out


def keygen(ctx: native.Context, shape: native.CircuitShape, fixed_values: list[np.ndarray], mapping: list[np.ndarray], circuit=None) -> ProvingKey:
    """keygen_vk + keygen_pk on the device.  fixed_values: Montgomery (n, 4) arrays; mapping: (n, 2) uint32."""
    _bind(ctx.lib)
    fv = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in fixed_values]
    mp = [np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, 2) for a in mapping]
    ft = (u64p * len(fv))(*[a.ctypes.data_as(u64p) for a in fv])
    mt = (u32p * len(mp))(*[a.ctypes.data_as(u32p) for a in mp])
    h = C.c_void_p()
    ctx._check(ctx.lib.zkw_keygen(ctx.h, C.byref(shape), ft, mt, C.byref(h)), "zkw_keygen")
    return ProvingKey(ctx, h, shape, circuit)


ADVICE_ON_DEVICE, ADVICE_CANONICAL, MULTIOPEN_SHPLONK, ADVICE_U64 = 1, 2, 4, 8
_PROOF_CAP = 1 << 20


def create_proof(ctx: native.Context, pk: ProvingKey, advice: list, seed, transcript: int, *, canonical: bool = False,
                 device_rows: list[int] | None = None, shplonk: bool = False, u64: bool = False) -> bytes:
    """create_proof on the device.  advice: one (rows, 4) uint64 array per advice column — host numpy arrays,
    or (with device_rows given) device tensors / addresses.  canonical=True: values are plain integers that
    the device converts to Montgomery form; u64=True: one uint64 per row ((rows,) arrays), widened on the device.  shplonk=True: SHPLONK multi-open (the reference's generate_proof)
    instead of GWC (generate_proof_evm).  seed: 32 bytes (from the OS: the blinding is then zero-knowledge) or an int
    below 2^64 (a reproducible stream for tests)."""
    _bind(ctx.lib)
    flags = (ADVICE_CANONICAL if canonical else 0) | (MULTIOPEN_SHPLONK if shplonk else 0) | (ADVICE_U64 if u64 else 0)
    if device_rows is not None:
        flags |= ADVICE_ON_DEVICE
        at = (u64p * len(advice))(*[C.cast(native._addr(a), u64p) for a in advice])
        rows = (C.c_size_t * len(advice))(*device_rows)
        keep = advice
    else:
        keep = [np.ascontiguousarray(a, dtype=np.uint64).reshape((-1,) if u64 else (-1, 4)) for a in advice]
        at = (u64p * len(keep))(*[a.ctypes.data_as(u64p) for a in keep])
        rows = (C.c_size_t * len(keep))(*[a.shape[0] for a in keep])
    buf = (C.c_uint8 * _PROOF_CAP)()
    n = C.c_size_t(0)
    if isinstance(seed, (bytes, bytearray)):
        if len(seed) != 32:
            raise ValueError("seed: expected 32 bytes")
        ctx._check(ctx.lib.zkw_create_proof_seeded(ctx.h, pk.h, at, rows, bytes(seed), transcript, flags, buf, _PROOF_CAP, C.byref(n)),
                   "zkw_create_proof_seeded")
    else:
        ctx._check(ctx.lib.zkw_create_proof_ex(ctx.h, pk.h, at, rows, C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), transcript, flags, buf, _PROOF_CAP,
                                               C.byref(n)), "zkw_create_proof_ex")
    del keep
    return bytes(buf[: n.value])


# ---- reference-facing API ------------------------------------------------------------------------------
class ProverState:
    """Device-resident SRS + proving key for one (degree, config): what the reference keeps on disk as
    ./params/kzg_bn254_<k>.srs and ./keys/proving_key.pk.

    circuit: the real P-256 ECDSA verification circuit (EcdsaCircuit, csrc/ecdsa_circuit.cpp) by default;
    synthetic=True selects the shape-identical PRNG-filled test system (SyntheticEcdsaCircuit) — a test shape for
    sizes the ECDSA circuit cannot fit and for the bit-exact comparison with the Python oracle prover; it attests
    nothing about a signature and is never reached through generate_proof / generate_proof_evm.

    A state owns one zkw_ctx (stream set, scratch arena, staging columns): prove() is serialised by a lock, so
    concurrent callers (the reference proves from Rocket worker threads, proving-server/src/main.rs:49-79) queue up
    on one state; use ProverPool for several provers per GPU."""

    def __init__(self, params: CircuitParams, device: int = 0, ctx: native.Context | None = None, synthetic: bool = False,
                 proving_key_path: str | None = None, verifying_key_path: str | None = None):
        import threading
        self.params = params
        self.synthetic = synthetic
        self._lock = threading.Lock()
        self._staging = None
        self.last_synth_ms = 0.0
        self._prover = None
        tau_mont = DEV_TAU_CANONICAL * (1 << 256) % FR_MODULUS
        tau = np.array([(tau_mont >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        if not synthetic and ctx is None:
            # the native resident prover (csrc/service.cpp): context + SRS + circuit + key (read from proving_key_path when
            # that file exists, else generated and written to the given paths) + page-locked staging, all behind the C ABI
            lib = native.load_library()
            _bind(lib)
            cp = native.CircuitParamsC(params.degree, params.num_advice, params.num_lookup_advice, params.num_fixed, params.lookup_bits,
                                       params.limb_bits, params.num_limbs)
            h = C.c_void_p()
            rc = lib.zkw_prover_create(device, C.byref(cp), tau.ctypes.data_as(u64p), os.fsencode(proving_key_path) if proving_key_path else None,
                                       os.fsencode(verifying_key_path) if verifying_key_path else None, C.byref(h))
            if rc != native.ZKW_OK:
                raise native.ZkwError(rc, "zkw_prover_create")
            self._prover = h
            self.ctx = native.Context(device, _borrowed=lib.zkw_prover_ctx(h))
            self.circuit = EcdsaCircuit(params)
            self.shape = self.circuit.shape
            self.pk = ProvingKey(self.ctx, C.c_void_p(lib.zkw_prover_pk(h)), self.shape, self.circuit, owned=False)
            return
        self.ctx = ctx or native.Context(device)
        self.ctx.srs_setup(params.degree, tau)                       # gen_srs(degree)
        if synthetic:
            self.circuit = SyntheticEcdsaCircuit(params)
            self.shape = native.CircuitShape.from_config(params.degree, params.num_advice, params.num_lookup_advice, params.num_fixed)
            fixed = [fr_to_mont(self.ctx, to_limbs(c)) for c in self.circuit.fixed_columns()]
        else:
            self.circuit = EcdsaCircuit(params)
            self.shape = self.circuit.shape
            fixed = [fr_to_mont(self.ctx, c) for c in self.circuit.fixed_columns()]
        self.pk = keygen(self.ctx, self.shape, fixed, self.circuit.permutation_mapping(), self.circuit)   # keygen_vk + keygen_pk
        if proving_key_path:
            self.pk.write(proving_key_path)
        if verifying_key_path:
            self.pk.write_vk(verifying_key_path)

    # -- witness -------------------------------------------------------------------------------------
    def _stage(self):
        if self._staging is None:
            if self.synthetic:
                self._staging = [self.ctx.host_array(r) for r in native.witness_rows(self.shape)]
            else:
                self._staging = [self.ctx.host_array(4 * r).reshape(r, 4) for r in self.circuit.rows]
                for a in self._staging:
                    a[:] = 0
        return self._staging

    def synthesize(self, assertion: bytes, allow_invalid: bool = False) -> list[np.ndarray]:
        """Advice columns in Montgomery form for `assertion` (160 bytes: pubkey_x | pubkey_y | r | s | msg_hash)."""
        if self.synthetic:
            return [fr_to_mont(self.ctx, to_limbs(c)) for c in self.circuit.synthesize(assertion)]
        cols = self.circuit.synthesize(*_split_assertion(assertion), allow_invalid=allow_invalid)
        return [fr_to_mont(self.ctx, c) for c in cols]

    def prove(self, assertion: bytes, transcript: int, seed=None, shplonk: bool = False) -> bytes:
        """witness synthesis on the host (into page-locked staging columns owned by this state), one H2D copy of the
        canonical advice values, proof bytes back.  Raises InvalidSignature before any device work when the
        signature does not verify."""
        import time
        if self._prover is not None:
            return self._prove_native(assertion, transcript, seed, shplonk)
        if seed is None:
            seed = os.urandom(32)                                    # the reference draws blinding from OsRng (ecdsa_p256.rs:362)
        with self._lock:
            staging = self._stage()
            t0 = time.perf_counter()
            if self.synthetic:
                # canonical values < 2^64: shipped as one u64 per row
                cols = native.synth_witness(self.shape, self.params.lookup_bits, assertion, out=staging)
                self.last_synth_ms = 1e3 * (time.perf_counter() - t0)
                return create_proof(self.ctx, self.pk, cols, seed, transcript, shplonk=shplonk, u64=True)
            cols = self.circuit.synthesize(*_split_assertion(assertion), out=staging)
            self.last_synth_ms = 1e3 * (time.perf_counter() - t0)
            # create_proof returns after the device is done with the staging columns (the proof bytes depend on
            # them), so the next call may overwrite them
            return create_proof(self.ctx, self.pk, cols, seed, transcript, shplonk=shplonk, canonical=True)

    def _prove_native(self, assertion: bytes, transcript: int, seed, shplonk: bool) -> bytes:
        if len(assertion) != 160:
            raise ValueError("assertion: expected pubkey_x | pubkey_y | r | s | msg_hash, 5 x 32 little-endian bytes")
        lib = self.ctx.lib
        buf = (C.c_uint8 * _PROOF_CAP)()
        n = C.c_size_t(0)
        rc = lib.zkw_prover_prove(self._prover, bytes(assertion), _seed_bytes(seed), transcript, MULTIOPEN_SHPLONK if shplonk else 0, buf, _PROOF_CAP,
                                  C.byref(n))
        self.last_synth_ms = float(lib.zkw_prover_last_synthesis_ms(self._prover))
        if rc == native.ZKW_ERR_SIGNATURE:
            raise InvalidSignature("signature does not verify: the ECDSA circuit has no satisfying assignment")
        if rc == native.ZKW_ERR_INVALID:
            raise ValueError("non-canonical field element in the assertion (or bad transcript / output buffer)")
        self.ctx._check(rc, "zkw_prover_prove")
        return bytes(buf[: n.value])

    def close(self):
        if self._prover is not None:
            self.pk.close()
            self.circuit.close()
            self.ctx.h = None                     # owned by the native prover
            self.ctx.lib.zkw_prover_destroy(self._prover)
            self._prover = None
            return
        self.pk.close()
        if not self.synthetic:
            self.circuit.close()


def _seed_bytes(seed):
    """None -> None (the library draws from the OS); bytes -> as is; int below 2^64 -> the zero-padded key the 64-bit entry
    points use (reproducible streams for tests)."""
    if seed is None:
        return None
    if isinstance(seed, (bytes, bytearray)):
        if len(seed) != 32:
            raise ValueError("seed: expected 32 bytes")
        return bytes(seed)
    return (seed & 0xFFFFFFFFFFFFFFFF).to_bytes(8, "little") + bytes(24)


def _split_assertion(assertion: bytes):
    if len(assertion) != 160:
        raise ValueError("assertion: expected pubkey_x | pubkey_y | r | s | msg_hash, 5 x 32 little-endian bytes")
    return tuple(assertion[32 * i: 32 * i + 32] for i in range(5))


class ProverPool:
    """Several ProverStates on ONE GPU proving independent assertions concurrently (the reference's batch
    use: Rocket's worker threads each proving one request, proving-server/src/main.rs:49-79).  Each worker
    owns a context (stream set, SRS tables, proving key, scratch arena); witness synthesis for the next
    assertion overlaps the device work of the others, and the latency-bound phases of one proof overlap the
    throughput-bound phases of another."""

    def __init__(self, params: CircuitParams, device: int | list[int] = 0, workers: int = 3, synthetic: bool = False):
        """device: one GPU, or a list of GPUs - `workers` provers on EACH of them, all driven from this process by the
        library's threads (zkw_prove_batch takes provers on any mix of devices); the one-process-per-GPU layout of bench.py
        is the alternative."""
        devices = [device] if isinstance(device, int) else list(device)
        self.states = [ProverState(params, d, synthetic=synthetic) for d in devices for _ in range(workers)]

    def prove_many(self, assertions: list[bytes], transcript: int, seed0: int | None = None) -> list[bytes]:
        """seed0 = None (default): every proof draws its own blinding seed from the OS (the reference's OsRng,
        ecdsa_p256.rs:362); an integer gives the deterministic stream seed0 + i (tests only)."""
        if all(st._prover is not None for st in self.states):
            return self._prove_many_native(assertions, transcript, seed0)
        import queue
        import threading
        todo: "queue.Queue[int]" = queue.Queue()
        for i in range(len(assertions)):
            todo.put(i)
        out: list = [None] * len(assertions)
        errors: list = []

        def work(st: ProverState):
            while True:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                try:
                    out[i] = st.prove(assertions[i], transcript, seed=None if seed0 is None else seed0 + i)
                except Exception as e:  # noqa: BLE001 - surfaced below
                    errors.append(e)
                    return

        threads = [threading.Thread(target=work, args=(st,)) for st in self.states]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out

    def _prove_many_native(self, assertions, transcript, seed0):
        """zkw_prove_batch: one host thread per prover inside the library, no Python threads."""
        lib = self.states[0].ctx.lib
        count = len(assertions)
        if any(len(a) != 160 for a in assertions):
            raise ValueError("assertion: expected 160 bytes")
        handles = (C.c_void_p * len(self.states))(*[st._prover for st in self.states])
        blob = b"".join(bytes(a) for a in assertions)
        seeds = None if seed0 is None else b"".join(_seed_bytes(seed0 + i) for i in range(count))
        stride = 1 << 14
        while True:
            out = (C.c_uint8 * (stride * max(count, 1)))()
            lens = (C.c_size_t * max(count, 1))()
            status = (C.c_int * max(count, 1))()
            rc = lib.zkw_prove_batch(handles, len(self.states), blob, count, seeds, transcript, 0, out, stride, lens, status)
            if rc == native.ZKW_ERR_INVALID and stride < _PROOF_CAP and any(status[i] == native.ZKW_ERR_INVALID and lens[i] == 0 for i in range(count)):
                # either a malformed assertion or proofs longer than the slot (wide configs): retry once with the largest slot
                stride = _PROOF_CAP
                continue
            break
        for i in range(count):
            if status[i] == native.ZKW_ERR_SIGNATURE:
                raise InvalidSignature(f"assertion {i}: signature does not verify")
            if status[i] == native.ZKW_ERR_INVALID:
                raise ValueError(f"assertion {i}: non-canonical field element")
            self.states[0].ctx._check(status[i], f"zkw_prove_batch[{i}]")
        raw = bytes(out)
        return [raw[stride * i: stride * i + lens[i]] for i in range(count)]

    def close(self):
        for st in self.states:
            st.close()


_STATES: dict = {}


def _config_for(degree: int) -> CircuitParams:
    """ECDSA_CONFIG env var (a path to the one-line JSON) or the reference's table for `degree`
    (ecdsa_p256.rs:95-100 reads ./src/configs/ecdsa_circuit.config, which is the degree-17 line)."""
    path = os.environ.get("ECDSA_CONFIG")
    if path:
        with open(path) as f:
            p = CircuitParams.from_json(f.read())
        if p.degree != degree:
            raise ValueError(f"ECDSA_CONFIG is for degree {p.degree}, asked for {degree}")
        return p
    return CircuitParams.for_degree(degree)


def download_keys(degree: int, proving_key_path: str | None = None, verifying_key_path: str | None = None, device: int = 0) -> ProverState:
    """gen_srs + keygen_vk + keygen_pk, then the two key files (ecdsa_p256.rs:256-272) — and the keys stay resident on the
    GPU for the generate_proof* calls that name the same proving_key_path.  Like the reference it always regenerates:
    existing files at the given paths are replaced."""
    key = (degree, proving_key_path, device)
    old = _STATES.pop(key, None)
    if old is not None:
        old.close()
    for p in (proving_key_path, verifying_key_path):
        if p:
            os.makedirs(os.path.dirname(os.path.abspath(p)), exist_ok=True)
            if os.path.exists(p):
                os.remove(p)
    _STATES[key] = ProverState(_config_for(degree), device, proving_key_path=proving_key_path, verifying_key_path=verifying_key_path)
    return _STATES[key]


def _state_for(degree: int, proving_key_path: str | None, device: int) -> ProverState:
    """The resident prover for (degree, proving_key_path): the reference re-reads SRS and key per request
    (ecdsa_p256.rs:338-343, 388-393); here the first request reads the key file (ProvingKey::read) and later ones reuse
    the resident key.  A missing file is not an error as in the reference (`expect("Unable to open proving key
    file")`): the key is then generated in memory."""
    key = (degree, proving_key_path, device)
    if key not in _STATES:
        path = proving_key_path if proving_key_path and os.path.exists(proving_key_path) else None
        _STATES[key] = ProverState(_config_for(degree), device, proving_key_path=path)
    return _STATES[key]


def _assertion_bytes(pubkey_x, pubkey_y, r, s, msg_hash) -> bytes:
    validate_assertion(pubkey_x, pubkey_y, r, s, msg_hash)
    return bytes(pubkey_x) + bytes(pubkey_y) + bytes(r) + bytes(s) + bytes(msg_hash)


def generate_proof(pubkey_x: bytes, pubkey_y: bytes, r: bytes, s: bytes, msg_hash: bytes, proving_key_path: str, degree: int,
                   device: int = 0, seed=None) -> bytes:
    """ecdsa_p256.rs:379-427 — Blake2b transcript, SHPLONK multi-open."""
    a = _assertion_bytes(pubkey_x, pubkey_y, r, s, msg_hash)
    return _state_for(degree, proving_key_path, device).prove(a, TRANSCRIPT_BLAKE2B, seed, shplonk=True)


def generate_proof_evm(pubkey_x: bytes, pubkey_y: bytes, r: bytes, s: bytes, msg_hash: bytes, proving_key_path: str, degree: int,
                       device: int = 0, seed=None) -> bytes:
    """ecdsa_p256.rs:329-377 — EVM (keccak) transcript, GWC multi-open."""
    a = _assertion_bytes(pubkey_x, pubkey_y, r, s, msg_hash)
    return _state_for(degree, proving_key_path, device).prove(a, TRANSCRIPT_EVM, seed)
