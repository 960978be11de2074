"""webauthn-halo2_b200 — B200-native (sm_100a) hot path of zkwebauthn/webauthn-halo2's Halo2 prover:
BN254 G1 MSM, BN254 Fr NTT and coset quotient evaluation behind the C ABI of include/zkw_b200.h.

The directory name carries a hyphen (it mirrors the reference's name), so import it with
importlib:  zkw = importlib.import_module("webauthn-halo2_b200").
"""
from .native import (  # noqa: F401
    BASES_CALLER, BASES_G, BASES_G_LAGRANGE, CircuitShape, Context, EXPORTS, LIB_PATH, QuotientInputs, ZkwError,
    load_library,
)
from .build import build  # noqa: F401
from .circuit import CircuitParams, EcdsaCircuit, InvalidSignature, SyntheticEcdsaCircuit, validate_assertion  # noqa: F401,E402
from .assertion import synthetic_assertion  # noqa: F401,E402
from .prover import (  # noqa: F401,E402
    TRANSCRIPT_BLAKE2B, TRANSCRIPT_EVM, ProverPool, ProverState, ProvingKey, create_proof, download_keys, fr_from_mont, fr_to_mont,
    generate_proof, generate_proof_evm, keygen,
)
