"""GPU parity: zkw_quotient_ecdsa vs the CPU oracle's evaluate_h restatement (bit-exact), on random
cosets for every column layout the reference's configs produce
(halo2-circuits/src/configs/bench_ecdsa.config: 1..291 gate columns, 1..53 lookup columns, 1..4
constant columns; constraint list from proving-server/P256Verifier.yul:406-547)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_inputs(oracle, shape, seed):
    en = 1 << shape.ext_k
    A, L, F = shape.num_advice, shape.num_lookup_advice, shape.num_fixed
    ncols = A + L + F
    chunk = shape.cs_degree - 2
    nsets = (ncols + chunk - 1) // chunk
    nlk = L or 1
    it = iter(range(seed * 1000, seed * 1000 + 10000))

    def vec():
        return oracle.fr_random(en, next(it))

    cols = {
        "advice": [vec() for _ in range(A + L)],
        "constants": [vec() for _ in range(F)],
        "table": vec(),
        "q_enable": [vec() for _ in range(A)],
        "q_lookup": vec() if L == 0 else None,
        "sigma": [vec() for _ in range(ncols)],
        "perm_z": [vec() for _ in range(nsets)],
        "lookup_z": [vec() for _ in range(nlk)],
        "lookup_a": [vec() for _ in range(nlk)],
        "lookup_s": [vec() for _ in range(nlk)],
        "l0": vec(), "l_last": vec(), "l_active": vec(),
    }
    ch = {name: oracle.fr_random(1, next(it))[0] for name in ("y", "beta", "gamma", "theta")}
    return cols, ch


@pytest.mark.parametrize("k,A,L,F", [(4, 1, 0, 1), (6, 1, 0, 1), (5, 2, 1, 1), (6, 4, 1, 1), (5, 8, 2, 1), (4, 17, 3, 1), (4, 5, 2, 2), (3, 7, 3, 4), (10, 1, 0, 1), (9, 4, 1, 1)])
def test_quotient_matches_oracle(ctx, zkw, oracle, k, A, L, F):
    oshape = oracle.make_shape(k, A, L, F)
    shape = zkw.CircuitShape(*[getattr(oshape, f) for f, _ in oshape._fields_])
    cols, ch = random_inputs(oracle, oshape, k * 100 + A)
    want = oracle.quotient_ecdsa(oshape, cols, ch)
    got = ctx.quotient(shape, cols, ch)
    assert np.array_equal(got, want)


def test_shape_from_config_matches_reference_configs(zkw):
    """k=19 line of bench_ecdsa.config -> selector mode, degree 5, ext 2^21, 1 permutation set;
    k=17 line (ecdsa_circuit.config) -> 4+1 advice, degree 4, ext 2^19, 3 sets (yul:429-518)."""
    s19 = zkw.CircuitShape.from_config(19, 1, 1, 1)
    assert (s19.ext_k, s19.cs_degree, s19.num_lookup_advice, s19.perm_sets, s19.lookups) == (21, 5, 0, 1, 1)
    s17 = zkw.CircuitShape.from_config(17, 4, 1, 1)
    assert (s17.ext_k, s17.cs_degree, s17.num_lookup_advice, s17.perm_sets, s17.lookups) == (19, 4, 1, 3, 1)


def test_quotient_rejects_bad_shapes(ctx, zkw, oracle):
    oshape = oracle.make_shape(4, 2, 1, 1)
    cols, ch = random_inputs(oracle, oshape, 9)
    bad = zkw.CircuitShape(4, 5, 2, 1, 1, 6, 4, 0)  # ext_k inconsistent with degree
    with pytest.raises(zkw.ZkwError):
        ctx.quotient(bad, cols, ch)
    sel = zkw.CircuitShape(4, 6, 2, 0, 1, 6, 5, 0)  # selector mode needs exactly one gate column
    with pytest.raises(zkw.ZkwError):
        ctx.quotient(sel, cols, ch)


@pytest.mark.parametrize("k,A,L,F", [(19, 1, 0, 1), (17, 4, 1, 1)])
def test_quotient_full_size_matches_oracle_bit_for_bit(ctx, zkw, oracle, k, A, L, F):
    """BASELINE sizes: k = 19 (bench_ecdsa.config:1; 2^21 extended rows, 14 input cosets) and k = 17
    (ecdsa_circuit.config:1; 2^19 extended rows), every row compared with the oracle's evaluate_h restatement."""
    oshape = oracle.make_shape(k, A, L, F)
    shape = zkw.CircuitShape(*[getattr(oshape, f) for f, _ in oshape._fields_])
    cols, ch = random_inputs(oracle, oshape, k)
    want = oracle.quotient_ecdsa(oshape, cols, ch)
    got = ctx.quotient(shape, cols, ch)
    assert np.array_equal(got, want)
