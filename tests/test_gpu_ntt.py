"""GPU parity: zkw_ntt_bn254_fr and the EvaluationDomain transforms vs the CPU oracle (bit-exact).

The oracle restates halo2_proofs::arithmetic::best_fft / poly::EvaluationDomain, which the reference
reaches through create_proof (halo2-circuits/src/ecc/ecdsa_p256.rs:366-373, 416-423, 555-562).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _omega(oracle, log_n):
    from oracle import pyref as pr
    w = pow(pr.FR_ROOT_OF_UNITY, 1 << (pr.FR_S - log_n), pr.R)
    return w, oracle.fr_to_mont([w])[0]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18])
def test_ntt_matches_oracle(ctx, oracle, log_n):
    n = 1 << log_n
    a = oracle.fr_random(n, 100 + log_n)
    _, om = _omega(oracle, log_n)
    got = ctx.ntt(a, om)
    want = oracle.best_fft(a, om)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("log_n,plan", [(16, "8,8"), (16, "10,6"), (16, "6,10"), (16, "3,3,10"), (16, "4,4,4,4"), (18, "9,9"), (18, "7,7,4"),
                                        (18, "1,7,10"), (13, "10,3"), (13, "3,10"), (12, "6,6")])
@pytest.mark.parametrize("staged", ["0", "1"])
def test_ntt_stage_splits_and_staged_twiddles(ctx, oracle, monkeypatch, log_n, plan, staged):
    """Every split of the log_n stages into passes gives the same transform, with and without the opt-in TMA path
    (ZKW_NTT_STAGE_TWIDDLES=1): passes after the first then take their last stage's twiddles from the per-(pass,
    tile group) table that one bulk copy stages into shared memory (any s0 >= C, including C = 0 where a tile is
    a single column)."""
    monkeypatch.setenv(f"ZKW_NTT_PLAN_{log_n}", plan)
    monkeypatch.setenv("ZKW_NTT_STAGE_TWIDDLES", staged)
    n = 1 << log_n
    a = oracle.fr_random(n, 300 + log_n)
    _, om = _omega(oracle, log_n)
    assert np.array_equal(ctx.ntt(a, om), oracle.best_fft(a, om))


@pytest.mark.parametrize("log_n", [1, 5, 10, 12, 17])
def test_inverse_ntt_with_scale(ctx, oracle, log_n):
    from oracle import pyref as pr
    n = 1 << log_n
    a = oracle.fr_random(n, 200 + log_n)
    w, om = _omega(oracle, log_n)
    om_inv = oracle.fr_to_mont([pow(w, -1, pr.R)])[0]
    n_inv = oracle.fr_to_mont([pow(n, -1, pr.R)])[0]
    fwd = ctx.ntt(a, om)
    back = ctx.ntt(fwd, om_inv, scale=n_inv)
    assert np.array_equal(back, a)


def test_ntt_small_case_against_python_definition(ctx, oracle):
    """a[i] = sum_j a_j w^(ij) straight from the definition (pure Python), n = 16."""
    from oracle import pyref as pr
    a = oracle.fr_random(16, 7)
    w, om = _omega(oracle, 4)
    got = oracle.fr_from_mont(ctx.ntt(a, om))
    assert got == pr.dft_naive(oracle.fr_from_mont(a), w)


def test_ntt_edge_vectors(ctx, oracle):
    from oracle import pyref as pr
    log_n = 11
    n = 1 << log_n
    _, om = _omega(oracle, log_n)
    zero = np.zeros((n, 4), dtype=np.uint64)
    assert np.array_equal(ctx.ntt(zero, om), zero)
    # delta at 0 -> all ones ; all (r-1) -> oracle
    d = zero.copy()
    d[0] = oracle.fr_to_mont([1])[0]
    ones = np.tile(oracle.fr_to_mont([1])[0], (n, 1))
    assert np.array_equal(ctx.ntt(d, om), ones)
    top = np.tile(oracle.fr_to_mont([pr.R - 1])[0], (n, 1))
    assert np.array_equal(ctx.ntt(top, om), oracle.best_fft(top, om))


def test_ntt_rejects_bad_length(ctx, zkw, oracle):
    _, om = _omega(oracle, 3)
    with pytest.raises(zkw.ZkwError):
        ctx.ntt(np.zeros((6, 4), dtype=np.uint64), om)


@pytest.mark.parametrize("deg,k", [(4, 1), (4, 3), (5, 4), (4, 8), (5, 9), (4, 10), (5, 11), (4, 13), (5, 15), (4, 17)])
def test_domain_transforms_match_oracle(ctx, oracle, deg, k):
    d = oracle.Domain.new(deg, k)
    a = oracle.fr_random(1 << k, 300 + k)
    coeff = ctx.lagrange_to_coeff(a)
    assert np.array_equal(coeff, d.lagrange_to_coeff(a))
    assert np.array_equal(ctx.coeff_to_lagrange(coeff), a)
    ext = ctx.coeff_to_extended(coeff, d.ext_k)
    assert np.array_equal(ext, d.coeff_to_extended(coeff))
    back = ctx.extended_to_coeff(ext)
    assert np.array_equal(back, d.extended_to_coeff(ext))
    assert np.array_equal(back[: 1 << k], coeff)
    assert not back[1 << k:].any()


def test_full_size_round_trip_k19(ctx, oracle):
    """BASELINE size (k = 19, extended 2^21): size-independent properties instead of an oracle run —
    iNTT(NTT(a)) = a, extended_to_coeff(coeff_to_extended(c)) = c zero-padded, and linearity."""
    k, ek = 19, 21
    a = oracle.fr_random(1 << k, 4242)
    coeff = ctx.lagrange_to_coeff(a)
    assert np.array_equal(ctx.coeff_to_lagrange(coeff), a)
    ext = ctx.coeff_to_extended(coeff, ek)
    back = ctx.extended_to_coeff(ext)
    assert np.array_equal(back[: 1 << k], coeff)
    assert not back[1 << k:].any()
    # spot-check 4 extended evaluations against Horner on the CPU (python ints)
    from oracle import pyref as pr
    dom = pr.EvaluationDomain(5, k)
    cv = oracle.fr_from_mont(coeff)
    for i in (0, 1, 12345, (1 << ek) - 1):
        x = dom.g_coset * pow(dom.extended_omega, i, pr.R) % pr.R
        assert oracle.fr_from_mont(ext[i:i + 1])[0] == pr.poly_eval(cv, x)


def test_full_size_transforms_k19_match_oracle_bit_for_bit(ctx, oracle):
    """BASELINE size, every element compared with the oracle: lagrange_to_coeff at 2^19, coeff_to_extended
    2^19 -> 2^21, extended_to_coeff at 2^21 (on an arbitrary vector, not only on an image of the extension),
    and a plain best_fft at 2^21."""
    k, ek = 19, 21
    d = oracle.Domain.new(5, k)
    assert d.ext_k == ek
    a = oracle.fr_random(1 << k, 777)
    coeff = ctx.lagrange_to_coeff(a)
    assert np.array_equal(coeff, d.lagrange_to_coeff(a))
    ext = ctx.coeff_to_extended(coeff, ek)
    assert np.array_equal(ext, d.coeff_to_extended(coeff))
    e = oracle.fr_random(1 << ek, 778)
    assert np.array_equal(ctx.extended_to_coeff(e), d.extended_to_coeff(e))
    _, om = _omega(oracle, ek)
    assert np.array_equal(ctx.ntt(e, om), oracle.best_fft(e, om))


def test_full_size_transforms_k17_match_oracle_bit_for_bit(ctx, oracle):
    """The server's degree (proving-server/src/main.rs:17): 2^17 rows, extended 2^19 (constraint degree 4)."""
    k = 17
    d = oracle.Domain.new(4, k)
    a = oracle.fr_random(1 << k, 779)
    coeff = ctx.lagrange_to_coeff(a)
    assert np.array_equal(coeff, d.lagrange_to_coeff(a))
    ext = ctx.coeff_to_extended(coeff, d.ext_k)
    assert np.array_equal(ext, d.coeff_to_extended(coeff))
    e = oracle.fr_random(1 << d.ext_k, 780)
    assert np.array_equal(ctx.extended_to_coeff(e), d.extended_to_coeff(e))
