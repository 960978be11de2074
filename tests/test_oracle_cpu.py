"""CPU suite: the C oracle (oracle/*.c) against the independent Python big-integer restatement
(oracle/pyref.py) on small cases, plus algebraic properties at larger sizes.  Runs without a GPU."""
import numpy as np
import pytest

from oracle import cpu, pyref as pr

R, P = pr.R, pr.P


def test_field_ops_match_python():
    a = cpu.fr_random(64, 1)
    av = cpu.fr_from_mont(a)
    rng = pr.SplitMix64(1)
    assert av == [rng.field(R) for _ in range(64)]
    for i in range(0, 64, 2):
        assert cpu.fr_from_mont(cpu.fr_mul(a[i], a[i + 1]).reshape(1, 4))[0] == av[i] * av[i + 1] % R
    # Montgomery form is value * 2^256 mod r, little-endian u64 limbs (halo2curves layout)
    assert cpu.limbs_to_int(a[0]) == pr.to_mont(av[0], R)
    assert cpu.limbs_to_int(cpu.fq_to_mont_one(5)) == pr.to_mont(5, P)


def test_g1_generator_and_fixed_base():
    g = cpu.g1_generator()
    assert cpu.g1_affine_to_ints(g) == pr.G1_GEN and cpu.g1_is_on_curve(g)
    sc = cpu.fr_random(5, 2)
    pts = cpu.g1_fixed_base_mul(sc)
    for s, p in zip(cpu.fr_from_mont(sc), pts):
        assert cpu.g1_affine_to_ints(p) == pr.g1_mul(pr.G1_GEN, s)
    # zero scalar -> identity encoded (0,0)
    z = cpu.g1_fixed_base_mul(np.zeros((1, 4), dtype=np.uint64))
    assert not z.any()


def test_srs_powers():
    tau = cpu.fr_random(1, 3)[0]
    t = cpu.fr_from_mont(tau.reshape(1, 4))[0]
    g = cpu.srs_powers(6, tau)
    for i in range(6):
        assert cpu.g1_affine_to_ints(g[i]) == pr.g1_mul(pr.G1_GEN, pow(t, i, R))


@pytest.mark.parametrize("n", [1, 3, 4, 31, 32, 70])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_best_multiexp_matches_python(n, threads):
    s = cpu.fr_random(n, 10 + n)
    b = cpu.g1_fixed_base_mul(cpu.fr_random(n, 20 + n))
    sv = cpu.fr_from_mont(s)
    bv = [cpu.g1_affine_to_ints(x) for x in b]
    want = pr.msm_naive(sv, bv)
    got = cpu.g1_affine_to_ints(cpu.g1_to_affine(cpu.best_multiexp(s, b, threads))[0])
    assert got == want
    if threads == 3:
        assert pr.best_multiexp(sv, bv, threads) == want  # the python restatement of the windowed routine agrees too


def test_best_multiexp_edge_cases():
    n = 40
    b = cpu.g1_fixed_base_mul(cpu.fr_random(n, 5))
    zero = np.zeros((n, 4), dtype=np.uint64)
    assert not cpu.g1_to_affine(cpu.best_multiexp(zero, b))[0].any()
    s = np.tile(cpu.fr_to_mont([R - 1])[0], (n, 1))
    got = cpu.g1_affine_to_ints(cpu.g1_to_affine(cpu.best_multiexp(s, b))[0])
    assert got == pr.msm_naive([R - 1] * n, [cpu.g1_affine_to_ints(x) for x in b])
    # empty input
    e = cpu.best_multiexp(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64))
    assert not e[8:].any()
    # windowed == naive at a size python cannot reach
    n = 3000
    s = cpu.fr_random(n, 6)
    b = cpu.g1_fixed_base_mul(cpu.fr_random(n, 7))
    assert np.array_equal(cpu.g1_to_affine(cpu.best_multiexp(s, b, 4)), cpu.g1_to_affine(cpu.msm_naive(s, b)))


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8])
def test_best_fft_matches_definition(log_n):
    n = 1 << log_n
    a = cpu.fr_random(n, 30 + log_n)
    w = pow(pr.FR_ROOT_OF_UNITY, 1 << (28 - log_n), R)
    out = cpu.fr_from_mont(cpu.best_fft(a, cpu.fr_to_mont([w])[0]))
    av = cpu.fr_from_mont(a)
    assert out == pr.dft_naive(av, w) == pr.best_fft(av, w, log_n)


@pytest.mark.parametrize("deg,k", [(4, 3), (5, 4), (4, 7)])
def test_domain_matches_python(deg, k):
    d = cpu.Domain.new(deg, k)
    pd = pr.EvaluationDomain(deg, k)
    assert (d.k, d.ext_k, d.quotient_poly_degree) == (pd.k, pd.extended_k, pd.quotient_poly_degree)
    assert cpu.fr_from_mont(d.arr("omega").reshape(1, 4))[0] == pd.omega
    assert cpu.fr_from_mont(d.arr("ext_omega").reshape(1, 4))[0] == pd.extended_omega
    a = cpu.fr_random(1 << k, 40 + k)
    av = cpu.fr_from_mont(a)
    c = d.lagrange_to_coeff(a)
    assert cpu.fr_from_mont(c) == pd.lagrange_to_coeff(av)
    e = d.coeff_to_extended(c)
    cv = cpu.fr_from_mont(c)
    assert cpu.fr_from_mont(e) == pd.coeff_to_extended(cv)
    for i in (0, 1, (1 << d.ext_k) - 1):   # definition: p(zeta * w_ext^i)
        assert cpu.fr_from_mont(e[i:i + 1])[0] == pr.poly_eval(cv, pd.g_coset * pow(pd.extended_omega, i, R) % R)
    back = d.extended_to_coeff(e)
    assert np.array_equal(back[: 1 << k], c) and not back[1 << k:].any()
    assert cpu.fr_from_mont(d.divide_by_vanishing_poly(e)) == pd.divide_by_vanishing_poly(cpu.fr_from_mont(e))


def test_ntt_round_trip_large():
    k = 16
    d = cpu.Domain.new(4, k)
    a = cpu.fr_random(1 << k, 50)
    assert np.array_equal(d.coeff_to_lagrange(d.lagrange_to_coeff(a)), a)


def _random_quotient_inputs(shape, seed):
    en = 1 << shape.ext_k
    A, L, F = shape.num_advice, shape.num_lookup_advice, shape.num_fixed
    ncols = A + L + F
    chunk = shape.cs_degree - 2
    nsets = (ncols + chunk - 1) // chunk
    nlk = L or 1
    it = iter(range(seed * 1000, seed * 1000 + 1000))
    vec = lambda: cpu.fr_random(en, next(it))
    cols = {"advice": [vec() for _ in range(A + L)], "constants": [vec() for _ in range(F)], "table": vec(),
            "q_enable": [vec() for _ in range(A)], "q_lookup": vec() if L == 0 else None,
            "sigma": [vec() for _ in range(ncols)], "perm_z": [vec() for _ in range(nsets)],
            "lookup_z": [vec() for _ in range(nlk)], "lookup_a": [vec() for _ in range(nlk)],
            "lookup_s": [vec() for _ in range(nlk)], "l0": vec(), "l_last": vec(), "l_active": vec()}
    ch = {name: cpu.fr_random(1, next(it))[0] for name in ("y", "beta", "gamma", "theta")}
    return cols, ch


@pytest.mark.parametrize("k,A,L,F", [(3, 1, 0, 1), (3, 4, 1, 1), (2, 5, 2, 2)])
def test_quotient_matches_python(k, A, L, F):
    shape = cpu.make_shape(k, A, L, F)
    cols, ch = _random_quotient_inputs(shape, k + A)
    got = cpu.fr_from_mont(cpu.quotient_ecdsa(shape, cols, ch))
    conv = lambda v: cpu.fr_from_mont(v)
    pcols = {kk: ([conv(x) for x in v] if isinstance(v, list) else (conv(v) if v is not None else None)) for kk, v in cols.items()}
    pch = {kk: cpu.fr_from_mont(v.reshape(1, 4))[0] for kk, v in ch.items()}
    pshape = {f: getattr(shape, f) for f, _ in shape._fields_}
    assert got == pr.quotient_ecdsa(pshape, pcols, pch)


def test_quotient_of_a_satisfied_system_is_a_polynomial():
    """If every constraint vanishes on the 2^k domain, h = numerator / (X^n - 1) has degree < 3n:
    build the simplest satisfied system (all-zero witness, z = 1, selectors off) and check that
    extended_to_coeff(h) has no coefficients beyond n*(deg-1)."""
    k = 4
    shape = cpu.make_shape(k, 2, 1, 1)
    d = cpu.Domain.new(shape.cs_degree, k)
    n, en = 1 << k, 1 << shape.ext_k
    one = cpu.fr_to_mont([1])[0]
    zero_l = np.zeros((n, 4), dtype=np.uint64)
    ones_l = np.tile(one, (n, 1))
    ext = lambda lag: d.coeff_to_extended(d.lagrange_to_coeff(lag))
    rnd = lambda s: ext(cpu.fr_random(n, s))
    l0 = zero_l.copy(); l0[0] = one
    llast = zero_l.copy(); llast[n - 7] = one
    lact = ones_l.copy(); lact[n - 7:] = 0
    cols = {"advice": [ext(zero_l)] * 3, "constants": [ext(zero_l)], "table": ext(zero_l), "q_enable": [rnd(1), rnd(2)],
            "q_lookup": None, "sigma": [rnd(3), rnd(4), rnd(5), rnd(6)], "perm_z": [ext(ones_l), ext(ones_l)],
            "lookup_z": [ext(ones_l)], "lookup_a": [ext(zero_l)], "lookup_s": [ext(zero_l)],
            "l0": ext(l0), "l_last": ext(llast), "l_active": ext(lact)}
    # with a zero witness the permutation products are only equal if sigma is the identity permutation:
    # use sigma_c = delta^c * omega^row, the identity mapping of halo2's permutation argument
    sig = []
    pd = pr.EvaluationDomain(shape.cs_degree, k)
    for c in range(4):
        vals = [pow(pr.FR_DELTA, c, R) * pow(pd.omega, row, R) % R for row in range(n)]
        sig.append(ext(cpu.fr_to_mont(vals)))
    cols["sigma"] = sig
    ch = {name: cpu.fr_random(1, 70 + i)[0] for i, name in enumerate(("y", "beta", "gamma", "theta"))}
    h = cpu.quotient_ecdsa(shape, cols, ch)
    hc = d.extended_to_coeff(h)
    assert not hc[n * (shape.cs_degree - 1):].any()
