"""GPU parity: zkw_msm_bn254_g1 vs the CPU oracle's best_multiexp restatement (compared after affine
normalisation: the only canonical form of a projective result).

best_multiexp is reached by the reference through ParamsKZG::commit{,_lagrange} inside create_proof /
keygen (halo2-circuits/src/ecc/ecdsa_p256.rs:259-260, 366-373, 416-423, 555-562).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bases(oracle, n, seed):
    return oracle.g1_fixed_base_mul(oracle.fr_random(n, seed))


def _affine(oracle, xyz):
    return oracle.g1_to_affine(np.asarray(xyz).reshape(1, 12))[0]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 100, 1000, 4097, 1 << 13, (1 << 14) + 7, 1 << 16])
def test_msm_caller_bases_matches_oracle(ctx, oracle, n):
    s = oracle.fr_random(n, 1000 + n)
    b = _bases(oracle, n, 2000 + n)
    got = ctx.msm(s, b)
    want = oracle.best_multiexp(s, b)
    assert np.array_equal(_affine(oracle, got), _affine(oracle, want))
    # the ABI returns the normalised representative (x, y, 1)
    assert np.array_equal(got[:8], _affine(oracle, want))


def test_msm_small_against_python_naive(ctx, oracle):
    from oracle import pyref as pr
    n = 12
    s = oracle.fr_random(n, 5)
    b = _bases(oracle, n, 6)
    got = oracle.g1_affine_to_ints(_affine(oracle, ctx.msm(s, b)))
    want = pr.msm_naive(oracle.fr_from_mont(s), [oracle.g1_affine_to_ints(x) for x in b])
    assert got == want


def test_msm_edge_scalars(ctx, oracle):
    from oracle import pyref as pr
    n = 2048
    b = _bases(oracle, n, 77)
    zero = np.zeros((n, 4), dtype=np.uint64)
    out = ctx.msm(zero, b)
    assert not out[8:].any()  # identity: Z = 0
    for val in (1, 2, pr.R - 1, (1 << 128) - 1, 1 << 253, (pr.R - 1) // 2):
        s = np.tile(oracle.fr_to_mont([val])[0], (n, 1))
        assert np.array_equal(_affine(oracle, ctx.msm(s, b)), _affine(oracle, oracle.best_multiexp(s, b))), hex(val)
    # witness-like skew: mostly zeros and bits, a few small limbs
    rng = np.random.default_rng(3)
    vals = [int(v) for v in rng.choice([0, 0, 0, 1, 1, 2, 3, (1 << 88) - 1, 1 << 64], size=n)]
    s = oracle.fr_to_mont(vals)
    assert np.array_equal(_affine(oracle, ctx.msm(s, b)), _affine(oracle, oracle.best_multiexp(s, b)))


def test_msm_edge_bases(ctx, oracle):
    n = 512
    s = oracle.fr_random(n, 91)
    b = _bases(oracle, n, 92)
    b[5] = 0            # identity point (0,0)
    b[100] = b[7]       # duplicate base
    b[200:264] = b[9]   # a run of equal bases
    s[200:264] = s[9]   # ... with equal scalars: forces the doubling branch inside a bucket
    assert np.array_equal(_affine(oracle, ctx.msm(s, b)), _affine(oracle, oracle.best_multiexp(s, b)))
    # P and -P with the same scalar cancel
    from oracle import pyref as pr
    neg = b[:2].copy()
    p = oracle.g1_affine_to_ints(b[0])
    neg[1] = oracle.g1_ints_to_affine((p[0], (-p[1]) % pr.P))
    neg[0] = b[0]
    out = ctx.msm(np.tile(s[0], (2, 1)), neg)
    assert not out[8:].any()


@pytest.mark.parametrize("n", [1 << 10, 1 << 14, 1 << 16])
def test_msm_resident_srs_with_window_tables(zkw, oracle, n):
    """zkw_srs_load keeps g / g_lagrange on the device with their window tables; MSMs name them by id."""
    c = zkw.Context(0)
    try:
        g = _bases(oracle, n, 31)
        gl = _bases(oracle, n, 32)
        c.srs_load(g, gl)
        s = oracle.fr_random(n, 33)
        assert np.array_equal(c.msm(s, which=zkw.BASES_G)[:8], _affine(oracle, oracle.best_multiexp(s, g)))
        assert np.array_equal(c.msm(s, which=zkw.BASES_G_LAGRANGE)[:8], _affine(oracle, oracle.best_multiexp(s, gl)))
        # a shorter MSM against a prefix of the resident basis (commit of a low-degree polynomial)
        m = n // 2 + 3
        assert np.array_equal(c.msm(s[:m], which=zkw.BASES_G)[:8], _affine(oracle, oracle.best_multiexp(s[:m], g[:m])))
        # without tables (per-window bucket groups) the answer is the same
        c.msm_config(window_bits=0, precompute=False)
        c.srs_load(g, None)
        assert np.array_equal(c.msm(s, which=zkw.BASES_G)[:8], _affine(oracle, oracle.best_multiexp(s, g)))
    finally:
        c.close()


@pytest.mark.parametrize("kind", ["all_equal", "plus_minus_one", "sorted_small", "half_zero", "two_values", "run_aligned", "few_hundred_values"])
def test_msm_skewed_scalars_over_window_tables(zkw, oracle, kind):
    """Witness-shaped scalar vectors put most entries into a handful of buckets: the equal-run accumulation
    cuts those buckets into thousands of partials (slots t + b) that the queued heavy-bucket kernel folds."""
    from oracle import pyref as pr
    n = 1 << 15
    rng = np.random.default_rng(7)
    if kind == "all_equal":
        vals = [0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF] * n
    elif kind == "plus_minus_one":
        vals = [1 if i % 3 else pr.R - 1 for i in range(n)]
    elif kind == "sorted_small":
        vals = sorted(int(x) for x in rng.integers(0, 1 << 18, n))          # a permuted lookup column
    elif kind == "half_zero":
        vals = [0 if i % 2 else int(x) for i, x in enumerate(rng.integers(0, 1 << 62, n))]
    elif kind == "two_values":
        vals = [(1 << 200) + 5 if i < n // 3 else (1 << 16) - 1 for i in range(n)]   # digit 2^16 - 1 recodes to -1 with a carry
    elif kind == "few_hundred_values":
        # ~160 entries per bucket in window 0 and ~5000 in window 1: 10 and 300 partials - the warp-per-bucket ("medium") and
        # CTA-per-bucket queues of the combine step, like the top digits of the real witness's 88-bit limbs
        pool = [int(x) for x in rng.integers(1, 1 << 16, 200)]
        vals = [pool[int(j)] + (int(j) % 7 << 16) for j in rng.integers(0, 200, n)]
    else:
        vals = [(i // 16) + 1 for i in range(n)]                             # runs of 16 equal digits: run and bucket boundaries coincide
    s = oracle.fr_to_mont(vals)
    g = _bases(oracle, n, 77)
    want = _affine(oracle, oracle.best_multiexp(s, g))
    c = zkw.Context(0)
    try:
        c.srs_load(g, None)
        assert np.array_equal(c.msm(s, which=zkw.BASES_G)[:8], want)
        assert np.array_equal(c.msm(s, g)[:8], want)      # caller bases: one bucket set per window
    finally:
        c.close()


@pytest.mark.parametrize("sort", ["direct", "binned"])
@pytest.mark.parametrize("n", [33, 1000, 4097, (1 << 14) + 7])
def test_msm_both_entry_sorts(zkw, oracle, monkeypatch, sort, n):
    """The bucket order of the entries comes from one of two sorts (msm.cu): the direct one (a global atomic per entry) and
    the binned one (coarse bins, shared-memory counting, chunks staged by bulk copy).  Both are forced here on every size,
    with uniform and with skewed scalars, over window tables and over caller bases."""
    monkeypatch.setenv("ZKW_MSM_BINNED_SORT", "0" if sort == "direct" else "1")
    monkeypatch.setenv("ZKW_MSM_BINNED_MIN_ENTRIES", "0")
    g = _bases(oracle, n, 4000 + n)
    rng = np.random.default_rng(n)
    skew = [int(v) for v in rng.choice([0, 1, 1, 2, 3, (1 << 18) - 1, (1 << 88) - 1, 12345678901234567890], size=n)]
    runs = [(i // 100) * 0x10001000100010001 + 7 for i in range(n)]          # constant stretches: whole warps agree on every digit
    c = zkw.Context(0)
    try:
        c.srs_load(g, None)
        for s in (oracle.fr_random(n, 4100 + n), oracle.fr_to_mont(skew), oracle.fr_to_mont(runs)):
            want = _affine(oracle, oracle.best_multiexp(s, g))
            assert np.array_equal(c.msm(s, which=zkw.BASES_G)[:8], want)
            assert np.array_equal(c.msm(s, g)[:8], want)
    finally:
        c.close()


def test_msm_before_srs_load_is_an_error(zkw, oracle):
    c = zkw.Context(0)
    try:
        with pytest.raises(zkw.ZkwError) as ei:
            c.msm(oracle.fr_random(4, 1), which=zkw.BASES_G)
        assert ei.value.status == -5
    finally:
        c.close()


@pytest.mark.parametrize("c_bits", [4, 7, 11, 13, 16, 17, 19, 20, 22])
def test_msm_window_sizes(zkw, oracle, c_bits):
    c = zkw.Context(0)
    try:
        c.msm_config(window_bits=c_bits, precompute=True)
        n = 3000
        s = oracle.fr_random(n, 55)
        b = _bases(oracle, n, 56)
        want = _affine(oracle, oracle.best_multiexp(s, b))
        assert np.array_equal(c.msm(s, b)[:8], want)
        c.srs_load(b, None)
        assert np.array_equal(c.msm(s, which=zkw.BASES_G)[:8], want)
    finally:
        c.close()


def test_msm_full_size_split_identity_k19(zkw, oracle):
    """BASELINE size (2^19 points, resident SRS with window tables): the oracle needs minutes there,
    so check size-independent properties — msm(whole) = msm(low half) + msm(high half) with the halves
    taken through the other code path (prefix of the resident basis / caller bases), and one oracle run
    on a 2^16 prefix."""
    import ctypes as C
    n = 1 << 19
    c = zkw.Context(0)
    try:
        g = _bases(oracle, n, 61)
        c.srs_load(g, None)
        a = oracle.fr_random(n, 62)
        whole = c.msm(a, which=zkw.BASES_G)
        m = 1 << 16
        assert np.array_equal(c.msm(a[:m], which=zkw.BASES_G)[:8], _affine(oracle, oracle.best_multiexp(a[:m], g[:m])))
        lo = c.msm(a[: n // 2], which=zkw.BASES_G)
        hi = c.msm(a[n // 2:], g[n // 2:])
        s = np.empty(12, dtype=np.uint64)
        u64p = C.POINTER(C.c_uint64)
        oracle.lib().zko_g1_add(s.ctypes.data_as(u64p), lo.ctypes.data_as(u64p), hi.ctypes.data_as(u64p))
        assert np.array_equal(_affine(oracle, s), whole[:8])
    finally:
        c.close()


def _witness_like_columns(oracle, n, lookup_bits=18):
    """Scalar vectors shaped like the columns a k = 19 proof commits to: an advice column (bits, lookup limbs,
    88-bit limbs, a few full-width values), the sorted permuted lookup column A', and a grand-product column Z
    (full-width values with long constant stretches)."""
    from oracle import pyref as pr
    rng = np.random.default_rng(1919)
    kind = rng.integers(0, 8, n)
    small = rng.integers(0, 1 << lookup_bits, n, dtype=np.uint64)
    wide = oracle.fr_from_mont(oracle.fr_random(n // 8 + 1, 7))
    adv = []
    for i in range(n):
        t = int(kind[i])
        if t < 2:
            adv.append(int(small[i]) & 1)
        elif t < 5:
            adv.append(int(small[i]))
        elif t < 7:
            adv.append((int(small[i]) << 70) | int(small[(i + 1) % n]) << 20 | 5)    # 88-bit limb
        else:
            adv.append(wide[i // 8])
    a_prime = sorted(int(x) for x in small)
    z = []
    cur = wide[0]
    for i in range(n):
        if int(kind[i]) == 0 and i % 97 == 0:
            cur = wide[(i // 8) % len(wide)]
        z.append(cur if i % 5 else (cur * (i + 1)) % pr.R)
    return {"advice": adv, "a_prime": a_prime, "z": z}


def test_msm_full_size_k19_matches_oracle_bit_for_bit(zkw, oracle):
    """BASELINE size: ONE 2^19-point MSM over the resident window-table path (c = 16, 592-CTA run plan, heavy
    bucket queue) compared bit for bit with the oracle's best_multiexp — uniform scalars and the three
    witness-shaped columns a k = 19 proof commits to."""
    n = 1 << 19
    c = zkw.Context(0)
    try:
        g = _bases(oracle, n, 61)
        c.srs_load(g, None)
        s = oracle.fr_random(n, 62)
        assert np.array_equal(c.msm(s, which=zkw.BASES_G)[:8], _affine(oracle, oracle.best_multiexp(s, g)))
        for name, vals in _witness_like_columns(oracle, n).items():
            s = oracle.fr_to_mont(vals)
            assert np.array_equal(c.msm(s, which=zkw.BASES_G)[:8], _affine(oracle, oracle.best_multiexp(s, g))), name
    finally:
        c.close()


def test_dedicated_squaring_equals_general_product_on_device(ctx):
    """Fp::sqr_lazy (irregular-row interleaved Montgomery squaring, used twice per bucket addition) against mul_lazy(a, a), limb
    for limb, on 2^18 pseudo-random values over the whole lazy range [0, 2m) and the corners 0, 1, m - 1, m, m + 1, 2m - 1, for
    Fr and Fq."""
    for seed in (1, 2, 0xB200):
        assert ctx.selftest_field(seed=seed, count=1 << 18) == (0, 0)
