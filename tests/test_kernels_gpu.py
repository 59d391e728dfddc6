"""Parity of the standalone CUDA kernels (through the C ABI) against the oracle.  Bit-exact."""
import random

import pytest

from oracle.fields import R_MOD, Q_MOD
from oracle import curve, poly

pytestmark = pytest.mark.gpu

EDGE_FR = [0, 1, 2, R_MOD - 1, R_MOD - 2, (1 << 256) % R_MOD, (1 << 255) % R_MOD, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]
EDGE_FQ = [0, 1, 2, Q_MOD - 1, Q_MOD - 2, (1 << 384) % Q_MOD, 0xFFFFFFFF, 1 << 32, (1 << 380)]


def test_fr_field_ops(pmlib):
    from polymath_b200 import kernels
    rnd = random.Random(1)
    a = [x for x in EDGE_FR for _ in EDGE_FR] + [rnd.randrange(R_MOD) for _ in range(20000)]
    b = [y for _ in EDGE_FR for y in EDGE_FR] + [rnd.randrange(R_MOD) for _ in range(20000)]
    assert kernels.fr_mul_batch(a, b) == [x * y % R_MOD for x, y in zip(a, b)]
    assert kernels.fr_add_batch(a, b) == [(x + y) % R_MOD for x, y in zip(a, b)]
    assert kernels.fr_sub_batch(a, b) == [(x - y) % R_MOD for x, y in zip(a, b)]


def test_fq_mul(pmlib):
    from polymath_b200 import kernels
    rnd = random.Random(2)
    a = [x for x in EDGE_FQ for _ in EDGE_FQ] + [rnd.randrange(Q_MOD) for _ in range(20000)]
    b = [y for _ in EDGE_FQ for y in EDGE_FQ] + [rnd.randrange(Q_MOD) for _ in range(20000)]
    assert kernels.fq_mul_batch(a, b) == [x * y % Q_MOD for x, y in zip(a, b)]


def test_field_inverse(pmlib):
    """Binary extended-Euclid inverse (field.cuh: Fp::inv) against Fermat's little theorem; 0 -> 0."""
    from polymath_b200 import kernels
    rnd = random.Random(3)
    a = EDGE_FR + [1 << k for k in range(1, 255, 7)] + [rnd.randrange(R_MOD) for _ in range(3000)]
    assert kernels.fr_inv_batch(a) == [pow(x, R_MOD - 2, R_MOD) for x in a]
    b = EDGE_FQ + [1 << k for k in range(1, 381, 7)] + [Q_MOD - (1 << k) for k in range(0, 380, 11)] + [rnd.randrange(Q_MOD) for _ in range(3000)]
    assert kernels.fq_inv_batch(b) == [pow(x, Q_MOD - 2, Q_MOD) for x in b]


@pytest.mark.parametrize("log_n", list(range(0, 15)))
def test_ntt_matches_oracle(pmlib, log_n):
    from polymath_b200 import kernels
    rnd = random.Random(100 + log_n)
    n = 1 << log_n
    vals = [rnd.randrange(R_MOD) for _ in range(n)]
    dom = poly.Domain(n)
    fwd = kernels.ntt_fr(vals)
    assert fwd == dom.fft(vals)
    assert kernels.ntt_fr(vals, inverse=True) == dom.ifft(vals)
    if log_n <= 6:
        assert fwd == poly.naive_dft(vals, dom.group_gen)


@pytest.mark.parametrize("log_n", [16, 20, 21, 23])
def test_ntt_roundtrip_and_linearity_large(pmlib, log_n):
    """Size-independent properties at sizes the Python oracle cannot reach."""
    from polymath_b200 import kernels
    rnd = random.Random(7 + log_n)
    n = 1 << log_n
    seeds = [rnd.randrange(R_MOD) for _ in range(64)]
    vals = [seeds[i % 64] * (i + 1) % R_MOD for i in range(n)]
    fwd = kernels.ntt_fr(vals)
    assert kernels.ntt_fr(fwd, inverse=True) == vals
    # spot-check evaluations by Horner at a few domain points
    dom = poly.Domain(n)
    for idx in (0, 1, n // 2 + 3, n - 1):
        assert fwd[idx] == poly.poly_eval(vals, pow(dom.group_gen, idx, R_MOD))


def test_ntt_coset(pmlib):
    from polymath_b200 import kernels
    rnd = random.Random(5)
    n = 1 << 10
    g = 7
    vals = [rnd.randrange(R_MOD) for _ in range(n)]
    dom = poly.Domain(n)
    shifted = [v * pow(g, i, R_MOD) % R_MOD for i, v in enumerate(vals)]
    ev = kernels.ntt_fr(vals, coset_gen=g)
    assert ev == dom.fft(shifted)
    assert kernels.ntt_fr(ev, inverse=True, coset_gen=g) == vals


def test_g1_compression_roundtrip_and_rejects(pmlib):
    """Device (de)compression against the oracle's restatement of ark-bls12-381's zcash encoding."""
    from polymath_b200 import kernels
    from polymath_b200.lib import PolymathB200Error
    rnd = random.Random(21)
    pts = [curve.G1_GEN, None, curve.g1_neg(curve.G1_GEN)] + _bases(200, rnd)
    pts += [curve.g1_neg(p) for p in pts[3:40]]
    enc = b"".join(curve.g1_compress(p) for p in pts)
    assert kernels.g1_compress_batch(pts) == enc
    assert kernels.g1_decompress_batch(enc) == pts
    assert kernels.g1_decompress_batch(enc[:48 * 20], validate=True) == pts[:20]
    # an infinity flag on a non-zero payload / with the sort flag is refused, as by the host codec and the oracle
    for bad_inf in (bytes([0xC0]) + b"\x01" * 47, bytes([0xE0]) + bytes(47), bytes([0xC1]) + bytes(47)):
        with pytest.raises(PolymathB200Error) as ei:
            kernels.g1_decompress_batch(enc[:48] + bad_inf)
        assert "infinity flag" in str(ei.value) and "index 1" in str(ei.value)
        with pytest.raises(ValueError):
            curve.g1_decompress(bad_inf)
    # a curve point outside the prime-order subgroup
    x = 1
    while True:
        rhs = (x ** 3 + 4) % Q_MOD
        y = pow(rhs, (Q_MOD + 1) // 4, Q_MOD)
        if y * y % Q_MOD == rhs and not curve.g1_in_subgroup((x, y)):
            break
        x += 1
    off = curve.g1_compress((x, y))
    assert kernels.g1_decompress_batch(enc[:96] + off) == pts[:2] + [(x, y)]
    bad_cases = {
        "prime-order subgroup": (enc[:96] + off, True),
        "compression flag": (enc[:48] + bytes(48), False),
        "field modulus": (bytes([0x9F]) + b"\xff" * 47, False),
    }
    nr = 1
    while pow((nr ** 3 + 4) % Q_MOD, (Q_MOD - 1) // 2, Q_MOD) == 1:
        nr += 1
    bad_cases["curve point"] = (enc[:144] + bytes([0x80]) + nr.to_bytes(47, "big"), False)
    for needle, (data, validate) in bad_cases.items():
        with pytest.raises(PolymathB200Error) as ei:
            kernels.g1_decompress_batch(data, validate=validate)
        assert needle in str(ei.value), (needle, str(ei.value))
    assert "index 3" in str(ei.value)


def _bases(k, rnd):
    tbl = curve.FixedBaseTable(curve.G1_GEN, window=8)
    return tbl.mul_many([rnd.randrange(1, R_MOD) for _ in range(k)])


def test_fixed_base_mul(pmlib):
    from polymath_b200 import kernels
    rnd = random.Random(11)
    scalars = [0, 1, 2, R_MOD - 1, (1 << 255) % R_MOD] + [rnd.randrange(R_MOD) for _ in range(300)]
    got = kernels.fixed_base_mul(scalars)
    tbl = curve.FixedBaseTable(curve.G1_GEN, window=8)
    assert got == tbl.mul_many(scalars)
    assert got[0] is None and got[1] == curve.G1_GEN


@pytest.mark.parametrize("n,c", [(1, 0), (2, 0), (7, 0), (33, 0), (300, 0), (300, 4), (1500, 0), (1500, 13), (4096, 0)])
def test_msm_matches_oracle(pmlib, n, c):
    from polymath_b200 import kernels
    rnd = random.Random(1000 + n + c)
    bases = _bases(n, rnd)
    scalars = [rnd.randrange(R_MOD) for _ in range(n)]
    want = poly.msm_pippenger(scalars, bases)
    assert kernels.msm_g1(bases, scalars, window_bits=c) == want
    if n <= 33:
        assert want == poly.msm_naive(scalars, bases)


@pytest.mark.parametrize("c,levels", [(10, 2), (10, 5), (10, 26), (13, 20), (16, 16), (7, 3)])
def test_msm_precomputed_levels(pmlib, c, levels):
    """Fixed-base tables: window w = g*levels + l reads 2^(c*l) P and shares group g's buckets."""
    from polymath_b200 import kernels
    rnd = random.Random(c * 100 + levels)
    n = 700
    bases = _bases(n, rnd)
    bases[5] = None
    scalars = [rnd.randrange(R_MOD) for _ in range(n)]
    scalars[7] = 0
    scalars[8] = R_MOD - 1
    assert kernels.msm_g1(bases, scalars, window_bits=c, levels=levels) == poly.msm_pippenger(scalars, bases)


def test_msm_edge_cases(pmlib):
    """Infinity bases, zero scalars, repeated points, +/- pairs, skewed scalars, heavy buckets, 104-byte stride."""
    from polymath_b200 import kernels
    rnd = random.Random(77)
    n = 600
    bases = _bases(n, rnd)
    for i in range(0, n, 7):
        bases[i] = None
    for i in range(3, n, 50):
        bases[i] = bases[i - 1]                     # duplicate points
    bases[11] = curve.g1_neg(bases[10])             # P and -P
    hot = rnd.randrange(R_MOD)
    scalars = [hot if i % 3 else rnd.randrange(R_MOD) for i in range(n)]
    for i in range(0, n, 13):
        scalars[i] = 0
    scalars[10] = scalars[11] = 5
    scalars[20] = R_MOD - 1
    scalars[21] = (1 << 255) % R_MOD
    want = poly.msm_pippenger(scalars, bases)
    assert kernels.msm_g1(bases, scalars) == want
    assert kernels.msm_g1(bases, scalars, window_bits=8, heavy_threshold=16) == want   # force the heavy-bucket path
    assert kernels.msm_g1(bases, scalars, stride=104) == want
    # everything cancels / empty
    assert kernels.msm_g1([bases[1], curve.g1_neg(bases[1])], [9, 9]) is None
    assert kernels.msm_g1([], []) is None
    assert kernels.msm_g1(bases[:5], [0] * 5) is None
    # msm_unchecked truncates to the shorter input (prover.rs:380-384 asserts scalars <= bases)
    assert kernels.msm_g1(bases, scalars[:100]) == poly.msm_pippenger(scalars[:100], bases[:100])


@pytest.mark.parametrize("rounds", [-1, 0, 3])
def test_msm_skewed_inputs(pmlib, rounds):
    """SURVEY.md 8d skew: 89 % of the scalars one repeated value, 10 % zero, 1 % uniform, 1 % infinity bases — whole
    warps land in ONE bucket per window (the warp-aggregated atomics of the sort, gated by the neighbour probe), mixed
    with lanes that do not (partial groups), through the walk, the heavy-run path and the pair rounds."""
    from polymath_b200 import kernels
    rnd = random.Random(8100 + rounds)
    n = 3000
    bases = _bases(n, rnd)
    hot = rnd.randrange(R_MOD)
    scalars = []
    for i in range(n):
        h = rnd.randrange(1000)
        scalars.append(hot if h < 890 else 0 if h < 990 else rnd.randrange(R_MOD))
        if rnd.randrange(100) == 0:
            bases[i] = None
    # runs of equal scalars shorter and longer than a warp, and an isolated pair of equal neighbours
    for i in range(64, 64 + 40):
        scalars[i] = 12345
    scalars[200] = scalars[201] = R_MOD - 2
    want = poly.msm_pippenger(scalars, bases)
    kernels.msm_set_tuning(rounds)
    try:
        assert kernels.msm_g1(bases, scalars) == want
        assert kernels.msm_g1(bases, scalars, window_bits=9, heavy_threshold=16) == want
        assert kernels.msm_g1(bases, scalars, window_bits=10, levels=4) == want
    finally:
        kernels.msm_set_tuning(-1)


@pytest.fixture
def msm_tuning(pmlib):
    from polymath_b200 import kernels
    yield kernels.msm_set_tuning
    kernels.msm_set_tuning(-1)


@pytest.mark.parametrize("n,c,rounds", [(1500, 6, 1), (1500, 6, 2), (1500, 6, 3), (1500, 6, 6), (1500, 6, 9), (4096, 4, 5),
                                         (300, 8, 2), (33, 4, 4), (2, 4, 1), (1, 0, 3)])
def test_msm_pair_rounds(msm_tuning, n, c, rounds):
    """Batched-affine pair rounds (k_pairs_forward / k_batch_invert / k_pairs_backward) before the XYZZ walk."""
    from polymath_b200 import kernels
    rnd = random.Random(4000 + n + c)
    bases = _bases(n, rnd)
    scalars = [rnd.randrange(R_MOD) for _ in range(n)]
    want = poly.msm_pippenger(scalars, bases)
    msm_tuning(rounds)
    assert kernels.msm_g1(bases, scalars, window_bits=c) == want
    if c:
        assert kernels.msm_g1(bases, scalars, window_bits=c, levels=3) == want


@pytest.mark.parametrize("sets,rounds", [(1, 0), (1, 3), (3, 2), (5, 0)])
def test_msm_bucket_set_passes(msm_tuning, monkeypatch, sets, rounds):
    """Large MSMs run their bucket sets in several passes (address space / workspace): forced on a small one."""
    from polymath_b200 import kernels
    monkeypatch.setenv("PM_MSM_SETS_PER_PASS", str(sets))
    rnd = random.Random(500 + sets)
    n = 1200
    bases = _bases(n, rnd)
    bases[3] = None
    hot = rnd.randrange(R_MOD)
    scalars = [hot if i % 4 == 0 else rnd.randrange(R_MOD) for i in range(n)]
    want = poly.msm_pippenger(scalars, bases)
    msm_tuning(rounds)
    assert kernels.msm_g1(bases, scalars, window_bits=7) == want                       # 37 bucket sets
    assert kernels.msm_g1(bases, scalars, window_bits=7, heavy_threshold=16) == want
    assert kernels.msm_g1(bases, scalars, window_bits=9, levels=4) == want              # 8 sets of 4 levels
    assert kernels.msm_g1(bases, scalars, window_bits=9, levels=29) == want             # one set


@pytest.mark.parametrize("rounds", [1, 2, 3, 5, 8])
def test_msm_pair_rounds_exceptional_pairs(msm_tuning, rounds):
    """P + P, P + (-P), infinity operands and results inside the pair rounds; heavy buckets skipped by them."""
    from polymath_b200 import kernels
    rnd = random.Random(78)
    distinct = _bases(6, rnd)
    n = 900
    # few distinct points, few distinct scalars: every bucket run is full of equal and opposite points
    bases = [distinct[rnd.randrange(6)] for _ in range(n)]
    for i in range(0, n, 5):
        bases[i] = curve.g1_neg(bases[i - 1])
    for i in range(0, n, 11):
        bases[i] = None
    pool = [rnd.randrange(R_MOD) for _ in range(4)] + [1, 2, R_MOD - 1, 3]
    scalars = [pool[rnd.randrange(len(pool))] for _ in range(n)]
    for i in range(0, n, 17):
        scalars[i] = 0
    want = poly.msm_pippenger(scalars, bases)
    msm_tuning(rounds)
    assert kernels.msm_g1(bases, scalars, window_bits=5) == want
    assert kernels.msm_g1(bases, scalars, window_bits=8, heavy_threshold=16) == want
    assert kernels.msm_g1(bases, scalars, window_bits=7, levels=4) == want
    # the run of one bucket is P, P, P, ... (pure doubling tree) and P, -P, P, -P (all cancel)
    assert kernels.msm_g1([distinct[0]] * 64, [5] * 64, window_bits=4) == curve.g1_mul(distinct[0], 320)
    assert kernels.msm_g1([distinct[0], curve.g1_neg(distinct[0])] * 32, [5] * 64, window_bits=4) is None
    assert kernels.msm_g1([distinct[0], curve.g1_neg(distinct[0])] * 32 + [distinct[1]], [5] * 65, window_bits=4) == curve.g1_mul(distinct[1], 5)


def test_msm_linearity_large(pmlib):
    """2^18 points: MSM(k*s) == k*MSM(s) and MSM(s) + MSM(t) == MSM(s+t), bases generated on the device."""
    from polymath_b200 import kernels
    rnd = random.Random(5)
    n = 1 << 18
    seeds = [rnd.randrange(R_MOD) for _ in range(97)]
    base_scalars = [seeds[i % 97] * (i + 1) % R_MOD for i in range(n)]
    bases = kernels.fixed_base_mul(base_scalars)
    s = [seeds[(i * 7) % 97] * (i + 3) % R_MOD for i in range(n)]
    # sum_i s_i * (b_i * G) = (sum_i s_i * b_i) * G
    dot = sum(x * y for x, y in zip(s, base_scalars)) % R_MOD
    assert kernels.msm_g1(bases, s) == curve.g1_mul(curve.G1_GEN, dot)
    kernels.msm_set_tuning(0)                         # XYZZ walk only
    try:
        assert kernels.msm_g1(bases, s) == curve.g1_mul(curve.G1_GEN, dot)
    finally:
        kernels.msm_set_tuning(-1)
