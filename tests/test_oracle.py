"""Pins of the oracle itself (CPU).  The reference holds no golden vectors, so the
oracle is anchored on public constants, published known-answer vectors of its dependencies,
naive-vs-fast cross checks, and the only property the reference's own tests assert:
`verify(prove(..)) == true` on tests/dummy.rs and tests/mimc.rs.  PARITY UNPINNED otherwise."""
import copy
import hashlib
import json
import os
import random

import pytest

from oracle import curve, fields, merlin, pairing, poly, polymath as opm, r1cs as orc, rng as orng
from oracle.fields import R_MOD, Q_MOD

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_field_constants():
    assert R_MOD.bit_length() == 255 and Q_MOD.bit_length() == 381
    assert (R_MOD - 1) % (1 << 32) == 0 and (R_MOD - 1) % (1 << 33) != 0
    w = fields.FR_TWO_ADIC_ROOT
    assert pow(w, 1 << 32, R_MOD) == 1 and pow(w, 1 << 31, R_MOD) != 1
    assert w == 10238227357739495823651030575849232062558860180284477541189508159991286009131
    assert fields.FR_MONT_R == 0x1824B159ACC5056F998C4FEFECBC4FF55884B7FA0003480200000001FFFFFFFE
    assert (-pow(R_MOD, -1, 1 << 64)) % (1 << 64) == 0xFFFFFFFEFFFFFFFF
    assert (-pow(Q_MOD, -1, 1 << 64)) % (1 << 64) == 0x89F3FFFCFFFCFFFD


def test_curve_constants_and_encoding():
    assert curve.g1_is_on_curve(curve.G1_GEN) and curve.g2_is_on_curve(curve.G2_GEN)
    assert curve.g1_mul(curve.G1_GEN, R_MOD - 1) == curve.g1_neg(curve.G1_GEN)
    assert curve.g2_mul(curve.G2_GEN, R_MOD) is None
    # zcash-format generator encodings (published with the BLS12-381 spec)
    assert curve.g1_compress(curve.G1_GEN).hex() == (
        "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
    assert curve.g2_compress(curve.G2_GEN).hex().startswith("93e02b6052719f607dacd3a088274f65596bd0d09920b61a")
    assert curve.g1_compress(None)[0] == 0xC0


def test_keccak_against_sha3():
    def sha3_256(msg):
        st, rate = bytearray(200), 136
        m = bytearray(msg) + b"\x06"
        m += b"\x00" * ((-len(m)) % rate)
        m[-1] |= 0x80
        for i in range(0, len(m), rate):
            for j in range(rate):
                st[j] ^= m[i + j]
            merlin.keccak_f1600(st)
        return bytes(st[:32])
    for msg in (b"", b"abc", b"x" * 500):
        assert sha3_256(msg) == hashlib.sha3_256(msg).digest()


def test_merlin_known_answer():
    t = merlin.MerlinTranscript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_chacha_core_rfc7539():
    key = [int.from_bytes(bytes(range(32))[4 * i:4 * i + 4], "little") for i in range(8)]
    blk = orng.chacha_block(key, [1, 0x09000000, 0x4A000000, 0], 20)
    out = b"".join(w.to_bytes(4, "little") for w in blk)
    assert out.hex().startswith("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e")


def test_pairing_bilinear():
    a, b = 1234567, 7654321
    e1 = pairing.pairing(curve.g1_mul(curve.G1_GEN, a), curve.g2_mul(curve.G2_GEN, b))
    e2 = fields.fq12_pow(pairing.pairing(curve.G1_GEN, curve.G2_GEN), a * b)
    assert e1 == e2 and e1 != fields.FQ12_ONE
    assert pairing.pairing_product_is_one([(curve.g1_mul(curve.G1_GEN, a), curve.G2_GEN),
                                           (curve.g1_neg(curve.G1_GEN), curve.g2_mul(curve.G2_GEN, a))])


def test_ntt_against_naive_dft():
    rnd = random.Random(3)
    for log_n in range(0, 7):
        n = 1 << log_n
        v = [rnd.randrange(R_MOD) for _ in range(n)]
        d = poly.Domain(n)
        assert d.fft(v) == poly.naive_dft(v, d.group_gen)
        assert d.ifft(d.fft(v)) == v
    d = poly.Domain(5)
    assert d.size == 8 and pow(d.group_gen, 8, R_MOD) == 1 and pow(d.group_gen, 4, R_MOD) != 1


def test_msm_pippenger_against_naive():
    rnd = random.Random(4)
    tbl = curve.FixedBaseTable(curve.G1_GEN, window=8)
    bases = tbl.mul_many([rnd.randrange(R_MOD) for _ in range(40)])
    bases[3] = None
    scalars = [rnd.randrange(R_MOD) for _ in range(40)]
    scalars[5] = 0
    assert poly.msm_pippenger(scalars, bases) == poly.msm_naive(scalars, bases)
    assert tbl.mul_many([5])[0] == curve.g1_mul(curve.G1_GEN, 5)


def test_lagrange_coefficients():
    d = poly.Domain(8)
    tau = 123456789
    lag = d.evaluate_all_lagrange_coefficients(tau)
    assert sum(lag) % R_MOD == 1
    vals = [5, 7, 11, 13, 17, 19, 23, 29]
    coeffs = d.ifft(vals)
    assert poly.poly_eval(coeffs, tau) == sum(a * b for a, b in zip(lag, vals)) % R_MOD


def _dummy_flow(seed, literal):
    rng = orng.StdRng.seed_from_u64(seed)
    pk = opm.generate_proving_key(orc.DummyCircuit(), rng, literal=literal)
    a, b = orng.fr_rand(rng), orng.fr_rand(rng)
    proof = opm.create_proof(orc.DummyCircuit(a, b), pk, rng, literal=literal)
    return pk, proof, a * b % R_MOD


def test_dummy_circuit_verifies_and_literal_equals_sparse():
    """tests/dummy.rs:37-74: setup -> prove -> assert verify; and the literal dense data flow of the
    reference equals the sparse closed form the device uses."""
    pk_l, proof_l, pub = _dummy_flow(0, literal=True)
    pk_s, proof_s, _ = _dummy_flow(0, literal=False)
    assert proof_l == proof_s
    assert pk_l.uj_wj_lcs_by_y_alpha_g1 == pk_s.uj_wj_lcs_by_y_alpha_g1
    assert (pk_l.vk.n, pk_l.vk.m0, pk_l.vk.sigma) == (8, 2, 11)
    assert len(pk_l.x_powers_y_gamma_z_g1) == 103 and len(pk_l.uj_wj_lcs_by_y_alpha_g1) == 7
    assert opm.verify_proof(pk_l.vk, proof_l, [pub])
    assert not opm.verify_proof(pk_l.vk, proof_l, [(pub + 1) % R_MOD])
    assert len(proof_l.serialize_compressed()) == 176 and len(pk_l.vk.serialize_compressed()) == 392


def test_small_mimc_literal_equals_sparse_and_verifies():
    rng = orng.StdRng.seed_from_u64(1)
    consts = [orng.fr_rand(rng) for _ in range(5)]
    pk = opm.generate_proving_key(orc.MiMCDemo(None, None, consts), copy.deepcopy(rng), literal=False)
    pk_l = opm.generate_proving_key(orc.MiMCDemo(None, None, consts), copy.deepcopy(rng), literal=True)
    assert pk.uj_wj_lcs_by_y_alpha_g1 == pk_l.uj_wj_lcs_by_y_alpha_g1
    xl, xr = 111, 222
    p1 = opm.create_proof(orc.MiMCDemo(xl, xr, consts), pk, copy.deepcopy(rng), literal=False)
    p2 = opm.create_proof(orc.MiMCDemo(xl, xr, consts), pk, copy.deepcopy(rng), literal=True)
    assert p1 == p2
    assert opm.verify_proof(pk.vk, p1, [orc.mimc_hash(xl, xr, consts)])


def test_mimc322_reference_configuration_verifies():
    """tests/mimc.rs:145-216 with MIMC_ROUNDS = 322 (n = 2048): the reference's own acceptance check."""
    rng = orng.StdRng.seed_from_u64(1)
    consts = [orng.fr_rand(rng) for _ in range(322)]
    pk = opm.generate_proving_key(orc.MiMCDemo(None, None, consts), rng)
    assert (pk.vk.n, pk.sap_matrices.size()) == (2048, (1292, 1295))
    assert len(pk.x_powers_y_gamma_z_g1) == 20503 and len(pk.uj_wj_lcs_by_y_alpha_g1) == 1293
    xl, xr = orng.fr_rand(rng), orng.fr_rand(rng)
    proof = opm.create_proof(orc.MiMCDemo(xl, xr, consts), pk, rng)
    assert opm.verify_proof(pk.vk, proof, [orc.mimc_hash(xl, xr, consts)])


def test_unsatisfied_witness_asserts():
    rng = orng.StdRng.seed_from_u64(2)
    pk = opm.generate_proving_key(orc.DummyCircuit(), rng)
    with pytest.raises(AssertionError):
        opm.create_proof_with_assignment(pk, [1, 7], [2, 3], rng)


def test_golden_fixtures_are_reproduced():
    """tests/golden/*.json were generated by tests/golden/make_golden.py from this oracle; they freeze its
    outputs so that a change of the oracle (or of the device path checked against them) is visible."""
    path = os.path.join(GOLDEN, "dummy_seed0.json")
    g = json.load(open(path))
    pk, proof, pub = _dummy_flow(g["seed"], literal=False)
    assert proof.serialize_compressed().hex() == g["proof_hex"]
    assert pk.vk.serialize_compressed().hex() == g["vk_hex"]
    assert pub == int(g["public_input"])


def test_g1_decompress_restatement():
    """g1_decompress inverts g1_compress on the pinned generator encoding and on random subgroup points; rejects
    x >= q, non-residues, missing compression flag and (validating) points outside the prime-order subgroup."""
    import random
    from oracle.fields import Q_MOD
    rnd = random.Random(11)
    assert curve.g1_decompress(curve.g1_compress(curve.G1_GEN)) == curve.G1_GEN
    assert curve.g1_decompress(curve.g1_compress(None)) is None
    for _ in range(6):
        p = curve.g1_mul(curve.G1_GEN, rnd.randrange(1, R_MOD))
        for q in (p, curve.g1_neg(p)):
            assert curve.g1_decompress(curve.g1_compress(q)) == q
    with pytest.raises(ValueError):
        curve.g1_decompress(bytes(48))                                   # compression flag missing
    with pytest.raises(ValueError):
        curve.g1_decompress(bytes([0x9F]) + b"\xff" * 47)                # x >= q
    # a curve point outside the subgroup: accepted unchecked, rejected when validating
    x = 1
    while True:
        rhs = (x ** 3 + 4) % Q_MOD
        y = pow(rhs, (Q_MOD + 1) // 4, Q_MOD)
        if y * y % Q_MOD == rhs and not curve.g1_in_subgroup((x, y)):
            break
        x += 1
    enc = curve.g1_compress((x, y))
    assert curve.g1_decompress(enc, validate=False) == (x, y)
    with pytest.raises(ValueError):
        curve.g1_decompress(enc, validate=True)
    # a non-residue abscissa
    x = 1
    while pow((x ** 3 + 4) % Q_MOD, (Q_MOD - 1) // 2, Q_MOD) == 1:
        x += 1
    with pytest.raises(ValueError):
        curve.g1_decompress(bytes([0x80 | (x >> 376)]) + (x % (1 << 376)).to_bytes(47, "big"), validate=False)
