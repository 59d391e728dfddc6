"""Host verifier of the C++ mirror (`pm_polymath_verify`, /root/reference/src/verifier.rs:19-62) against the oracle.

No GPU: the pairing check stays on the host (BASELINE.json north_star).  The oracle (oracle/pairing.py,
oracle/polymath.py) is the checker; the golden proofs were produced by the oracle prover
(tests/golden/make_golden.py) and are accepted by the oracle verifier in tests/test_oracle.py.
"""
import ctypes as C
import json
import os
import random

import pytest

from oracle import curve as oc
from oracle import pairing as opair
from oracle.fields import R_MOD
from polymath_b200 import codec
from polymath_b200.api import Polymath, _bind
from polymath_b200.lib import PolymathB200Error, load

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _golden(name):
    d = json.load(open(os.path.join(GOLDEN, name)))
    pub = [int(d["public_input"])] if "public_input" in d else [int(d["image"])]
    return bytes.fromhex(d["vk_hex"]), pub, bytes.fromhex(d["proof_hex"])


def _g2_wire(q):
    if q is None:
        return bytes(192)
    (x0, x1), (y0, y1) = q
    return b"".join(codec.fq_to_wire(v) for v in (x0, x1, y0, y1))


def _pairing_is_one(pairs):
    lib = load()
    _bind(lib)
    ok = C.c_int(-1)
    g1 = b"".join(codec.g1_to_wire(p) for p, _ in pairs)
    g2 = b"".join(_g2_wire(q) for _, q in pairs)
    rc = lib.pm_host_pairing_product_is_one(g1, g2, len(pairs), C.byref(ok))
    assert rc == 0, lib.pm_last_error()
    return bool(ok.value)


def test_pairing_bilinearity_matches_oracle():
    rnd = random.Random(11)
    g1, g2 = oc.G1_GEN, oc.G2_GEN
    for _ in range(3):
        a, b = rnd.randrange(1, R_MOD), rnd.randrange(1, R_MOD)
        pa, qb = oc.g1_mul(g1, a), oc.g2_mul(g2, b)
        pab = oc.g1_mul(g1, a * b % R_MOD)
        # e(aP, bQ) * e(-abP, Q) == 1
        good = [(pa, qb), (oc.g1_neg(pab), g2)]
        bad = [(pa, qb), (oc.g1_neg(oc.g1_mul(g1, (a * b + 1) % R_MOD)), g2)]
        assert _pairing_is_one(good)
        assert not _pairing_is_one(bad)
    # infinity operands contribute the factor one; the oracle agrees on a small case
    assert _pairing_is_one([(None, g2), (g1, None)])
    small = [(oc.g1_mul(g1, 5), oc.g2_mul(g2, 7)), (oc.g1_neg(oc.g1_mul(g1, 35)), g2)]
    assert _pairing_is_one(small) == opair.pairing_product_is_one(small) is True
    assert not _pairing_is_one([(g1, g2)])


@pytest.mark.parametrize("name", ["dummy_seed0.json", "mimc8_seed7.json", "mimc322_seed1.json"])
def test_verify_accepts_golden_proofs_and_rejects_tampering(name):
    vk, pub, proof = _golden(name)
    assert Polymath.verify(vk, pub, proof) is True
    # wrong public input (the reference has no negative test; the oracle verifier rejects the same way)
    assert Polymath.verify(vk, [(pub[0] + 1) % R_MOD], proof) is False
    # a(x1) changed: still canonical, must be rejected
    a_at = int.from_bytes(proof[96:128], "little")
    bad = proof[:96] + ((a_at + 1) % R_MOD).to_bytes(32, "little") + proof[128:]
    assert Polymath.verify(vk, pub, bad) is False
    # [a]_1 and [c]_1 swapped: valid points, wrong proof
    assert Polymath.verify(vk, pub, proof[48:96] + proof[:48] + proof[96:]) is False


def test_verify_rejects_malformed_encodings():
    vk, pub, proof = _golden("dummy_seed0.json")
    # non-canonical scalar (>= r)
    bad = proof[:96] + (R_MOD).to_bytes(32, "little") + proof[128:]
    with pytest.raises(PolymathB200Error):
        Polymath.verify(vk, pub, bad)
    # compression flag cleared
    with pytest.raises(PolymathB200Error):
        Polymath.verify(vk, pub, bytes([proof[0] & 0x7F]) + proof[1:])
    # x coordinate not on the curve: search a few single-byte perturbations for one that fails to decode
    hit = False
    for delta in range(1, 40):
        cand = proof[:47] + bytes([(proof[47] + delta) & 0xFF]) + proof[48:]
        try:
            assert Polymath.verify(vk, pub, cand) is False      # decodes to another subgroup point: plain reject
        except PolymathB200Error:
            hit = True
            break
    assert hit, "expected at least one perturbed x without a curve/subgroup point"
    # truncated G2 flag in the key
    with pytest.raises(PolymathB200Error):
        Polymath.verify(vk[:48] + bytes([vk[48] & 0x7F]) + vk[49:], pub, proof)


def test_verify_matches_oracle_verifier_on_fresh_proofs():
    """Oracle prover -> product verifier and oracle verifier agree (accept and reject)."""
    from oracle import polymath as opm, r1cs as orc
    from oracle.rng import StdRng as ORng, fr_rand
    rng = ORng.seed_from_u64(99)
    pk = opm.generate_proving_key(orc.DummyCircuit(), rng)
    vk_bytes = pk.vk.serialize_compressed()
    proofs, pubs = [], []
    for _ in range(3):
        x, y = fr_rand(rng), fr_rand(rng)
        cs = orc.synthesize(orc.DummyCircuit(x, y), setup_mode=False)
        pr = opm.create_proof_with_assignment(pk, cs.instance_assignment, cs.witness_assignment, rng)
        pub = [x * y % R_MOD]
        assert opm.verify_proof(pk.vk, pr, pub)
        assert Polymath.verify(vk_bytes, pub, pr.serialize_compressed())
        proofs.append(pr.serialize_compressed())
        pubs.append(pub)
    # batch: all good -> accept; one public input off -> reject; proofs permuted against inputs -> reject
    seed = bytes(range(32))
    assert Polymath.verify_batch(vk_bytes, pubs, proofs, seed) is True
    assert Polymath.verify_batch(vk_bytes, pubs[:1], proofs[:1], seed) is True
    assert Polymath.verify_batch(vk_bytes, [pubs[0], [(pubs[1][0] + 1) % R_MOD], pubs[2]], proofs, seed) is False
    assert Polymath.verify_batch(vk_bytes, pubs, [proofs[1], proofs[0], proofs[2]], seed) is False
    assert Polymath.verify_batch(vk_bytes, [], [], seed) is True


def test_public_inputs_are_never_reduced():
    """x and x + r must not alias.  The Python binding refuses values outside [0, r); the C ABI refuses raw
    limb vectors >= r (PM_ERR_ARG) instead of running unreduced arithmetic on them."""
    import ctypes as C
    from polymath_b200.api import _bind
    from polymath_b200.lib import load
    vk_bytes, pub, proof = _golden("mimc8_seed7.json")
    assert Polymath.verify(vk_bytes, pub, proof) is True
    with pytest.raises(ValueError):
        Polymath.verify(vk_bytes, [pub[0] + R_MOD], proof)
    with pytest.raises(ValueError):
        Polymath.verify(vk_bytes, [-1], proof)
    with pytest.raises(ValueError):
        Polymath.verify_batch(vk_bytes, [[pub[0] + R_MOD]], [proof], bytes(32))
    lib = load()
    _bind(lib)
    ok = C.c_int(7)
    # Montgomery limbs of the valid input plus r: still < 2^256 for most values, never < r
    from polymath_b200 import codec
    limbs = int.from_bytes(codec.fr_to_wire(pub[0]), "little") + R_MOD
    if limbs < 1 << 256:
        rc = lib.pm_polymath_verify(vk_bytes, limbs.to_bytes(32, "little"), 1, proof, C.byref(ok))
        assert rc == 2 and ok.value == 0
    rc = lib.pm_polymath_verify(vk_bytes, b"\xff" * 32, 1, proof, C.byref(ok))
    assert rc == 2 and ok.value == 0
    rc = lib.pm_polymath_verify_batch(vk_bytes, 1, b"\xff" * 32, 1, proof, bytes(32), C.byref(ok))
    assert rc == 2 and ok.value == 0


def test_verify_batch_empty_needs_no_seed():
    """count == 0 accepts without touching the seed (NULL allowed); coefficients are bound to the statements, so the
    verdict does not depend on the seed value."""
    import ctypes as C
    from polymath_b200.api import _bind
    from polymath_b200.lib import load
    vk_bytes, pub, proof = _golden("mimc8_seed7.json")
    lib = load()
    _bind(lib)
    ok = C.c_int(0)
    assert lib.pm_polymath_verify_batch(vk_bytes, 0, None, 1, None, None, C.byref(ok)) == 0 and ok.value == 1
    for seed in (bytes(32), bytes(range(32)), b"\xaa" * 32):
        assert Polymath.verify_batch(vk_bytes, [pub, pub], [proof, proof], seed) is True
        bad = proof[:96] + bytes(32) + proof[128:]
        assert Polymath.verify_batch(vk_bytes, [pub, pub], [proof, bad], seed) is False


def test_verify_edge_encodings():
    """Infinity commitments, a zero evaluation, a proof from another statement: decoded like ark-serialize, rejected
    without a crash or a false accept."""
    vk_bytes, pub, proof = _golden("mimc8_seed7.json")
    inf = bytes([0xC0]) + bytes(47)
    cases = [
        inf + proof[48:],                                   # [a]_1 = O
        proof[:48] + inf + proof[96:],                      # [c]_1 = O
        proof[:128] + inf,                                  # [d]_1 = O
        inf + inf + bytes(32) + inf,                        # everything trivial
        proof[:96] + bytes(32) + proof[128:],               # a(x1) = 0
    ]
    for cand in cases:
        got = Polymath.verify(vk_bytes, pub, cand)
        assert got is False
    # infinity with stray bits must not decode
    with pytest.raises(PolymathB200Error):
        Polymath.verify(vk_bytes, pub, bytes([0xC0]) + bytes(46) + b"\x01" + proof[48:])
    with pytest.raises(PolymathB200Error):
        Polymath.verify(vk_bytes, pub, bytes([0xE0]) + bytes(47) + proof[48:])     # infinity + sort flag
    # a valid proof of ANOTHER key/statement is rejected under this key
    vk2, pub2, proof2 = _golden("dummy_seed0.json")
    assert Polymath.verify(vk_bytes, pub, proof2) is False
    assert Polymath.verify(vk2, pub2, proof) is False
    # more / fewer public inputs than the key expects: plain reject, like the reference (no length check there)
    assert Polymath.verify(vk_bytes, pub + [5], proof) is False
    assert Polymath.verify(vk_bytes, [], proof) is False


def test_pairing_product_with_many_pairs():
    """More than eight live pairs (the shared Miller loop folds the rest): prod e(a_i P, Q) * e(-(sum a_i) P, Q) == 1."""
    rnd = random.Random(5)
    g1, g2 = oc.G1_GEN, oc.G2_GEN
    coeffs = [rnd.randrange(1, 1 << 64) for _ in range(11)]
    q = oc.g2_mul(g2, 9)
    pairs = [(oc.g1_mul(g1, a), q) for a in coeffs] + [(oc.g1_neg(oc.g1_mul(g1, sum(coeffs) % R_MOD)), q)]
    assert _pairing_is_one(pairs)
    assert not _pairing_is_one(pairs[:-1] + [(oc.g1_neg(oc.g1_mul(g1, (sum(coeffs) + 1) % R_MOD)), q)])
    # infinity entries interleaved do not disturb the product
    assert _pairing_is_one(pairs[:5] + [(None, q), (g1, None)] + pairs[5:])


def _host_hash(kind, data):
    lib = load()
    _bind(lib)
    out = C.create_string_buffer(32)
    assert lib.pm_host_hash(kind, data, len(data), out) == 0
    return out.raw


def test_keccak256_and_blake3_match_known_answers_and_the_oracle():
    import hashlib
    import blake3
    from oracle import merlin as om
    # Keccak-256 known answers (original padding, not SHA3-256)
    assert om.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert om.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert om.keccak256(b"abc") != hashlib.sha3_256(b"abc").digest()
    # BLAKE3 known answer for the empty input (official test vector)
    assert blake3.blake3(b"").hexdigest() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"
    rnd = random.Random(3)
    for ln in [0, 1, 2, 63, 64, 65, 135, 136, 137, 271, 272, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 3073, 4096, 5000,
               7 * 1024, 8 * 1024 + 1, 20000]:
        data = bytes(rnd.randrange(256) for _ in range(ln))
        assert _host_hash(1, data) == om.keccak256(data), ln
        assert _host_hash(2, data) == blake3.blake3(data).digest(), ln


@pytest.mark.parametrize("name", ["merlin", "keccak256", "blake3"])
def test_verify_with_each_transcript_like_tests_dummy_rs(name):
    """tests/dummy.rs:76-80 runs setup -> prove -> verify with all three transcripts: oracle prover, product verifier."""
    from oracle import polymath as opm, r1cs as orc
    from oracle import merlin as om
    from oracle.rng import StdRng as ORng, fr_rand
    cls = om.TRANSCRIPTS[name]
    rng = ORng.seed_from_u64(31)
    pk = opm.generate_proving_key(orc.DummyCircuit(), rng)
    x, y = fr_rand(rng), fr_rand(rng)
    cs = orc.synthesize(orc.DummyCircuit(x, y), setup_mode=False)
    pr = opm.create_proof_with_assignment(pk, cs.instance_assignment, cs.witness_assignment, rng, transcript_cls=cls)
    pub = [x * y % R_MOD]
    assert opm.verify_proof(pk.vk, pr, pub, transcript_cls=cls)
    vk_bytes, proof = pk.vk.serialize_compressed(), pr.serialize_compressed()
    assert Polymath.verify(vk_bytes, pub, proof, transcript=name) is True
    for other in om.TRANSCRIPTS:
        if other != name:
            assert Polymath.verify(vk_bytes, pub, proof, transcript=other) is False
            assert not opm.verify_proof(pk.vk, pr, pub, transcript_cls=om.TRANSCRIPTS[other])
