"""oracle/fast.py (the CPU prove that bench.py times: every O(n) step in oracle/cpu_ref.cpp) must return the very
bytes of oracle/polymath.py's line-by-line restatement of prover.rs, on the reference's own circuits."""
import numpy as np
import pytest

from oracle import fast, polymath as opm, r1cs as orc
from oracle.fields import R_MOD
from oracle.poly import Domain
from oracle.rng import StdRng as ORng, fr_rand


def _flow(circ_setup, circ_prove, seed):
    rng = ORng.seed_from_u64(seed)
    pk = opm.generate_proving_key(circ_setup, rng)
    cs = orc.synthesize(circ_prove, setup_mode=False)
    inst, wit = cs.instance_assignment, cs.witness_assignment
    trace = {}
    want = opm.create_proof_with_assignment(pk, inst, wit, rng, trace=trace)
    sap = pk.sap_matrices
    m0, m, nr = sap.m0_m_n()
    n = pk.vk.n
    tm = {}
    got = fast.prove(fast.key_from_oracle(pk), fast.Csr.from_rows(sap.a), fast.Csr.from_rows(sap.b), fast.Csr.from_rows(sap.c),
                     m0, m - m0, nr, n, pk.vk.sigma, pk.vk.omega, fast.fr_wire(inst), fast.fr_wire(wit), trace["ra"], timings=tm)
    assert got.serialize_compressed() == want.serialize_compressed()
    assert opm.verify_proof(pk.vk, got, inst[1:])
    assert set(tm) == {"sap", "ntt", "msm_phase1", "opening", "msm_d"}
    return pk, inst, wit, trace


def test_fast_prove_matches_the_restatement_dummy():
    _flow(orc.DummyCircuit(), orc.DummyCircuit(5, 7), seed=11)             # tests/dummy.rs: n = 8


def test_fast_prove_matches_the_restatement_mimc():
    consts = [3, 1 << 200, 7, 0, 11, 13, 17, 19]                          # a zero constant: the `lin` row has one entry
    pk, inst, wit, trace = _flow(orc.MiMCDemo(None, None, consts), orc.MiMCDemo(21, 34, consts), seed=12)
    # an unsatisfying witness trips the same assert as prover.rs:108
    sap = pk.sap_matrices
    m0, m, nr = sap.m0_m_n()
    bad = list(wit)
    bad[3] = (bad[3] + 1) % R_MOD
    with pytest.raises(AssertionError, match="prover.rs:108"):
        fast.prove(fast.key_from_oracle(pk), fast.Csr.from_rows(sap.a), fast.Csr.from_rows(sap.b), fast.Csr.from_rows(sap.c),
                   m0, m - m0, nr, pk.vk.n, pk.vk.sigma, pk.vk.omega, fast.fr_wire(inst), fast.fr_wire(bad), trace["ra"])


def test_fast_prove_bench_dummy_shape():
    """benches/bench.rs:38-61: unused witnesses (infinity bases), identical rows, empty last row."""
    a, b = 123456789, 987654321
    _flow(orc.BenchDummyCircuit(a, b, 40, 40), orc.BenchDummyCircuit(a, b, 40, 40), seed=13)


def test_chunked_division_and_horner_match_the_sequential_recurrence():
    import ctypes as C
    import random
    rnd = random.Random(5)
    lib = fast._lib()
    for ln in (2, 3, 17, 1000, 4097):
        num = [rnd.randrange(R_MOD) for _ in range(ln)]
        x1 = rnd.randrange(R_MOD)
        q_want, rem_want = opm.divide_by_linear(num, x1)
        nw, xw = fast.fr_wire(num), fast.fr_wire([x1])
        q = np.zeros((ln - 1, 4), np.uint64)
        rem = np.zeros(4, np.uint64)
        assert lib.orc_divide_linear(fast._p(nw), ln, fast._p(xw), fast._p(q), fast._p(rem)) == 0
        assert [fast.fr_int(r) for r in q] == q_want and fast.fr_int(rem) == rem_want
        ev = np.zeros(4, np.uint64)
        lib.orc_horner(fast._p(nw), ln, fast._p(xw), fast._p(ev))
        assert fast.fr_int(ev) == sum(cf * pow(x1, i, R_MOD) for i, cf in enumerate(num)) % R_MOD
