"""End-to-end parity of the CUDA setup / prove path (through the C ABI and the C++ host mirror)
against the oracle on the reference's own test circuits.  Bit-exact: proof and key bytes."""
import copy

import pytest

from oracle import polymath as opm
from oracle import r1cs as orc
from oracle.fields import R_MOD
from oracle.rng import StdRng as OStdRng, fr_rand as o_fr_rand
from oracle import curve

pytestmark = pytest.mark.gpu


def _r1cs_from_cs(cs):
    from polymath_b200.api import R1CS
    a, b, c = cs.to_matrices()
    return R1CS(cs.num_instance_variables, cs.num_witness_variables, a, b, c)


def _compare_keys(pk_dev, pk_or):
    assert pk_dev.export_key(0) == pk_or.x_powers_g1
    assert pk_dev.export_key(1) == pk_or.x_powers_y_alpha_g1
    assert pk_dev.export_key(2) == pk_or.x_powers_zh_by_y_alpha_g1
    assert pk_dev.export_key(3) == pk_or.x_powers_y_gamma_g1
    assert pk_dev.export_key(4) == pk_or.x_powers_y_gamma_z_g1
    assert pk_dev.export_key(5) == pk_or.uj_wj_lcs_by_y_alpha_g1
    assert pk_dev.export_key(5, stride=104) == pk_or.uj_wj_lcs_by_y_alpha_g1


def _run_flow(make_circuit_setup, make_circuit_prove, seed, proofs=1, compare_trace=True):
    """Reference flow: seed rng -> setup -> (witness draws) -> prove, on the device and in the oracle."""
    from polymath_b200.api import Polymath, StdRng
    # oracle
    orng = OStdRng.seed_from_u64(seed)
    # device mirror
    drng = StdRng.seed_from_u64(seed)
    setup_circ, extra = make_circuit_setup(orng)
    for _ in range(extra):               # keep the device RNG in step with draws made while building the circuit
        drng.fr_rand()
    cs = orc.synthesize(setup_circ, setup_mode=True)
    pk_or = opm.generate_proving_key(setup_circ, orng)
    pk_dev, vk_bytes = Polymath.setup(_r1cs_from_cs(cs), drng)
    assert vk_bytes == pk_or.vk.serialize_compressed()
    _compare_keys(pk_dev, pk_or)
    for _ in range(proofs):
        circ, draws, public = make_circuit_prove(orng)
        for _ in range(draws):
            drng.fr_rand()
        pcs = orc.synthesize(circ, setup_mode=False)
        trace = {}
        proof_or = opm.create_proof_with_assignment(pk_or, pcs.instance_assignment, pcs.witness_assignment, orng, trace=trace)
        proof_dev = Polymath.prove(pk_dev, pcs.instance_assignment, pcs.witness_assignment, drng)
        if compare_trace:
            n = trace["n"]
            assert pk_dev.debug_read(0) == trace["u"]
            assert pk_dev.debug_read(1) == trace["w"]
            assert pk_dev.debug_read(2) == trace["wu"]
            assert pk_dev.debug_read(3) == trace["u2"]
            assert pk_dev.debug_read(4)[len(pcs.instance_assignment) + len(pcs.witness_assignment):] == trace["y"]
            assert pk_dev.debug_read(6) == trace["d_coeffs"]
        assert proof_dev == proof_or.serialize_compressed()
        assert opm.verify_proof(pk_or.vk, proof_or, public)
        # the host verifier of the library (verifier.rs:19-62) agrees with the oracle's: accept, and reject a wrong input
        assert Polymath.verify(vk_bytes, public, proof_dev)
        assert not Polymath.verify(vk_bytes, [(public[0] + 1) % R_MOD] + list(public[1:]), proof_dev)
        # the RNG streams stay aligned after the proof
        assert drng.next_u64() == orng.next_u64()
    pk_dev.close()


def test_dummy_circuit_flow(pmlib):
    """tests/dummy.rs:37-74 with an explicit seed."""
    state = {}

    def mk_setup(rng):
        return orc.DummyCircuit(), 0

    def mk_prove(rng):
        a, b = o_fr_rand(rng), o_fr_rand(rng)
        state["pub"] = [a * b % R_MOD]
        return orc.DummyCircuit(a, b), 2, state["pub"]

    _run_flow(mk_setup, mk_prove, seed=12345, proofs=2)


@pytest.mark.parametrize("rounds,proofs", [(5, 2), (40, 1), (322, 1)])
def test_mimc_flow(pmlib, rounds, proofs):
    """tests/mimc.rs:145-216 (rounds = 322 is the reference's configuration, n = 2048)."""
    consts = []

    def mk_setup(rng):
        consts[:] = [o_fr_rand(rng) for _ in range(rounds)]
        return orc.MiMCDemo(None, None, consts), rounds

    def mk_prove(rng):
        xl, xr = o_fr_rand(rng), o_fr_rand(rng)
        return orc.MiMCDemo(xl, xr, consts), 2, [orc.mimc_hash(xl, xr, consts)]

    _run_flow(mk_setup, mk_prove, seed=1, proofs=proofs, compare_trace=(rounds <= 40))


def test_bench_dummy_circuit_flow(pmlib):
    """benches/bench.rs:38-61 shape (unused witnesses -> infinity bases, repeated rows, empty last row)."""
    def mk_setup(rng):
        a, b = o_fr_rand(rng), o_fr_rand(rng)
        mk_setup.ab = (a, b)
        return orc.BenchDummyCircuit(a, b, 60, 60), 2

    def mk_prove(rng):
        a, b = mk_setup.ab
        return orc.BenchDummyCircuit(a, b, 60, 60), 0, [a * b % R_MOD]

    _run_flow(mk_setup, mk_prove, seed=0, proofs=1)


def test_unsatisfied_witness_is_rejected(pmlib):
    """prover.rs:108 panics on a non-zero remainder; the C ABI reports PM_ERR_UNSATISFIED."""
    from polymath_b200.api import Polymath, StdRng
    from polymath_b200.lib import PolymathB200Error
    cs = orc.synthesize(orc.DummyCircuit(), setup_mode=True)
    pk, _ = Polymath.setup(_r1cs_from_cs(cs), StdRng.seed_from_u64(3))
    with pytest.raises(PolymathB200Error) as ei:
        Polymath.prove(pk, [1, 7], [2, 3], StdRng.seed_from_u64(4))   # 2*3 != 7
    assert ei.value.code == 3
    # and the context is still usable afterwards
    proof = Polymath.prove(pk, [1, 6], [2, 3], StdRng.seed_from_u64(4))
    assert len(proof) == 176
    # the reference panics (prover.rs:107-108) BEFORE it draws r_a (prover.rs:110): after a refused witness the
    # caller's generator must stand where it stood
    rng, fresh = StdRng.seed_from_u64(4), StdRng.seed_from_u64(4)
    with pytest.raises(PolymathB200Error):
        Polymath.prove(pk, [1, 7], [2, 3], rng)
    assert rng.next_u64() == fresh.next_u64()
    # values outside [0, r) are refused, not reduced
    with pytest.raises(ValueError):
        Polymath.prove(pk, [1, 6 + R_MOD], [2, 3], StdRng.seed_from_u64(4))
    pk.close()


def test_load_host_key_matches_device_setup(pmlib):
    """pm_ctx_create (upload of a host ProvingKey, both 96- and 104-byte strides) proves identically."""
    from polymath_b200.api import Polymath, StdRng, KEY_NAMES
    consts = [3, 5, 7, 11]
    circ = orc.MiMCDemo(None, None, consts)
    cs = orc.synthesize(circ, setup_mode=True)
    pk_or = opm.generate_proving_key(circ, OStdRng.seed_from_u64(9))
    vectors = {name: getattr(pk_or, name) for name in KEY_NAMES}
    pcs = orc.synthesize(orc.MiMCDemo(123, 456, consts), setup_mode=False)
    proofs = []
    for stride in (96, 104):
        pk = Polymath.load_key(_r1cs_from_cs(cs), pk_or.vk.n, pk_or.vk.sigma, vectors, stride=stride)
        proofs.append(Polymath.prove(pk, pcs.instance_assignment, pcs.witness_assignment, StdRng.seed_from_u64(2)))
        pk.close()
    # compressed key vectors (`serialize_compressed` bytes): decoded on the device; exported again they are identical
    comp = {name: b"".join(curve.g1_compress(p) for p in vectors[name]) for name in KEY_NAMES}
    pk = Polymath.load_key(_r1cs_from_cs(cs), pk_or.vk.n, pk_or.vk.sigma, comp, stride=48)
    proofs.append(Polymath.prove(pk, pcs.instance_assignment, pcs.witness_assignment, StdRng.seed_from_u64(2)))
    for which, name in enumerate(KEY_NAMES):
        assert pk.export_key(which, stride=48) == comp[name]
    pk.close()
    want = opm.create_proof_with_assignment(pk_or, pcs.instance_assignment, pcs.witness_assignment, OStdRng.seed_from_u64(2))
    assert proofs[0] == proofs[1] == proofs[2] == want.serialize_compressed()


def test_setup_g2_and_trapdoor_api(pmlib):
    from polymath_b200.api import Polymath
    cs = orc.synthesize(orc.DummyCircuit(), setup_mode=True)
    x, z = 0x1234567890ABCDEF1234567890ABCDEF, R_MOD - 5
    pk, xg2, zg2 = Polymath.setup_with_trapdoors(_r1cs_from_cs(cs), x, z)
    assert xg2 == curve.g2_mul(curve.G2_GEN, x)
    assert zg2 == curve.g2_mul(curve.G2_GEN, z)
    pk.close()


@pytest.mark.parametrize("c", [8, 13])
def test_fixed_base_tables_path(pmlib, c, monkeypatch):
    """The precomputed-level MSM configuration the large circuits use (forced on a small one)."""
    monkeypatch.setenv("PM_MSM_PRECOMP_MIN", "1")
    monkeypatch.setenv("PM_MSM_PRECOMP", str(c))
    consts = []

    def mk_setup(rng):
        consts[:] = [o_fr_rand(rng) for _ in range(20)]
        return orc.MiMCDemo(None, None, consts), 20

    def mk_prove(rng):
        xl, xr = o_fr_rand(rng), o_fr_rand(rng)
        return orc.MiMCDemo(xl, xr, consts), 2, [orc.mimc_hash(xl, xr, consts)]

    _run_flow(mk_setup, mk_prove, seed=31 + c, proofs=2)


@pytest.mark.parametrize("rounds", [2, 4])
def test_prover_with_pair_rounds(pmlib, rounds):
    """The batched-affine accumulation the large circuits use (forced on a small one): proofs stay byte-identical."""
    from polymath_b200 import kernels
    consts = []

    def mk_setup(rng):
        consts[:] = [o_fr_rand(rng) for _ in range(40)]
        return orc.MiMCDemo(None, None, consts), 40

    def mk_prove(rng):
        xl, xr = o_fr_rand(rng), o_fr_rand(rng)
        return orc.MiMCDemo(xl, xr, consts), 2, [orc.mimc_hash(xl, xr, consts)]

    kernels.msm_set_tuning(rounds)
    try:
        _run_flow(mk_setup, mk_prove, seed=77 + rounds, proofs=1)
    finally:
        kernels.msm_set_tuning(-1)


@pytest.mark.parametrize("name", ["keccak256", "blake3"])
def test_dummy_circuit_with_the_hash_transcripts(pmlib, name):
    """tests/dummy.rs:76-80: the same flow with `Keccak256Transcript` / `Blake3Transcript`; the device phases are the same,
    only the host-side challenges differ.  Proof bytes must equal the oracle's and both verifiers must accept."""
    from oracle import merlin as om
    from polymath_b200.api import Polymath, StdRng
    cls = om.TRANSCRIPTS[name]
    seed = 606
    orng, drng = OStdRng.seed_from_u64(seed), StdRng.seed_from_u64(seed)
    cs = orc.synthesize(orc.DummyCircuit(), setup_mode=True)
    pk_or = opm.generate_proving_key(orc.DummyCircuit(), orng)
    pk_dev, vk_bytes = Polymath.setup(_r1cs_from_cs(cs), drng)
    assert vk_bytes == pk_or.vk.serialize_compressed()
    for _ in range(2):
        a, b = o_fr_rand(orng), o_fr_rand(orng)
        drng.fr_rand(); drng.fr_rand()
        pcs = orc.synthesize(orc.DummyCircuit(a, b), setup_mode=False)
        want = opm.create_proof_with_assignment(pk_or, pcs.instance_assignment, pcs.witness_assignment, orng, transcript_cls=cls)
        got = Polymath.prove(pk_dev, pcs.instance_assignment, pcs.witness_assignment, drng, transcript=name)
        assert got == want.serialize_compressed()
        pub = [a * b % R_MOD]
        assert opm.verify_proof(pk_or.vk, want, pub, transcript_cls=cls)
        assert Polymath.verify(vk_bytes, pub, got, transcript=name)
        assert not Polymath.verify(vk_bytes, pub, got, transcript="merlin")
    pk_dev.close()
