"""The device path must reproduce the committed golden fixtures (tests/golden/*.json, frozen oracle
outputs) byte for byte, and proofs of circuits far beyond the oracle's reach must satisfy the
reference's own acceptance property: `verify` accepts (tests/mimc.rs:214)."""
import json
import os

import pytest

from oracle import polymath as opm, r1cs as orc, curve
from oracle.fields import R_MOD
from oracle.merlin import MerlinFieldTranscript

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _r1cs_from_cs(cs):
    from polymath_b200.api import R1CS
    a, b, c = cs.to_matrices()
    return R1CS(cs.num_instance_variables, cs.num_witness_variables, a, b, c)


def test_golden_dummy(pmlib):
    from polymath_b200.api import Polymath, StdRng
    g = json.load(open(os.path.join(GOLDEN, "dummy_seed0.json")))
    rng = StdRng.seed_from_u64(g["seed"])
    cs = orc.synthesize(orc.DummyCircuit(), setup_mode=True)
    pk, vk = Polymath.setup(_r1cs_from_cs(cs), rng)
    assert vk.hex() == g["vk_hex"]
    a, b = rng.fr_rand(), rng.fr_rand()
    assert (a, b) == (int(g["a"]), int(g["b"]))
    proof = Polymath.prove(pk, [1, a * b % R_MOD], [a, b], rng)
    assert proof.hex() == g["proof_hex"]
    pk.close()


@pytest.mark.parametrize("name", ["mimc8_seed7.json", "mimc322_seed1.json"])
def test_golden_mimc(pmlib, name):
    from polymath_b200.api import Polymath, StdRng
    from polymath_b200 import circuits
    g = json.load(open(os.path.join(GOLDEN, name)))
    rng = StdRng.seed_from_u64(g["seed"])
    consts = [rng.fr_rand() for _ in range(g["rounds"])]
    pk, vk = Polymath.setup(circuits.mimc_r1cs(consts), rng)
    assert vk.hex() == g["vk_hex"] and pk.n == g["n"]
    xl, xr = rng.fr_rand(), rng.fr_rand()
    assert (xl, xr) == (int(g["xl"]), int(g["xr"]))
    inst, wit = circuits.mimc_assignment(xl, xr, consts)
    assert inst[1] == int(g["image"])
    assert Polymath.prove(pk, inst, wit, rng).hex() == g["proof_hex"]
    pk.close()


def test_golden_kernels(pmlib):
    from polymath_b200 import kernels
    g = json.load(open(os.path.join(GOLDEN, "kernels.json")))
    vals = [int(v) for v in g["ntt_in"]]
    assert kernels.ntt_fr(vals) == [int(v) for v in g["ntt_fwd"]]
    assert kernels.ntt_fr(vals, inverse=True) == [int(v) for v in g["ntt_inv"]]
    bases = kernels.fixed_base_mul([int(v) for v in g["base_scalars"]])
    got = kernels.msm_g1(bases, [int(v) for v in g["scalars"]])
    assert got == (int(g["msm"][0]), int(g["msm"][1]))


class _Challenges:
    """Host side of the phase API for a test: the reference's Fiat-Shamir algebra via the oracle helpers."""

    def __init__(self, vk, instance):
        self.vk, self.instance = vk, instance
        self.t = MerlinFieldTranscript(opm.B_POLYMATH)

    def first(self, a_pt, c_pt):
        self.x1 = opm.compute_x1(self.t, self.instance, [a_pt, c_pt])
        y1 = opm.compute_y1(self.x1, self.vk.sigma)
        self.y1_alpha = opm.neg_power(y1, opm.MINUS_ALPHA)
        self.y1_gamma = opm.neg_power(y1, opm.MINUS_GAMMA)
        return self.x1, self.y1_alpha

    def second(self, a_at_x1):
        pi = opm.compute_pi_at_x1(self.vk, self.instance, self.x1, self.y1_gamma)
        c_at_x1 = opm.compute_c_at_x1(self.y1_gamma, self.y1_alpha, a_at_x1, pi)
        x2 = opm.compute_x2(self.t, self.x1, [a_at_x1, c_at_x1])
        return x2, c_at_x1


_LARGE = [int(v) for v in os.environ.get("PM_TEST_LARGE", "").split(",") if v]   # e.g. PM_TEST_LARGE=24 (minutes)


@pytest.mark.parametrize("log_n", [14, 17, 22] + _LARGE)
def test_large_synthetic_circuit_verifies(pmlib, log_n):
    """S-mimc(2^log_n): setup + prove entirely on the device, then the oracle's pairing
    check must accept — the property the reference's own tests assert.  2^17 exercises the fixed-base tables,
    2^22 the first domain whose squaring transform (2^23) takes three NTT passes and the size from which a sharded
    prove runs its transforms through the all-to-all; 2^20, BASELINE.json's headline size, is covered byte for byte by
    test_bench_check_proof_matches_golden_and_the_oracle_accepts below."""
    from polymath_b200 import circuits
    from polymath_b200.api import Polymath
    n = 1 << log_n
    r1cs, inst, wit, rng = circuits.synthetic_mimc(n, seed=5)
    x, z = rng.fr_rand(), rng.fr_rand()
    pk, x_g2, z_g2 = Polymath.setup_with_trapdoors(r1cs, x, z)
    assert pk.n == n and pk.sigma == n + 3
    vk = opm.VerifyingKey(one_g1=curve.G1_GEN, one_g2=curve.G2_GEN, x_g2=x_g2, z_g2=z_g2, n=n, m0=2, sigma=n + 3,
                          omega=__import__("oracle.poly", fromlist=["Domain"]).Domain(n).group_gen)
    ra = [rng.fr_rand(), rng.fr_rand()]
    a, c, a_at_x1, d = Polymath.prove_phases(pk, inst, wit, ra, _Challenges(vk, inst))
    proof = opm.Proof(a_g1=a, c_g1=c, a_at_x1=a_at_x1, d_g1=d)
    assert opm.verify_proof(vk, proof, inst[1:])
    assert not opm.verify_proof(vk, proof, [(inst[1] + 1) % R_MOD])
    # a second proof of the same statement with other blinding: different bytes, accepted as well (zero-knowledge
    # blinding actually enters the commitments), and the library's host verifier agrees with the oracle's on both
    ra2 = [rng.fr_rand(), rng.fr_rand()]
    a2, c2, a2_at_x1, d2 = Polymath.prove_phases(pk, inst, wit, ra2, _Challenges(vk, inst))
    proof2 = opm.Proof(a_g1=a2, c_g1=c2, a_at_x1=a2_at_x1, d_g1=d2)
    assert (a2, c2, d2) != (a, c, d)
    assert opm.verify_proof(vk, proof2, inst[1:])
    vk_bytes = vk.serialize_compressed()
    assert Polymath.verify(vk_bytes, inst[1:], proof.serialize_compressed())
    assert Polymath.verify(vk_bytes, inst[1:], proof2.serialize_compressed())
    pk.close()


def test_bench_check_proof_matches_golden_and_the_oracle_accepts(pmlib):
    """The proof bench.py prints as `proof_check` (S-mimc(2^20), setup seed 1, blinding seed CHECK_SEED) is pinned in
    tests/golden/bench_proofs.json: the one-GPU path must reproduce it byte for byte here, and the ORACLE's pairing
    check must accept those bytes — so a bench run on N GPUs that matches the golden has matched an oracle-accepted,
    unsharded proof of the same seed (VERDICT r1, next #1b)."""
    import bench
    from polymath_b200 import keydump
    from polymath_b200.api import Polymath, StdRng
    golden = json.load(open(os.path.join(GOLDEN, "bench_proofs.json")))["mimc_2p20_seed1_check"]
    log_n = 20
    n = 1 << log_n
    r1cs, inst, wit, rng = keydump.build_workload("mimc", log_n, 1)
    x, z = rng.fr_rand(), rng.fr_rand()          # sample_element_outside_domain x2 (generator.rs:72,77)
    assert pow(x, n, R_MOD) != 1 and pow(z, n, R_MOD) != 1
    pk, x_g2, z_g2 = Polymath.setup_with_trapdoors(r1cs, x, z)
    vk = opm.VerifyingKey(one_g1=curve.G1_GEN, one_g2=curve.G2_GEN, x_g2=x_g2, z_g2=z_g2, n=n, m0=2, sigma=n + 3,
                          omega=__import__("oracle.poly", fromlist=["Domain"]).Domain(n).group_gen)
    got = Polymath.prove(pk, inst, wit, StdRng.seed_from_u64(bench.CHECK_SEED))
    assert got.hex() == golden
    proof = opm.Proof(a_g1=curve.g1_decompress(got[:48]), c_g1=curve.g1_decompress(got[48:96]),
                      a_at_x1=int.from_bytes(got[96:128], "little"), d_g1=curve.g1_decompress(got[128:]))
    assert opm.verify_proof(vk, proof, inst[1:])
    assert Polymath.verify(vk.serialize_compressed(), inst[1:], got)
    pk.close()


def test_bench_dummy_shape_at_scale(pmlib):
    """benches/bench.rs circuit shape with 2^13 variables: thousands of infinity bases, all constraint rows
    equal (every y-scalar identical -> one giant bucket per window), empty last row; proof must verify."""
    from polymath_b200 import circuits
    from polymath_b200.api import Polymath, StdRng
    rng = StdRng.seed_from_u64(0)
    a, b = rng.fr_rand(), rng.fr_rand()
    nv = nc = (1 << 13) - 100
    r1cs, inst, wit = circuits.bench_dummy(nv, nc, a, b)
    x, z = rng.fr_rand(), rng.fr_rand()
    pk, x_g2, z_g2 = Polymath.setup_with_trapdoors(r1cs, x, z)
    n = pk.n
    lcs = pk.export_key(5)
    assert sum(1 for p in lcs if p is None) >= nv - 10        # unused witness columns -> [0]G
    vk = opm.VerifyingKey(one_g1=curve.G1_GEN, one_g2=curve.G2_GEN, x_g2=x_g2, z_g2=z_g2, n=n, m0=2, sigma=n + 3,
                          omega=__import__("oracle.poly", fromlist=["Domain"]).Domain(n).group_gen)
    ra = [rng.fr_rand(), rng.fr_rand()]
    pa, pc, a_at_x1, pd = Polymath.prove_phases(pk, inst, wit, ra, _Challenges(vk, inst))
    assert opm.verify_proof(vk, opm.Proof(a_g1=pa, c_g1=pc, a_at_x1=a_at_x1, d_g1=pd), inst[1:])
    pk.close()
