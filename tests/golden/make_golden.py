"""Generate the golden fixtures under tests/golden/ from the oracle.

The reference cannot run in this environment and ships no vectors, so these
fixtures freeze the ORACLE's outputs (parity unpinned against real arkworks).  They are consumed by
tests/test_oracle.py (CPU) and tests/test_golden_gpu.py (device path must reproduce them).
Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import polymath as opm, r1cs as orc, rng as orng, poly, curve  # noqa: E402
from oracle.fields import R_MOD  # noqa: E402
import random  # noqa: E402


def dummy(seed):
    rng = orng.StdRng.seed_from_u64(seed)
    pk = opm.generate_proving_key(orc.DummyCircuit(), rng)
    a, b = orng.fr_rand(rng), orng.fr_rand(rng)
    proof = opm.create_proof(orc.DummyCircuit(a, b), pk, rng)
    return dict(circuit="tests/dummy.rs DummyCircuit", seed=seed, a=str(a), b=str(b), public_input=str(a * b % R_MOD),
                vk_hex=pk.vk.serialize_compressed().hex(), proof_hex=proof.serialize_compressed().hex())


def mimc(seed, rounds):
    rng = orng.StdRng.seed_from_u64(seed)
    consts = [orng.fr_rand(rng) for _ in range(rounds)]
    pk = opm.generate_proving_key(orc.MiMCDemo(None, None, consts), rng)
    xl, xr = orng.fr_rand(rng), orng.fr_rand(rng)
    proof = opm.create_proof(orc.MiMCDemo(xl, xr, consts), pk, rng)
    return dict(circuit="tests/mimc.rs MiMCDemo", rounds=rounds, seed=seed, xl=str(xl), xr=str(xr),
                image=str(orc.mimc_hash(xl, xr, consts)), n=pk.vk.n,
                vk_hex=pk.vk.serialize_compressed().hex(), proof_hex=proof.serialize_compressed().hex())


def kernels():
    rnd = random.Random(2024)
    n = 64
    vals = [rnd.randrange(R_MOD) for _ in range(n)]
    dom = poly.Domain(n)
    tbl = curve.FixedBaseTable(curve.G1_GEN, window=8)
    bs = [rnd.randrange(1, R_MOD) for _ in range(48)]
    bases = tbl.mul_many(bs)
    scalars = [rnd.randrange(R_MOD) for _ in range(48)]
    msm = poly.msm_pippenger(scalars, bases)
    return dict(ntt_in=[str(v) for v in vals], ntt_fwd=[str(v) for v in dom.fft(vals)], ntt_inv=[str(v) for v in dom.ifft(vals)],
                base_scalars=[str(v) for v in bs], scalars=[str(v) for v in scalars], msm=[str(msm[0]), str(msm[1])])


if __name__ == "__main__":
    json.dump(dummy(0), open(os.path.join(HERE, "dummy_seed0.json"), "w"), indent=1)
    json.dump(mimc(1, 322), open(os.path.join(HERE, "mimc322_seed1.json"), "w"), indent=1)
    json.dump(mimc(7, 8), open(os.path.join(HERE, "mimc8_seed7.json"), "w"), indent=1)
    json.dump(kernels(), open(os.path.join(HERE, "kernels.json"), "w"), indent=1)
    print("golden fixtures written")
