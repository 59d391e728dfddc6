"""Synthetic SAP workloads of SURVEY.md §8(d) for bench.py and the large-size tests.

`S-mimc(n)`: the MiMC Feistel chain of /root/reference/tests/mimc.rs:74-143 with
ROUNDS = n/4 - 1, so that the SAP has exactly n rows (m0 = 2, n_r = 2*ROUNDS, mw = 2*ROUNDS + 1).
`S-dummy(n)`: /root/reference/benches/bench.rs:38-61 (a*b = c repeated, unused witness copies).
These build the R1CS matrices exactly as ark-relations' `to_matrices()` lays them out
(column 0 = one, 1 = the public image, 2.. = witnesses; rows sorted by variable).
Input generation only — no proving logic here.
"""
from .api import R1CS, StdRng
from .codec import R_MOD


def mimc_rounds_for_domain(n: int) -> int:
    assert n >= 8 and n & (n - 1) == 0
    return n // 4 - 1


def mimc_r1cs(constants):
    """R1CS of MiMCDemo (tests/mimc.rs:74-143) for len(constants) rounds."""
    rounds = len(constants)
    m0 = 2
    a_rows, b_rows, c_rows = [], [], []
    xl_col, xr_col = m0 + 0, m0 + 1
    next_w = 2
    for i, ci in enumerate(constants):
        tmp_col = m0 + next_w
        next_w += 1
        lin = [(ci, 0), (1, xl_col)] if ci else [(1, xl_col)]
        a_rows.append(lin)
        b_rows.append(lin)
        c_rows.append([(1, tmp_col)])
        if i == rounds - 1:
            new_col = 1
            c2 = [(1, new_col), (R_MOD - 1, xr_col)]
        else:
            new_col = m0 + next_w
            next_w += 1
            c2 = [(R_MOD - 1, xr_col), (1, new_col)]
        a_rows.append([(1, tmp_col)])
        b_rows.append(lin)
        c_rows.append(c2)
        xr_col, xl_col = xl_col, new_col
    return R1CS(m0, next_w, a_rows, b_rows, c_rows)


def mimc_assignment(xl, xr, constants):
    """(instance, witness) of MiMCDemo for the preimage (xl, xr): instance = [1, image]."""
    wit = [xl % R_MOD, xr % R_MOD]
    rounds = len(constants)
    image = None
    for i, ci in enumerate(constants):
        t = (xl + ci) % R_MOD
        tmp = t * t % R_MOD
        new_xl = (t * tmp + xr) % R_MOD
        wit.append(tmp)
        if i == rounds - 1:
            image = new_xl
        else:
            wit.append(new_xl)
        xl, xr = new_xl, xl
    return [1, image], wit


def synthetic_mimc(n: int, seed: int = 1, rng=None):
    """S-mimc(n): constants and preimage from StdRng::seed_from_u64(seed).  `rng`: any object with
    `fr_rand()` positioned at that seed (the CPU reference arm passes the oracle's generator, so that its process
    never loads the CUDA library)."""
    rng = rng or StdRng.seed_from_u64(seed)
    rounds = mimc_rounds_for_domain(n)
    constants = [rng.fr_rand() for _ in range(rounds)]
    r1cs = mimc_r1cs(constants)
    xl, xr = rng.fr_rand(), rng.fr_rand()
    instance, witness = mimc_assignment(xl, xr, constants)
    return r1cs, instance, witness, rng


def bench_dummy(num_variables: int, num_constraints: int, a: int, b: int):
    """S-dummy: benches/bench.rs:38-61."""
    m0 = 2
    a_col, b_col, c_col = m0 + 0, m0 + 1, 1
    a_rows = [[(1, a_col)]] * (num_constraints - 1) + [[]]
    b_rows = [[(1, b_col)]] * (num_constraints - 1) + [[]]
    c_rows = [[(1, c_col)]] * (num_constraints - 1) + [[]]
    r1cs = R1CS(m0, num_variables - 1, a_rows, b_rows, c_rows)
    instance = [1, a * b % R_MOD]
    witness = [a, b] + [a] * (num_variables - 3)
    return r1cs, instance, witness
