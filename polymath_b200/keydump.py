"""Export a device-built proving key to files, for consumers that must not load this library.

    python -m polymath_b200.keydump --log-n 20 --seed 1 --out DIR [--workload mimc|dummy]

Runs `Polymath::setup` for the S-mimc / S-dummy circuit of bench.py on the current GPU and writes the six G1
vectors of the `ProvingKey` (src/data_structures.rs:56-73) as raw 96-byte Montgomery affine records
(`<name>.bin`, (0,0) = infinity), plus `vk.bin` (compressed VerifyingKey).  `bench.py --impl reference` spawns
this in a SUBPROCESS: the CPU arm then proves with a real key (its proof is checked against the device's) while
its own process never maps libpolymath_b200.so.
"""
import argparse
import ctypes as C
import json
import os
import sys

KEY_NAMES = ["x_powers_g1", "x_powers_y_alpha_g1", "x_powers_zh_by_y_alpha_g1", "x_powers_y_gamma_g1",
             "x_powers_y_gamma_z_g1", "uj_wj_lcs_by_y_alpha_g1"]


def export_raw(pk, which):
    """Vector `which` of a ProvingKey handle as bytes (96 B per point), without Python-level decoding."""
    from .lib import check
    ln = pk.key_len(which)
    buf = C.create_string_buffer(max(ln * 96, 1))
    check(pk._lib.pm_ctx_export_key(pk._h, which, buf, 96))
    return buf, ln


def build_workload(workload, log_n, seed, rng=None):
    """(r1cs, instance, witness, rng) of bench.py's workloads; `rng`: see circuits.synthetic_mimc."""
    from . import circuits
    n = 1 << log_n
    if workload == "dummy":
        if rng is None:
            from .api import StdRng
            rng = StdRng.seed_from_u64(seed)
        a, b = rng.fr_rand(), rng.fr_rand()          # benches/bench.rs:65-68
        nv = nc = n // 2 - 2                         # SAP rows 2 (m0 + n_r) = n exactly
        r1cs, inst, wit = circuits.bench_dummy(nv, nc, a, b)
        return r1cs, inst, wit, rng
    return circuits.synthetic_mimc(n, seed=seed, rng=rng)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--workload", default="mimc", choices=["mimc", "dummy"])
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    from .api import Polymath
    from .lib import require_device
    require_device()
    r1cs, inst, wit, rng = build_workload(a.workload, a.log_n, a.seed)
    pk, vk = Polymath.setup(r1cs, rng)
    os.makedirs(a.out, exist_ok=True)
    lens = {}
    for i, name in enumerate(KEY_NAMES):
        buf, ln = export_raw(pk, i)
        with open(os.path.join(a.out, name + ".bin"), "wb") as fh:
            fh.write(memoryview(buf)[:ln * 96])
        lens[name] = ln
    with open(os.path.join(a.out, "vk.bin"), "wb") as fh:
        fh.write(vk)
    with open(os.path.join(a.out, "meta.json"), "w") as fh:
        json.dump({"log_n": a.log_n, "seed": a.seed, "workload": a.workload, "n": pk.n, "sigma": pk.sigma, "lens": lens}, fh)
    pk.close()
    print("key written to", a.out)


if __name__ == "__main__":
    sys.exit(main())
