"""Run setup + a few proves of S-mimc(2^log_n) on one GPU, optionally as virtual rank 0 of `world`
(partials only), for profiling with ncu.  Prints phase times."""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polymath_b200 import circuits, codec, sharded  # noqa: E402
from polymath_b200.api import StdRng, _lib  # noqa: E402
from polymath_b200.lib import check  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--dummy", action="store_true", help="S-dummy (benches/bench.rs shape) instead of S-mimc")
    a = ap.parse_args()
    lib = _lib()
    sharded.bind(lib)
    if a.dummy:
        rng = StdRng.seed_from_u64(0)
        av, bv = rng.fr_rand(), rng.fr_rand()
        nv = nc = (1 << (a.log_n - 1)) - 2          # SAP rows = 2 (m0 + n_r) = 2^log_n
        r1cs, inst, wit = circuits.bench_dummy(nv, nc, av, bv)
    else:
        r1cs, inst, wit, rng = circuits.synthetic_mimc(1 << a.log_n, seed=1)
    x, z = rng.fr_rand(), rng.fr_rand()
    h = C.c_void_p()
    xg2, zg2 = C.create_string_buffer(192), C.create_string_buffer(192)
    t0 = time.perf_counter()
    check(lib.pm_setup_sharded(C.byref(r1cs.view), codec.fr_to_wire(x), codec.fr_to_wire(z), 0, a.world, C.byref(h), xg2, zg2))
    print("setup %.3f s" % (time.perf_counter() - t0))
    check(lib.pm_ctx_set_assignment(h, codec.frs_to_wire(inst), codec.frs_to_wire(wit)))
    ra = codec.frs_to_wire([rng.fr_rand(), rng.fr_rand()])
    part1, part3 = C.create_string_buffer(384), C.create_string_buffer(192)
    ao, co, ev, do = (C.create_string_buffer(96), C.create_string_buffer(96), C.create_string_buffer(32), C.create_string_buffer(96))
    for it in range(a.iters):
        t0 = time.perf_counter()
        check(lib.pm_prove_phase1_partial(h, ra, part1))
        t1 = time.perf_counter()
        check(lib.pm_prove_phase1_finish(h, part1.raw, 1, ao, co))
        check(lib.pm_prove_phase2(h, codec.fr_to_wire(12345), codec.fr_to_wire(6789), ev))
        t2 = time.perf_counter()
        # arbitrary challenges: the opening remainder is then non-zero, which phase 3 reports after doing all the work
        rc = lib.pm_prove_phase3_partial(h, codec.fr_to_wire(1111), codec.fr_to_wire(2222), part3)
        t3 = time.perf_counter()
        ms = (C.c_double * 3)()
        lib.pm_ctx_phase_ms(h, ms)
        print("iter %d: wall p1 %.2f ms, finish+p2 %.2f ms, p3 %.2f ms (rc %d); device p1 %.2f p2 %.2f p3 %.2f" %
              (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, rc, ms[0], ms[1], ms[2]))


if __name__ == "__main__":
    main()
