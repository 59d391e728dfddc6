#!/bin/bash
# compute-sanitizer (memcheck, then racecheck + synccheck on the shared-memory kernels) over small proves and MSM / NTT cases
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, random
sys.path.insert(0, '.')
from polymath_b200 import circuits, kernels, codec
from polymath_b200.api import Polymath, StdRng
from oracle.fields import R_MOD
for log_n in (8, 11, 13):
    r1cs, inst, wit, rng = circuits.synthetic_mimc(1 << log_n, seed=5)
    pk, vk = Polymath.setup(r1cs, rng)
    proof = Polymath.prove(pk, inst, wit, rng)
    assert Polymath.verify(vk, inst[1:], proof)
    pk.close()
    print("prove ok", log_n, flush=True)
rnd = random.Random(3)
for lg in (5, 9, 12, 13):
    v = [rnd.randrange(R_MOD) for _ in range(1 << lg)]
    w = kernels.ntt_fr(kernels.ntt_fr(v), inverse=True)
    assert w == v
    print("ntt ok", lg, flush=True)
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/r2_ab_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2_ab_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/r2_ab_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/r2_ab_racecheck.log
