#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_sharded_gpu.py -m gpu -x -q -k "msm_edge or msm_pair_rounds or msm_skewed or msm_bucket_set or exceptional or resident_kernels_with_virtual_ranks-10 or resident_kernels_with_virtual_ranks-14 or precomputed or ntt_matches or ntt_coset or fixed_base or g1_compression" > gpurun_out/r2_ac_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/r2_ac_memcheck.log
cat > /tmp/san2.py <<'PY'
import sys
sys.path.insert(0, '.')
from polymath_b200 import circuits, kernels
from polymath_b200.api import Polymath, StdRng
for log_n, rounds in ((16, -1), (12, 2)):
    kernels.msm_set_tuning(rounds)
    r1cs, inst, wit, rng = circuits.synthetic_mimc(1 << log_n, seed=5)
    pk, vk = Polymath.setup(r1cs, rng)
    proof = Polymath.prove(pk, inst, wit, rng)
    assert Polymath.verify(vk, inst[1:], proof)
    pk.close()
    print("prove ok", log_n, rounds, flush=True)
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/r2_ac_memcheck2.log 2>&1; echo "memcheck2 rc=$?"
tail -4 gpurun_out/r2_ac_memcheck2.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/r2_ac_racecheck2.log 2>&1; echo "racecheck2 rc=$?"
tail -4 gpurun_out/r2_ac_racecheck2.log
