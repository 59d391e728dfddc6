#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_d_pytest.txt
cat gpurun_out/r2_d_pytest.txt
timeout 600 python bench.py --workload dummy --steps 5 --warmup 3 --no-cpu-baseline --write-golden gpurun_out/bench_proofs_dummy.json > gpurun_out/bench_r2_d_dummy.json 2> gpurun_out/bench_r2_d.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_d_dummy.json'));print('dummy',d['value'],d['phase_ms'],d['proof_verified'],d['roofline']['kernel_ms'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_d.json 2>> gpurun_out/bench_r2_d.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_d.json'));print('mimc',d['value'],d['phase_ms'],d['proof_check']['matches_golden'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
for fx in 100000 400000; do PM_MSM_ROUND_FIXED_NS=$fx timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_d_fx$fx.json 2>> gpurun_out/bench_r2_d.err; python -c "import json;d=json.load(open('gpurun_out/bench_r2_d_fx$fx.json'));print('fixed $fx',d['value'],d['phase_ms'],d['roofline']['kernel_ms'],d['roofline']['kernel'][-60:])"; done
tail -c 1500 gpurun_out/bench_r2_d.err
