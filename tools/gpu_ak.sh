#!/bin/bash
# final sanity of the round: full GPU suite, smoke, bench (N=1) and the reference arm, from a clean rebuild
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_ak_pytest.txt; cat gpurun_out/r2_ak_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_ak.json 2> gpurun_out/bench_r2_ak.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_ak.json'));print('mimc',d['value'],d['e2e']['value'],d['phase_ms'],d['proof_check']['matches_golden'],d['roofline']['frac'],d['cpu_baseline']['value'],d['cpu_baseline']['field_mul'])"
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r2_ak_ref.json 2> gpurun_out/bench_r2_ak_ref.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_ak_ref.json'));print('ref',d['value'],d['cpu_baseline']['cores'],d['cpu_baseline']['proof_matches_golden'],d['cpu_baseline']['field_mul'])"
tail -c 400 gpurun_out/bench_r2_ak_ref.err
