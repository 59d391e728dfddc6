#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_f_pytest.txt
cat gpurun_out/r2_f_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_f.json'));print('mimc',d['value'],d['phase_ms'],d['proof_check']['matches_golden'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
tail -c 1500 gpurun_out/bench_r2_f.err
