#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_h_pytest.txt
cat gpurun_out/r2_h_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_h.json 2> gpurun_out/bench_r2_h.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_h.json'));print('mimc',d['value'],d['phase_ms'],d['proof_check']['matches_golden'],d['roofline']['kernel_ms'],d['roofline']['frac'],d.get('kernel_sweep'),d.get('cpu_baseline'))"
tail -c 1500 gpurun_out/bench_r2_h.err
