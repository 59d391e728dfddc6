#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sharded_gpu.py tests/test_prover_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python tools/sweep.py --ntt 16,18,20,21,22,23,24,26,28 --msm 16,18,20,22,24,26 --iters 3 --coset --codec 20 > gpurun_out/sweep_r2_ag.jsonl 2>&1
timeout 900 python tools/sweep.py --skip-basics --ntt "" --msm 16,18,20,22,24,26 --iters 3 --skew > gpurun_out/sweep_r2_ag_skew.jsonl 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_ag.json 2> gpurun_out/bench_r2_ag.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_ag.json'));print('mimc',d['value'],d['e2e']['value'],d['phase_ms'],d['proof_check']['matches_golden'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['roofline']['traffic'],d.get('kernel_sweep'),d['cpu_baseline']['value'])"
