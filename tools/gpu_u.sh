#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1
timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 22,24 --iters 3 2>&1 | cut -c60-200
timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 22,24 --iters 3 --skew 2>&1 | cut -c60-200
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_prover_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_u.json 2> gpurun_out/bench_r2_u.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_u.json'));print('mimc',d['value'],d['e2e']['value'],d['phase_ms'],d['proof_check']['matches_golden'],d['roofline']['kernel_ms'],d['roofline']['frac'],d.get('kernel_sweep'),d['cpu_baseline']['value'])"
