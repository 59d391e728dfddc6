#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 20,22,24 --iters 2 2>&1 | cut -c60-300
timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 20,22,24 --iters 2 --skew 2>&1 | cut -c60-300
timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1
timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 3 --dummy 2>&1 | tail -1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_prover_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5
