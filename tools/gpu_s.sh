#!/bin/bash
set -x
timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1
timeout 600 python tools/prove_once.py --log-n 20 --world 8 --iters 4 2>&1 | tail -1
timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 22 --iters 3 --codec 20 2>&1 | cut -c1-200
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_prover_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5
