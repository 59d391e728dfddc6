#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/sweep.py --skip-basics --ntt 20,21,22,23,24,26,28 --msm "" --iters 3 --coset 2>&1 | cut -c1-175
