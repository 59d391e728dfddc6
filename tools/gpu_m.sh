#!/bin/bash
set -x
mkdir -p gpurun_out
for lg in 11 14 16; do
timeout 600 python tools/prove_once.py --log-n $lg --world 1 --iters 6 2>&1 | tail -2
done
timeout 600 python tools/prove_once.py --log-n 20 --world 8 --iters 4 2>&1 | tail -1
timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_prover_gpu.py tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -5
