#!/bin/bash
set -x
mkdir -p gpurun_out
for b in 4 2 3; do
echo "== PM_RED_BITS=$b"
PM_RED_BITS=$b timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -2
PM_RED_BITS=$b timeout 600 python tools/prove_once.py --log-n 20 --world 8 --iters 4 2>&1 | tail -2
done
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_prover_gpu.py -m gpu -x -q 2>&1 | tail -5
