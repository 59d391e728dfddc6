#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sharded_gpu.py -m gpu -x -q -k "ntt or sharded or resident" 2>&1 | tail -3
for b in 0 1; do
echo "== PM_NTT_BIG_TILE=$b"
PM_NTT_BIG_TILE=$b timeout 900 python tools/sweep.py --skip-basics --ntt 21,22,23,24,25 --msm "" --iters 3 2>&1 | cut -c1-150
done
