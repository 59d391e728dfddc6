#!/bin/bash
set -x
mkdir -p gpurun_out
for g in 128 64 32; do
echo "== PM_L2_FETCH=$g"
PM_L2_FETCH=$g timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1
PM_L2_FETCH=$g timeout 600 python tools/sweep.py --skip-basics --ntt 21,24 --msm 22 --iters 3 2>&1 | cut -c1-200
done
