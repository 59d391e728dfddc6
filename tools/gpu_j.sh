#!/bin/bash
# one GPU as virtual rank 0 of 8 / of 2 at 2^20: phase times and ncu launch lists
set -x
mkdir -p gpurun_out
for w in 8 2; do
timeout 600 python tools/prove_once.py --log-n 20 --world $w --iters 4 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2_j_w$w.csv \
    python tools/prove_once.py --log-n 20 --world $w --iters 2 > gpurun_out/r2_j_ncu_w$w.log 2>&1
done
