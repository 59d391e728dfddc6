#!/bin/bash
# small circuits: wall / device phase times and the launch list at 2^11, 2^14
set -x
mkdir -p gpurun_out
for lg in 11 14; do
timeout 600 python tools/prove_once.py --log-n $lg --world 1 --iters 6 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2_l_2p$lg.csv \
    python tools/prove_once.py --log-n $lg --world 1 --iters 2 > gpurun_out/r2_l_ncu_2p$lg.log 2>&1
done
