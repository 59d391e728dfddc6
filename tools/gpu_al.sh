#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_al_bench_2p20.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_al_ncu_list.log 2>&1
tail -2 gpurun_out/launches_r2_al_bench_2p20.csv | cut -c1-200
