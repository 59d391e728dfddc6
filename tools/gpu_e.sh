#!/bin/bash
set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_e_$name.json 2>> gpurun_out/bench_r2_e.err; python -c "import json;d=json.load(open('gpurun_out/bench_r2_e_$name.json'));print('$name',round(d['value'],2),{k:round(v,2) for k,v in d['phase_ms'].items()},round(d['roofline']['kernel_ms'],2),d['roofline']['kernel'][45:80])"; }
run bias1 PM_MSM_ROUNDS_BIAS=1
run bias2 PM_MSM_ROUNDS_BIAS=2
run p1bias0 PM_P1_ROUNDS_BIAS=0
run p1bias2 PM_P1_ROUNDS_BIAS=2
run ppt32 PM_MSM_PAIRS_PER_THREAD=32
run atab0 PM_A_TABLES=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_e_bench_2p20.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_e_ncu_list.log 2>&1
tail -c 600 gpurun_out/bench_r2_e.err
