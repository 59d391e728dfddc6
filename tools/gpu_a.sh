#!/bin/bash
# round-2 capture A: GPU tests, bench line, launch list, --set full of the kernels VERDICT r1 asked for
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_a_pytest.txt
cat gpurun_out/r2_a_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
tail -c 3000 gpurun_out/bench_r2_a.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_a_bench_2p20.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_a_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_digits|k_reduce_level|k_pairs_forward|k_pairs_backward|k_ntt_columns|k_ntt_rows|k_inv_|k_accumulate_rounds|k_sum_slices|k_reduce_top' \
    --launch-skip 0 -c 110 -o gpurun_out/prof_r2_a -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_a_ncu_full.log 2>&1
ncu -i gpurun_out/prof_r2_a.ncu-rep --page raw --csv > gpurun_out/prof_r2_a_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
