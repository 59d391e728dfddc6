#!/bin/bash
# round-2 capture B: launch list of one prove + --set full of the [d]_1 MSM kernels (phase 3) and the NTT (phase 1).
# The .ncu-rep files stay in /tmp on the box (hundreds of MB); only the raw CSV pages come back.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_b_bench_2p20.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_b_ncu_list.log 2>&1
PM_CUDA_PROFILER=phase3 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_digits|k_reduce_level|k_pairs_forward|k_pairs_backward|k_inv_top|k_inv_down|k_accumulate_rounds|k_sum_slices|k_reduce_top' \
    -c 14 -o /tmp/prof_r2_b_phase3 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_b_ncu_phase3.log 2>&1
ncu -i /tmp/prof_r2_b_phase3.ncu-rep --page raw --csv > gpurun_out/prof_r2_b_phase3_raw.csv 2>/dev/null
PM_CUDA_PROFILER=phase1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_ntt_columns|k_ntt_rows' -c 4 -o /tmp/prof_r2_b_ntt -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_b_ncu_ntt.log 2>&1
ncu -i /tmp/prof_r2_b_ntt.ncu-rep --page raw --csv > gpurun_out/prof_r2_b_ntt_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/*.ncu-rep | tail -12
