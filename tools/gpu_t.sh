#!/bin/bash
# round-2 final capture T: full GPU suite, bench (N=1), launch list of one prove, --set full of the MSM / NTT kernels
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t_pytest.txt
cat gpurun_out/r2_t_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_t.json 2> gpurun_out/bench_r2_t.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_t.json'));print('mimc',d['value'],d['e2e'],d['phase_ms'],d['proof_check']['matches_golden'],d['roofline']['kernel_ms'],d['roofline']['frac'],d.get('kernel_sweep'),d['cpu_baseline']['value'])"
timeout 600 python bench.py --steps 5 --warmup 3 --workload dummy --no-cpu-baseline > gpurun_out/bench_r2_t_dummy.json 2>> gpurun_out/bench_r2_t.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_t_dummy.json'));print('dummy',d['value'],d['phase_ms'],d['proof_verified'])"
timeout 600 python bench.py --steps 5 --warmup 3 --log-n 16 --no-cpu-baseline > gpurun_out/bench_r2_t_2p16.json 2>> gpurun_out/bench_r2_t.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_t_2p16.json'));print('2p16',d['value'],d['e2e']['value'],d['phase_ms'],d['proof_verified'])"
tail -c 800 gpurun_out/bench_r2_t.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_t_bench_2p20.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_t_ncu_list.log 2>&1
PM_CUDA_PROFILER=phase3 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_digits|k_reduce_level|k_pairs_forward|k_pairs_backward|k_invert_up|k_invert_down|k_accumulate_rounds|k_sum_slices|k_reduce_top' \
    -c 16 -o /tmp/prof_r2_t_phase3 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_t_ncu_phase3.log 2>&1
ncu -i /tmp/prof_r2_t_phase3.ncu-rep --page raw --csv > gpurun_out/prof_r2_t_phase3_raw.csv 2>/dev/null
PM_CUDA_PROFILER=phase1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_ntt_columns|k_ntt_rows|k_sap_constraint_rows' -c 5 -o /tmp/prof_r2_t_ntt -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_t_ncu_ntt.log 2>&1
ncu -i /tmp/prof_r2_t_ntt.ncu-rep --page raw --csv > gpurun_out/prof_r2_t_ntt_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
