#!/bin/bash
# DRAM traffic of the whole [d]_1 bucket-accumulation stage (all rounds, both halves) and of one Fr NTT 2^21 -> profiles/r2_traffic.json
set -x
mkdir -p gpurun_out
PM_CUDA_PROFILER=phase3 timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k regex:'k_pairs_forward|k_pairs_backward|k_invert_up|k_invert_top|k_invert_down|k_accumulate_rounds|k_accumulate_heavy|k_heavy_finish' \
    -c 120 --csv --log-file gpurun_out/traffic_r2_y_stage.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_y_stage.log 2>&1
cat > /tmp/ntt21.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
from polymath_b200.lib import require_device, check
lib = require_device(); d = C.c_double()
check(lib.pm_bench_ntt(21, 0, 1, C.byref(d)))
PY
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_columns|k_ntt_rows' -c 8 --csv --log-file gpurun_out/traffic_r2_y_ntt.csv python /tmp/ntt21.py > gpurun_out/r2_y_ntt.log 2>&1
tail -3 gpurun_out/traffic_r2_y_stage.csv; tail -5 gpurun_out/traffic_r2_y_ntt.csv
