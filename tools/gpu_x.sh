#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python tools/sweep.py --ntt 16,18,20,21,22,24,26,28 --msm 16,18,20,22,24,26 --iters 3 --codec 20 > gpurun_out/sweep_r2_x.jsonl 2>&1
timeout 900 python tools/sweep.py --skip-basics --ntt "" --msm 16,18,20,22,24,26 --iters 3 --skew > gpurun_out/sweep_r2_x_skew.jsonl 2>&1
cat gpurun_out/sweep_r2_x.jsonl gpurun_out/sweep_r2_x_skew.jsonl | cut -c1-230
