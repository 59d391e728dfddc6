#!/bin/bash
# 2 GPUs: the virtual-rank self-test at 2^20 (GPU 0), then the sharded bench incl. the 2^24 leg
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k "resident_kernels" 2>&1 | tail -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_g_n2.json 2> gpurun_out/bench_r2_g_n2.err
tail -c 2500 gpurun_out/bench_r2_g_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_g_n2.json'))
print('N=2', d['value'], d['e2e']['value'], d['phase_ms'], d['proof_check']['matches_golden'], d['kernel_sweep'])
print('leg24', d['leg_2p24'])
PY
