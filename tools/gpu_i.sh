#!/bin/bash
# 8 GPUs: 2^20 bench without the 2^24 leg, per-phase / per-kernel times of rank 0
set -x
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-2p24 > gpurun_out/bench_r2_i_n8.json 2> gpurun_out/bench_r2_i_n8.err
echo rc=$?
grep -v "^\s*$" gpurun_out/bench_r2_i_n8.err | grep -v "OMP_NUM\|\*\*\*\*\|barrier\|return func" | tail -15
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_i_n8.json') if l.startswith('{')][-1])
print('N=8', d['value'], d['e2e']['value'], d['phase_ms'], d['proof_check']['matches_golden'], d.get('kernel_sweep'))
print({k:v for k,v in d.items() if k in ('stage_ms','msm_shapes','bounds','gpu_launches')})
PY
