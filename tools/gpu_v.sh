#!/bin/bash
# 8 GPUs exactly as the driver launches the scaling bench (2^20 + the 2^24 leg), then N=4 without the 2^24 leg
set -x
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2_v_n8.json 2> gpurun_out/bench_r2_v_n8.err
echo rc=$?
grep -v "^\s*$" gpurun_out/bench_r2_v_n8.err | grep -v "OMP_NUM\|\*\*\*\*\|barrier\|return func" | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-2p24 > gpurun_out/bench_r2_v_n4.json 2> gpurun_out/bench_r2_v_n4.err
echo rc=$?
python - <<'PY'
import json
for n in (8,4):
    d=json.loads([l for l in open('gpurun_out/bench_r2_v_n%d.json'%n) if l.startswith('{')][-1])
    print('N=%d'%n, d['value'], d['e2e']['value'], d['phase_ms'], d['proof_check']['matches_golden'], d.get('kernel_sweep'))
    l=d.get('leg_2p24')
    if l: print(' leg24', {k:l[k] for k in l if k not in ('proof_check','parallelism')}, l['proof_check']['verified'], l['proof_check']['proof_hex'][:32])
PY
