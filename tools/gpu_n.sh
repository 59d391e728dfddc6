#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 20,22,24 --iters 2 > gpurun_out/sweep_r2_n_uniform.jsonl 2>&1
timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 20,22,24 --iters 2 --skew > gpurun_out/sweep_r2_n_skew.jsonl 2>&1
cat gpurun_out/sweep_r2_n_uniform.jsonl gpurun_out/sweep_r2_n_skew.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-2p24 > gpurun_out/bench_r2_n_n2.json 2> gpurun_out/bench_r2_n_n2.err
echo rc=$?
grep -v "^\s*$" gpurun_out/bench_r2_n_n2.err | grep -v "OMP_NUM\|\*\*\*\*\|barrier\|return func" | tail -15
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n_n2.json') if l.startswith('{')][-1])
print('N=2', d['value'], d['e2e']['value'], d['phase_ms'], d['proof_check']['matches_golden'], d.get('kernel_sweep'))
PY
