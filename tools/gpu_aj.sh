#!/bin/bash
# final state, as the driver runs it: N = 8 (with the 2^24 leg), N = 4 and N = 2 (2^24 leg at N = 2 as well)
set -x
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
for n in 8 2 4; do
  extra=""; [ "$n" = "4" ] && extra="--no-2p24"
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 $extra > gpurun_out/bench_r2_aj_n$n.json 2> gpurun_out/bench_r2_aj_n$n.err
  echo "rc=$? n=$n"
  grep -v "^\s*$" gpurun_out/bench_r2_aj_n$n.err | grep -v "OMP_NUM\|\*\*\*\*\|barrier\|return func" | tail -8
done
python - <<'PY'
import json
for n in (2,4,8):
    d=json.loads([l for l in open('gpurun_out/bench_r2_aj_n%d.json'%n) if l.startswith('{')][-1])
    print('N=%d'%n, round(d['value'],2), round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['phase_ms'].items()}, d['proof_check']['matches_golden'], {k:round(v,1) for k,v in d.get('kernel_sweep').items()})
    l=d.get('leg_2p24')
    if l: print(' leg24', round(l['value'],1), round(l['e2e_ms'],1), {k:round(v,1) for k,v in l['phase_ms'].items()}, l['proof_check']['verified'], l['proof_check']['proof_hex'][:16])
PY
