#!/bin/bash
# round-2 capture C: new bench line (golden proof_check written), GPU tests, reference arm, S-dummy workload, halves A/B
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --write-golden tests/golden/bench_proofs.json > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err
tail -c 1500 gpurun_out/bench_r2_c.err
cp tests/golden/bench_proofs.json gpurun_out/bench_proofs.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_c.json'))
print({k:d[k] for k in ('value','phase_ms','setup_s','kernel_sweep','proof_verified')}, d['e2e'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['cpu_baseline'] and {k:d['cpu_baseline'][k] for k in ('value','cores','proof_matches_device','split_ms')}, d['proof_check']['matches_golden'], d['setup'])
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_c_pytest.txt
cat gpurun_out/r2_c_pytest.txt
PM_MSM_HALVES=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c_nohalves.json 2>> gpurun_out/bench_r2_c.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_c_nohalves.json'));print('nohalves',d['value'],d['phase_ms'],d['roofline']['kernel_ms'])"
timeout 600 python bench.py --workload dummy --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c_dummy.json 2>> gpurun_out/bench_r2_c.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_c_dummy.json'));print('dummy',d['value'],d['phase_ms'],d['proof_verified'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_c_ref.json 2>> gpurun_out/bench_r2_c.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2_c_ref.json'));print('ref',d['value'],d['cpu_baseline']['cores'],d['cpu_baseline']['proof_matches_golden'],d['cpu_baseline']['split_ms_last'])"
tail -c 1000 gpurun_out/bench_r2_c.err
