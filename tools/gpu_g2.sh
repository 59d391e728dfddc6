#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
run() { name=$1; shift; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-2p24 > gpurun_out/bench_r2_g2_$name.json 2> gpurun_out/bench_r2_g2_$name.err; echo "rc=$? $name"; grep -v "^\s*$" gpurun_out/bench_r2_g2_$name.err | grep -v "OMP_NUM\|\*\*\*\*\|barrier\|return func" | tail -25; head -c 400 gpurun_out/bench_r2_g2_$name.json; echo; }
run replicated PM_RESIDENT_MIN_LOG=30
run resident PM_RESIDENT_MIN_LOG=20
