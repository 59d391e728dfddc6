#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for lg in 10 11 12 14 16; do timeout 300 python tools/prove_once.py --log-n $lg --world 1 --iters 5 2>&1 | tail -1 | cut -c1-140; done
python __graft_entry__.py smoke 2>&1 | tail -2
