#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1 | cut -c1-140; env "$@" timeout 300 python tools/prove_once.py --log-n 20 --world 8 --iters 4 2>&1 | tail -1 | cut -c1-140; }
run PM_INV_FAN1=32
run PM_INV_FAN1=8
run PM_INV_FAN1=4
run PM_INV_FAN1=2
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_prover_gpu.py -m gpu -x -q 2>&1 | tail -3
