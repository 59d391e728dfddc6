#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1 | cut -c1-140; env "$@" timeout 300 python tools/prove_once.py --log-n 20 --world 8 --iters 4 2>&1 | tail -1 | cut -c1-140; }
run PM_INV_FAN0=32
run PM_INV_FAN0=16
run PM_INV_FAN0=8
run PM_INV_FAN0=16 PM_INV_FAN1=8
