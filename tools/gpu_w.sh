#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python tools/prove_once.py --log-n 20 --world 8 --iters 4 2>&1 | tail -1 | cut -c1-140; }
run X=0
run PM_MSM_PRECOMP_D=18
run PM_MSM_PRECOMP_D=19
run PM_MSM_PRECOMP_C=16
run PM_MSM_PRECOMP_C=17
run PM_MSM_PRECOMP_C=19
run PM_MSM_PRECOMP_C=20
run PM_MSM_ROUNDS_BIAS=1
run PM_P1_ROUNDS_BIAS=1
run PM_MSM_HALVES=0
run PM_A_TABLES=0
