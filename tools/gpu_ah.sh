#!/bin/bash
run() { echo "== $*"; for lg in 11 13; do env "$@" timeout 300 python tools/prove_once.py --log-n $lg --world 1 --iters 5 2>&1 | tail -1 | cut -c1-140; done; }
run X=1
run PM_MSM_PRECOMP_MIN=256
run PM_MSM_PRECOMP_MIN=256 PM_MSM_PRECOMP_C=12
run PM_MSM_PRECOMP_MIN=256 PM_MSM_PRECOMP_C=13
run PM_MSM_PRECOMP_MIN=256 PM_MSM_PRECOMP_C=14
run PM_MSM_PRECOMP_MIN=256 PM_MSM_PRECOMP_C=12 PM_MSM_PRECOMP_D=14
run PM_MSM_PRECOMP_MIN=256 PM_MSM_PRECOMP_C=13 PM_MSM_PRECOMP_D=14
