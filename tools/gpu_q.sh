#!/bin/bash
set -x
for g in 3 4 5 6; do
echo "== PM_FWD_CTAS=$g"
PM_FWD_CTAS=$g timeout 600 python tools/prove_once.py --log-n 20 --world 1 --iters 4 2>&1 | tail -1
PM_FWD_CTAS=$g timeout 600 python tools/sweep.py --skip-basics --ntt "" --msm 22 --iters 3 2>&1 | cut -c60-200
done
