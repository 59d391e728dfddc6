#!/bin/bash
run() { echo "== $*"; env "$@" timeout 900 python tools/sweep.py --skip-basics --ntt 21,22,24 --msm "" --iters 5 2>&1 | grep '"inverse": 0' | cut -c28-130; }
run X=1
run PM_NTT_ROW_TILE=10
run PM_NTT_ROW_TILE=9
run PM_NTT_COL_TILE=10 PM_NTT_ROW_TILE=10
run PM_NTT_COL_TILE=9 PM_NTT_ROW_TILE=9
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "ntt" 2>&1 | tail -2
