#!/bin/bash
for b in 11 10; do
echo "== PM_NTT_TILE_BITS=$b"
PM_NTT_TILE_BITS=$b timeout 900 python tools/sweep.py --skip-basics --ntt 18,19,20 --msm "" --iters 5 2>&1 | cut -c1-150
done
