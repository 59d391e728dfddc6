#!/bin/bash
python - <<'PY'
import ctypes as C, sys
sys.path.insert(0,'.')
from polymath_b200.lib import require_device, check
lib = require_device()
d = C.c_double()
for f,name in ((0,'fr_mul'),(4,'fr_sqr'),(5,'fr_kara'),(1,'fq_mul'),(2,'fq_sqr'),(3,'fq_kara')):
    for rep in range(2):
        check(lib.pm_bench_field_mul(f, C.byref(d)))
    print(name, round(d.value/1e9,2), 'G/s')
PY
