"""polymath_b200 — B200 (sm_100a) prover backend for the Polymath zk-SNARK (sigma0-dev/polymath).

The product is the CUDA library `libpolymath_b200.so` (C ABI in `include/polymath_b200.h`);
this package is the thin Python binding used by the tests and `bench.py`.  There is no CPU
fallback: importing `polymath_b200.lib` raises if the library has not been built, and every
compute call fails loudly when no CUDA device is present.
"""
from .lib import load, PolymathB200Error  # noqa: F401
