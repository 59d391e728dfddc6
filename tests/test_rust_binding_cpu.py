"""The Rust sys crate binds exactly the C ABI: ffi.rs is regenerated from include/polymath_b200.h and compared with the
committed copy, and every `pm_*` symbol the shared library exports is bound (and vice versa)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_rust_sys  # noqa: E402


def test_ffi_rs_is_generated_from_the_header():
    assert open(gen_rust_sys.OUT).read() == gen_rust_sys.generate(), "run python tools/gen_rust_sys.py"


def test_every_exported_symbol_is_bound():
    so = os.path.join(ROOT, "polymath_b200", "libpolymath_b200.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pm_\w+)", out))
    bound = set(gen_rust_sys.function_names())
    assert exported == bound, (sorted(exported - bound), sorted(bound - exported))
    ffi = open(gen_rust_sys.OUT).read()
    for name in bound:
        assert "pub fn %s(" % name in ffi
    # the safe layer only calls functions that exist
    lib = open(os.path.join(ROOT, "rust", "polymath-b200-sys", "src", "lib.rs")).read()
    for name in set(re.findall(r"ffi::(pm_\w+)\(", lib)):
        assert name in bound, name
    assert "assert_eq!(ffi::BOUND_SYMBOLS.len(), %d);" % len(bound) in lib
