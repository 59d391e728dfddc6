"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, and the
host-side (non-GPU) product code — RNG, transcript — matches the oracle.  No GPU needed."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "polymath_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from polymath_b200.lib import load
    lib = load()
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.pm_abi_version() == 1


def test_no_device_means_loud_failure():
    from polymath_b200.lib import load
    lib = load()
    if lib.pm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from polymath_b200 import kernels
    from polymath_b200.lib import PolymathB200Error
    with pytest.raises(PolymathB200Error):
        kernels.fr_mul_batch([1], [2])


def test_host_rng_matches_oracle():
    from polymath_b200.api import StdRng
    from oracle.rng import StdRng as ORng, fr_rand
    for seed in (0, 1, 12345, 2**64 - 1):
        a, b = StdRng.seed_from_u64(seed), ORng.seed_from_u64(seed)
        assert [a.next_u64() for _ in range(40)] == [b.next_u64() for _ in range(40)]
        assert [a.fr_rand() for _ in range(50)] == [fr_rand(b) for _ in range(50)]


def test_host_merlin_vector():
    from polymath_b200.api import _bind
    from polymath_b200.lib import load
    lib = load()
    _bind(lib)
    out = C.create_string_buffer(32)
    assert lib.pm_merlin_test_vector(out) == 0
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
