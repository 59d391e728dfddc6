"""The C++ restatement (oracle/cpu_ref.cpp, used as CPU baseline and mid-size checker) against the
Python-integer oracle.  CPU only."""
import random

import pytest

from oracle import cpp, curve, poly
from oracle.fields import R_MOD, Q_MOD
from polymath_b200 import codec


def test_cpp_field_mul():
    import ctypes as C
    lib = cpp.load()
    rnd = random.Random(1)
    a = [rnd.randrange(R_MOD) for _ in range(200)] + [0, 1, R_MOD - 1]
    b = [rnd.randrange(R_MOD) for _ in range(200)] + [R_MOD - 1, R_MOD - 1, R_MOD - 1]
    out = C.create_string_buffer(len(a) * 32)
    lib.orc_fr_mul(codec.frs_to_wire(a), codec.frs_to_wire(b), out, len(a))
    assert codec.frs_from_wire(out.raw) == [x * y % R_MOD for x, y in zip(a, b)]
    a = [rnd.randrange(Q_MOD) for _ in range(200)] + [Q_MOD - 1]
    b = [rnd.randrange(Q_MOD) for _ in range(200)] + [Q_MOD - 1]
    out = C.create_string_buffer(len(a) * 48)
    lib.orc_fq_mul(b"".join(codec.fq_to_wire(v) for v in a), b"".join(codec.fq_to_wire(v) for v in b), out, len(a))
    got = [codec.fq_from_wire(out.raw[i:i + 48]) for i in range(0, len(a) * 48, 48)]
    assert got == [x * y % Q_MOD for x, y in zip(a, b)]


@pytest.mark.parametrize("log_n", [1, 2, 5, 10, 12])
def test_cpp_ntt(log_n):
    rnd = random.Random(log_n)
    n = 1 << log_n
    v = [rnd.randrange(R_MOD) for _ in range(n)]
    d = poly.Domain(n)
    buf = bytearray(codec.frs_to_wire(v))
    cpp.ntt_wire(buf, log_n, False)
    assert codec.frs_from_wire(bytes(buf)) == d.fft(v)
    cpp.ntt_wire(buf, log_n, True)
    assert codec.frs_from_wire(bytes(buf)) == v


@pytest.mark.parametrize("n", [1, 5, 40, 700])
def test_cpp_msm(n):
    rnd = random.Random(n)
    tbl = curve.FixedBaseTable(curve.G1_GEN, window=8)
    bases = tbl.mul_many([rnd.randrange(1, R_MOD) for _ in range(n)])
    scalars = [rnd.randrange(R_MOD) for _ in range(n)]
    if n > 5:
        bases[2] = None
        scalars[3] = 0
        scalars[4] = R_MOD - 1
    got = codec.g1_from_wire(cpp.msm_wire(codec.g1s_to_wire(bases), codec.frs_to_wire(scalars), n))
    assert got == poly.msm_pippenger(scalars, bases)
    assert cpp.load().orc_msm_window_bits(n) == poly.ark_window_bits(n)
