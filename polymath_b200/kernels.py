"""Python-level wrappers of the standalone kernel entry points (parity tests and the kernel sweep)."""
import ctypes as C

from . import codec
from .lib import check, require_device


def fr_mul_batch(a, b):
    lib = require_device()
    out = C.create_string_buffer(len(a) * 32)
    check(lib.pm_fr_mul_batch(codec.frs_to_wire(a), codec.frs_to_wire(b), out, len(a)))
    return codec.frs_from_wire(out.raw)


def fr_add_batch(a, b):
    lib = require_device()
    out = C.create_string_buffer(len(a) * 32)
    check(lib.pm_fr_add_batch(codec.frs_to_wire(a), codec.frs_to_wire(b), out, len(a)))
    return codec.frs_from_wire(out.raw)


def fr_sub_batch(a, b):
    lib = require_device()
    out = C.create_string_buffer(len(a) * 32)
    check(lib.pm_fr_sub_batch(codec.frs_to_wire(a), codec.frs_to_wire(b), out, len(a)))
    return codec.frs_from_wire(out.raw)


def fq_mul_batch(a, b):
    lib = require_device()
    out = C.create_string_buffer(len(a) * 48)
    wa = b"".join(codec.fq_to_wire(v) for v in a)
    wb = b"".join(codec.fq_to_wire(v) for v in b)
    check(lib.pm_fq_mul_batch(wa, wb, out, len(a)))
    return [codec.fq_from_wire(out.raw[i:i + 48]) for i in range(0, len(a) * 48, 48)]


def fr_inv_batch(a):
    lib = require_device()
    out = C.create_string_buffer(len(a) * 32)
    check(lib.pm_fr_inv_batch(codec.frs_to_wire(a), out, len(a)))
    return codec.frs_from_wire(out.raw)


def fq_inv_batch(a):
    lib = require_device()
    out = C.create_string_buffer(len(a) * 48)
    check(lib.pm_fq_inv_batch(b"".join(codec.fq_to_wire(v) for v in a), out, len(a)))
    return [codec.fq_from_wire(out.raw[i:i + 48]) for i in range(0, len(a) * 48, 48)]


def ntt_fr(values, inverse=False, coset_gen=None):
    lib = require_device()
    n = len(values)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    buf = C.create_string_buffer(codec.frs_to_wire(values), n * 32)
    g = codec.fr_to_wire(coset_gen) if coset_gen is not None else None
    check(lib.pm_ntt_fr(buf, log_n, 1 if inverse else 0, g))
    return codec.frs_from_wire(buf.raw)


def msm_g1(bases, scalars, window_bits=0, heavy_threshold=0, stride=96, levels=1):
    lib = require_device()
    n = min(len(bases), len(scalars))
    if levels > 1:
        out = C.create_string_buffer(96)
        check(lib.pm_msm_g1_levels(codec.g1s_to_wire(bases[:n]), 96, codec.frs_to_wire(scalars[:n]), n, window_bits, levels, out))
        return codec.g1_from_wire(out.raw)
    if stride == 96:
        wb = codec.g1s_to_wire(bases[:n])
    else:
        wb = b"".join((codec.g1_to_wire(p) + (b"\x01" if p is None else b"\x00") + bytes(stride - 97)) for p in bases[:n])
    out = C.create_string_buffer(96)
    check(lib.pm_msm_g1_window(wb, stride, codec.frs_to_wire(scalars[:n]), n, window_bits, heavy_threshold, out))
    return codec.g1_from_wire(out.raw)


def msm_set_tuning(rounds=-1):
    """Pair rounds of every later MSM (-1 = automatic, 0 = XYZZ walk only); results never depend on it."""
    lib = require_device()
    check(lib.pm_msm_set_tuning(rounds))


def g1_decompress_batch(encodings: bytes, validate=False):
    """48-byte compressed encodings -> list of affine points / None (`deserialize_compressed[_unchecked]`)."""
    lib = require_device()
    n = len(encodings) // 48
    out = C.create_string_buffer(max(n * 96, 1))
    check(lib.pm_g1_decompress_batch(bytes(encodings), n, 1 if validate else 0, out))
    return codec.g1s_from_wire(out.raw[:n * 96])


def g1_compress_batch(points) -> bytes:
    """Affine points / None -> concatenated 48-byte compressed encodings (`serialize_compressed`)."""
    lib = require_device()
    n = len(points)
    out = C.create_string_buffer(max(n * 48, 1))
    check(lib.pm_g1_compress_batch(codec.g1s_to_wire(points), n, out))
    return out.raw[:n * 48]


def fixed_base_mul(scalars):
    lib = require_device()
    out = C.create_string_buffer(len(scalars) * 96)
    check(lib.pm_fixed_base_mul(codec.frs_to_wire(scalars), len(scalars), out))
    return codec.g1s_from_wire(out.raw)
