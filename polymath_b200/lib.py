"""ctypes binding of libpolymath_b200.so (C ABI: include/polymath_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpolymath_b200.so")

FR_BYTES, FQ_BYTES, G1_BYTES = 32, 48, 96


class PolymathB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("polymath_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load the CUDA library.  Raises if it has not been built (python build.py): no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built; run `python build.py` (the product path has no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    u8p, sz = C.c_char_p, C.c_size_t
    lib.pm_last_error.restype = C.c_char_p
    lib.pm_abi_version.restype = C.c_int
    lib.pm_device_count.restype = C.c_int
    lib.pm_kernel_launches.restype = C.c_uint64
    for name in ("pm_fr_mul_batch", "pm_fr_add_batch", "pm_fr_sub_batch", "pm_fq_mul_batch"):
        f = getattr(lib, name)
        f.argtypes = [u8p, u8p, u8p, sz]
        f.restype = C.c_int
    for name in ("pm_fr_inv_batch", "pm_fq_inv_batch"):
        f = getattr(lib, name)
        f.argtypes = [u8p, u8p, sz]
        f.restype = C.c_int
    lib.pm_ntt_fr.argtypes = [u8p, C.c_uint, C.c_int, u8p]
    lib.pm_ntt_fr.restype = C.c_int
    lib.pm_msm_g1.argtypes = [u8p, sz, u8p, sz, u8p]
    lib.pm_msm_g1.restype = C.c_int
    lib.pm_msm_g1_window.argtypes = [u8p, sz, u8p, sz, C.c_int, C.c_int, u8p]
    lib.pm_msm_g1_window.restype = C.c_int
    lib.pm_msm_g1_levels.argtypes = [u8p, sz, u8p, sz, C.c_int, C.c_int, u8p]
    lib.pm_msm_g1_levels.restype = C.c_int
    lib.pm_msm_set_tuning.argtypes = [C.c_int]
    lib.pm_msm_set_tuning.restype = C.c_int
    lib.pm_g1_decompress_batch.argtypes = [u8p, sz, C.c_int, u8p]
    lib.pm_g1_decompress_batch.restype = C.c_int
    lib.pm_g1_compress_batch.argtypes = [u8p, sz, u8p]
    lib.pm_g1_compress_batch.restype = C.c_int
    lib.pm_fixed_base_mul.argtypes = [u8p, sz, u8p]
    lib.pm_fixed_base_mul.restype = C.c_int
    dp = C.POINTER(C.c_double)
    lib.pm_bench_imad_peak.argtypes = [dp]
    lib.pm_bench_field_mul.argtypes = [C.c_int, dp]
    lib.pm_bench_ntt.argtypes = [C.c_uint, C.c_int, C.c_int, dp]
    lib.pm_bench_g1_codec.argtypes = [sz, dp, dp]
    lib.pm_bench_g1_codec.restype = C.c_int
    lib.pm_bench_last_msm.argtypes = [dp]
    lib.pm_bench_last_msm.restype = C.c_int
    lib.pm_bench_msm.argtypes = [sz, C.c_int, C.c_int, dp, dp]
    lib.pm_bench_msm_levels.argtypes = [sz, C.c_int, C.c_int, C.c_int, dp, dp]
    lib.pm_bench_msm_levels.restype = C.c_int
    for name in ("pm_bench_imad_peak", "pm_bench_field_mul", "pm_bench_ntt", "pm_bench_msm"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise PolymathB200Error(code, (load().pm_last_error() or b"").decode())


def require_device():
    lib = load()
    if lib.pm_device_count() <= 0:
        raise PolymathB200Error(1, "no CUDA device visible; polymath_b200 has no CPU fallback")
    return lib
