"""ctypes wrapper of oracle/libpm_oracle.so (C++/OpenMP restatement; test infrastructure / CPU baseline)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libpm_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError("oracle/libpm_oracle.so not built; run `make -C oracle` (or python build.py)")
        lib = C.CDLL(_PATH)
        lib.orc_ntt.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orc_msm.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.orc_msm_window_bits.argtypes = [C.c_size_t]
        lib.orc_fr_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orc_fq_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orc_make_bases.argtypes = [C.c_size_t, C.c_void_p]
        _lib = lib
    return _lib


def ntt_wire(buf: bytearray, log_n: int, inverse: bool):
    """In-place NTT on a buffer of Montgomery-form Fr (32 B each)."""
    arr = (C.c_char * len(buf)).from_buffer(buf)
    load().orc_ntt(C.addressof(arr), log_n, 1 if inverse else 0)


def msm_wire(bases: bytes, scalars: bytes, n: int) -> bytes:
    out = C.create_string_buffer(96)
    load().orc_msm(bases, scalars, n, out)
    return out.raw


def num_threads() -> int:
    return load().orc_num_threads()


def use_all_cores() -> int:
    """Run the OpenMP loops on every core this process may use (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    want = int(os.environ.get("PM_REF_THREADS", n))
    lib = load()
    lib.orc_set_num_threads.argtypes = [C.c_int]
    lib.orc_set_num_threads.restype = None
    lib.orc_set_num_threads(want)
    return lib.orc_num_threads()


def make_bases_wire(n: int) -> bytes:
    out = C.create_string_buffer(96 * n)
    load().orc_make_bases(n, out)
    return out.raw
