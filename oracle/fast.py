"""A complete CPU prove at the benchmark sizes (oracle; TEST INFRASTRUCTURE / CPU BASELINE ONLY).

`create_proof_with_assignment` of /root/reference/src/prover.rs:66-237 with the reference's own data flow —
the same five large MSMs (prover.rs:114-121,229), the four iFFTs (:94,:96,:161,:165), fft(2n)/square/ifft(2n)
(:315-328), the Horner evaluation (:132), the sparse assembly and the division by (X - x1) (:142-225) — where every
O(n) step runs in oracle/cpu_ref.cpp (C++/OpenMP on all host cores) and the Fiat-Shamir transcript, the O(m0)
scalar algebra and the serialisation are oracle/polymath.py's.  The dense O(n*m) SAP evaluation of the reference
(prover.rs:87-96) cannot run beyond n = 2^11; the sparse closed form replaces it, checked against
the literal one in tests/test_oracle.py.

It is what `bench.py --impl reference` and the `cpu_baseline` leg time, and — given the same key, witness and r_a —
it must return the very bytes the device path returns (tests/test_fast_oracle_cpu.py pins it to
oracle/polymath.py; bench.py compares it with the GPU proof at n = 2^20).

Buffers are numpy uint64 arrays in the device ABI's wire form (Montgomery little-endian limbs: Fr = 4 words,
G1 affine = 12 words, (0,0) = infinity), so exported device keys feed it without conversion.
PARITY UNPINNED against real arkworks (oracle/__init__.py).
"""
import ctypes as C
import time

import numpy as np

from . import cpp
from . import polymath as opm
from .curve import g1_add
from .fields import R_MOD, Q_MOD, FR_MONT_R, FR_MONT_RINV, FQ_MONT_RINV
from .merlin import MerlinFieldTranscript

P = R_MOD
KEY_NAMES = ["x_powers_g1", "x_powers_y_alpha_g1", "x_powers_zh_by_y_alpha_g1", "x_powers_y_gamma_g1",
             "x_powers_y_gamma_z_g1", "uj_wj_lcs_by_y_alpha_g1"]      # data_structures.rs:56-73 field order


def _lib():
    lib = cpp.load()
    if not getattr(lib, "_fast_bound", False):
        vp, sz = C.c_void_p, C.c_size_t
        lib.orc_spmv.argtypes = [vp, vp, vp, sz, vp, vp]
        lib.orc_sap_evals.argtypes = [sz, sz, sz, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.orc_fr_square.argtypes = [vp, sz]
        lib.orc_quotient.argtypes = [vp, vp, sz, vp]
        lib.orc_two_ra_u.argtypes = [vp, sz, vp, vp]
        lib.orc_horner.argtypes = [vp, sz, vp, vp]
        lib.orc_d_numerator.argtypes = [sz, sz, vp, vp, vp, vp, vp, vp]
        lib.orc_divide_linear.argtypes = [vp, sz, vp, vp, vp]
        lib._fast_bound = True
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fr_wire(vals):
    """ints -> (k, 4) uint64 Montgomery limbs"""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % P) * FR_MONT_R % P
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def fr_int(row):
    m = 0
    for k in range(4):
        m |= int(row[k]) << (64 * k)
    return m * FR_MONT_RINV % P


def g1_int(row):
    """12 uint64 words (x | y, Montgomery) -> affine ints or None"""
    x = y = 0
    for k in range(6):
        x |= int(row[k]) << (64 * k)
        y |= int(row[6 + k]) << (64 * k)
    if x == 0 and y == 0:
        return None
    return x * FQ_MONT_RINV % Q_MOD, y * FQ_MONT_RINV % Q_MOD


def msm(bases, scalars, count=None):
    """`msm()` of prover.rs:380-384 over wire buffers -> affine ints / None"""
    n = min(len(bases), len(scalars)) if count is None else count
    out = np.zeros(12, dtype=np.uint64)
    if n:
        b = np.ascontiguousarray(bases[:n])
        s = np.ascontiguousarray(scalars[:n])
        cpp.load().orc_msm(_p(b), _p(s), n, _p(out))
    return g1_int(out)


def _ntt(a, log_n, inverse):
    cpp.load().orc_ntt(_p(a), log_n, 1 if inverse else 0)


class Csr:
    """One R1CS matrix as `to_matrices()` rows flattened: row_ptr uint64[nr+1], col uint32[nnz], val uint64[nnz,4]."""

    def __init__(self, row_ptr, col, val):
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.val = np.ascontiguousarray(val, dtype=np.uint64).reshape(-1, 4)

    @classmethod
    def from_rows(cls, rows):
        rp, col, vals = [0], [], []
        for row in rows:
            seen = set()
            for cf, j in row:
                if j in seen:          # m_at keeps the first entry of a column (common.rs:100-105)
                    continue
                seen.add(j)
                col.append(j)
                vals.append(cf)
            rp.append(len(col))
        return cls(np.array(rp, dtype=np.uint64), np.array(col, dtype=np.uint32), fr_wire(vals) if vals else np.zeros((0, 4), np.uint64))


def prove(key, a: Csr, b: Csr, c: Csr, m0, mw, nr, n, sigma, omega, x_wire, w_wire, ra, timings=None,
          transcript_cls=MerlinFieldTranscript):
    """key: dict KEY_NAMES -> (len, 12) uint64 arrays; x_wire (m0, 4), w_wire (mw, 4) uint64; ra = [r0, r1] ints.
    Returns oracle.polymath.Proof.  Raises AssertionError where the reference panics."""
    lib = _lib()
    t0 = time.perf_counter()
    marks = {}

    def mark(name):
        nonlocal t0
        t1 = time.perf_counter()
        marks[name] = marks.get(name, 0.0) + (t1 - t0)
        t0 = t1

    log_n = n.bit_length() - 1
    assert 1 << log_n == n and sigma == n + 3
    zp = np.ascontiguousarray(np.concatenate([x_wire, w_wire]))
    az, bz, cz = (np.zeros((nr, 4), np.uint64) for _ in range(3))
    for mat, out in ((a, az), (b, bz), (c, cz)):
        lib.orc_spmv(_p(mat.row_ptr), _p(mat.col), _p(mat.val), nr, _p(zp), _p(out))
    y = np.zeros((m0 + nr, 4), np.uint64)
    u, w, wu = (np.zeros((n, 4), np.uint64) for _ in range(3))
    assert lib.orc_sap_evals(m0, nr, n, _p(x_wire), _p(az), _p(bz), _p(cz), _p(y), _p(u), _p(w), _p(wu)) == 0
    ww = w.copy()                                    # witness-column part of W.z == W.z (prover.rs:165)
    mark("sap")
    _ntt(u, log_n, True)                             # poly_coeffs x4: prover.rs:94,96,161,165
    _ntt(w, log_n, True)
    _ntt(wu, log_n, True)
    _ntt(ww, log_n, True)
    u2 = np.zeros((2 * n, 4), np.uint64)             # square_polynomial, prover.rs:315-328
    u2[:n] = u
    _ntt(u2, log_n + 1, False)
    lib.orc_fr_square(_p(u2), 2 * n)
    _ntt(u2, log_n + 1, True)
    hnum = np.zeros((2 * n, 4), np.uint64)
    st = lib.orc_quotient(_p(u2), _p(w), n, _p(hnum))
    assert st != 4, "h is zero or too large (prover.rs:107)"
    assert st == 0, "witness does not satisfy the SAP (prover.rs:108)"
    mark("ntt")
    ra_w = fr_wire(ra)
    # compute_a_g1, prover.rs:330-338
    a_g1 = g1_add(msm(key["x_powers_g1"], u), msm(key["x_powers_y_alpha_g1"], ra_w))
    # compute_r_g1, prover.rs:340-357
    two_ra_u = np.zeros((n + 1, 4), np.uint64)
    lib.orc_two_ra_u(_p(u), n, _p(ra_w), _p(two_ra_u))
    ra_sq = fr_wire([ra[0] * ra[0], 2 * ra[0] * ra[1], ra[1] * ra[1]])
    r_g1 = g1_add(g1_add(msm(key["x_powers_g1"], two_ra_u), msm(key["x_powers_y_alpha_g1"], ra_sq)),
                  msm(key["x_powers_y_gamma_g1"], ra_w))
    h_g1 = msm(key["x_powers_zh_by_y_alpha_g1"], hnum[n:2 * n - 1])                 # prover.rs:118
    z_tail = np.ascontiguousarray(np.concatenate([x_wire, w_wire, y]))             # z[1..].concat(), prover.rs:120-121
    lcs_g1 = msm(key["uj_wj_lcs_by_y_alpha_g1"], z_tail)
    c_g1 = g1_add(g1_add(lcs_g1, h_g1), r_g1)                                       # prover.rs:123
    mark("msm_phase1")

    x_ints = [fr_int(r) for r in x_wire]
    t = transcript_cls(opm.B_POLYMATH)
    x1 = opm.compute_x1(t, x_ints, [a_g1, c_g1])                                    # prover.rs:125-126
    y1 = opm.compute_y1(x1, sigma)
    y1_alpha = opm.neg_power(y1, opm.MINUS_ALPHA)
    ev = np.zeros(4, np.uint64)
    x1_w = fr_wire([x1])
    lib.orc_horner(_p(u), n, _p(x1_w), _p(ev))
    a_at_x1 = (fr_int(ev) + (ra[0] + ra[1] * x1) * y1_alpha) % P                   # prover.rs:132
    y1_gamma = opm.neg_power(y1, opm.MINUS_GAMMA)

    class _Vk:
        pass
    vk = _Vk()
    vk.n, vk.omega = n, omega
    pi_at_x1 = opm.compute_pi_at_x1(vk, x_ints, x1, y1_gamma)
    c_at_x1 = opm.compute_c_at_x1(y1_gamma, y1_alpha, a_at_x1, pi_at_x1)
    x2 = opm.compute_x2(t, x1, [a_at_x1, c_at_x1])                                  # prover.rs:189
    size = 2 * (n - 1) + 8 * sigma + 1
    num = np.zeros((size, 4), np.uint64)
    consts = fr_wire([ra[0], ra[1], x2, a_at_x1, c_at_x1])
    lib.orc_d_numerator(n, sigma, _p(u), _p(wu), _p(ww), _p(hnum), _p(consts), _p(num))
    q = np.zeros((size - 1, 4), np.uint64)
    rem = np.zeros(4, np.uint64)
    lib.orc_divide_linear(_p(num), size, _p(x1_w), _p(q), _p(rem))
    assert not rem.any(), "opening remainder non-zero (prover.rs:221)"
    mark("opening")
    d_g1 = msm(key["x_powers_y_gamma_z_g1"], q)                                     # prover.rs:229
    mark("msm_d")
    if timings is not None:
        timings.update(marks)
    return opm.Proof(a_g1=a_g1, c_g1=c_g1, a_at_x1=a_at_x1, d_g1=d_g1)


def key_from_oracle(pk):
    """oracle.polymath.ProvingKey (affine ints) -> wire arrays (tests; small keys only)."""
    from .fields import FQ_MONT_R
    out = {}
    for name in KEY_NAMES:
        pts = getattr(pk, name)
        arr = np.zeros((len(pts), 12), np.uint64)
        for i, pt in enumerate(pts):
            if pt is None:
                continue
            for half, v in enumerate(pt):
                m = v * FQ_MONT_R % Q_MOD
                for k in range(6):
                    arr[i, 6 * half + k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
        out[name] = arr
    return out


def synthetic_key(n, m0, cols):
    """Arbitrary valid curve points in the shape of a ProvingKey (sequential multiples of G, orc_make_bases): the MSM
    work of a prove does not depend on the base values, so the CPU arm can be TIMED without a setup; the resulting
    proof is of course not a valid one.  Used only when no exported key is available."""
    lens = {"x_powers_g1": n + 1, "x_powers_y_alpha_g1": 3, "x_powers_zh_by_y_alpha_g1": n - 1, "x_powers_y_gamma_g1": 2,
            "x_powers_y_gamma_z_g1": 2 * (n - 1) + 8 * (n + 3) + 1, "uj_wj_lcs_by_y_alpha_g1": cols - m0}
    biggest = max(lens.values())
    raw = np.frombuffer(cpp.make_bases_wire(biggest), dtype=np.uint64).reshape(-1, 12)
    return {name: raw[:ln] for name, ln in lens.items()}
