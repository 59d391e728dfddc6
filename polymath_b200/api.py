"""Python mirror of the reference's public surface for the proving hot path.

`Polymath.setup` / `Polymath.prove` follow `impl SNARK for Polymath<Bls12_381, MerlinFieldTranscript<Fr>>`
(/root/reference/src/lib.rs:52-91) at the level below circuit synthesis: they take the R1CS
matrices (`cs.to_matrices()`, generator.rs:46) and the instance / witness assignments
(prover.rs:53-58) that the caller's circuit produces.  All computation happens in
libpolymath_b200.so (CUDA); this file only marshals bytes.  Verification (pairing check) stays
on the host (BASELINE.json north_star): `Polymath.verify` / `verify_batch` call the library's host-only
verifier (`pm_polymath_verify`, src/verifier.rs:19-62) and need no GPU.
"""
import ctypes as C

from . import codec
from .lib import check, require_device, load


class R1CSView(C.Structure):
    _fields_ = [
        ("num_instance_variables", C.c_uint64),
        ("num_r1cs_witness_variables", C.c_uint64),
        ("num_r1cs_constraints", C.c_uint64),
        ("a_row_ptr", C.c_void_p), ("a_col", C.c_void_p), ("a_val", C.c_void_p),
        ("b_row_ptr", C.c_void_p), ("b_col", C.c_void_p), ("b_val", C.c_void_p),
        ("c_row_ptr", C.c_void_p), ("c_col", C.c_void_p), ("c_val", C.c_void_p),
    ]


class PKView(C.Structure):
    _fields_ = [
        ("r1cs", R1CSView),
        ("n", C.c_uint64), ("sigma", C.c_uint64), ("point_stride", C.c_size_t),
        ("x_powers_g1", C.c_void_p), ("x_powers_g1_len", C.c_uint64),
        ("x_powers_y_alpha_g1", C.c_void_p), ("x_powers_y_alpha_g1_len", C.c_uint64),
        ("x_powers_zh_by_y_alpha_g1", C.c_void_p), ("x_powers_zh_by_y_alpha_g1_len", C.c_uint64),
        ("x_powers_y_gamma_g1", C.c_void_p), ("x_powers_y_gamma_g1_len", C.c_uint64),
        ("x_powers_y_gamma_z_g1", C.c_void_p), ("x_powers_y_gamma_z_g1_len", C.c_uint64),
        ("uj_wj_lcs_by_y_alpha_g1", C.c_void_p), ("uj_wj_lcs_by_y_alpha_g1_len", C.c_uint64),
    ]


TRANSCRIPTS = {"merlin": 0, "keccak256": 1, "blake3": 2}     # PM_TRANSCRIPT_* (src/transcript/{merlin,keccak256,blake3}.rs)

KEY_NAMES = ("x_powers_g1", "x_powers_y_alpha_g1", "x_powers_zh_by_y_alpha_g1", "x_powers_y_gamma_g1",
             "x_powers_y_gamma_z_g1", "uj_wj_lcs_by_y_alpha_g1")


def _bind(lib):
    if getattr(lib, "_api_bound", False):
        return
    vp, u8p = C.c_void_p, C.c_char_p
    lib.pm_rng_seed_from_u64.argtypes = [C.c_uint64]
    lib.pm_rng_seed_from_u64.restype = vp
    lib.pm_rng_from_seed.argtypes = [u8p]
    lib.pm_rng_from_seed.restype = vp
    lib.pm_rng_free.argtypes = [vp]
    lib.pm_rng_free.restype = None
    lib.pm_rng_next_u64.argtypes = [vp]
    lib.pm_rng_next_u64.restype = C.c_uint64
    lib.pm_rng_fr_rand.argtypes = [vp, u8p]
    lib.pm_rng_fr_rand.restype = None
    lib.pm_merlin_test_vector.argtypes = [u8p]
    lib.pm_ctx_create.argtypes = [C.POINTER(PKView), C.POINTER(vp)]
    lib.pm_ctx_destroy.argtypes = [vp]
    lib.pm_ctx_destroy.restype = None
    lib.pm_setup.argtypes = [C.POINTER(R1CSView), u8p, u8p, C.POINTER(vp), u8p, u8p]
    lib.pm_ctx_dims.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.pm_ctx_key_len.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
    lib.pm_ctx_export_key.argtypes = [vp, C.c_int, u8p, C.c_size_t]
    lib.pm_prove_phase1.argtypes = [vp, u8p, u8p, u8p, u8p, u8p]
    lib.pm_ctx_set_assignment.argtypes = [vp, u8p, u8p]
    lib.pm_prove_phase1_resident.argtypes = [vp, u8p, u8p, u8p]
    lib.pm_prove_phase2.argtypes = [vp, u8p, u8p, u8p]
    lib.pm_prove_phase3.argtypes = [vp, u8p, u8p, u8p]
    lib.pm_ctx_debug_read.argtypes = [vp, C.c_int, u8p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.pm_ctx_phase_ms.argtypes = [vp, C.POINTER(C.c_double)]
    lib.pm_polymath_setup.argtypes = [C.POINTER(R1CSView), vp, C.POINTER(vp), u8p]
    lib.pm_polymath_prove.argtypes = [vp, u8p, u8p, vp, u8p]
    lib.pm_polymath_prove_resident.argtypes = [vp, u8p, vp, u8p]
    lib.pm_polymath_verify.argtypes = [u8p, u8p, C.c_size_t, u8p, C.POINTER(C.c_int)]
    lib.pm_polymath_verify_transcript.argtypes = [u8p, u8p, C.c_size_t, u8p, C.c_int, C.POINTER(C.c_int)]
    lib.pm_polymath_prove_transcript.argtypes = [vp, u8p, u8p, vp, C.c_int, u8p]
    lib.pm_host_hash.argtypes = [C.c_int, u8p, C.c_size_t, u8p]
    lib.pm_polymath_verify_batch.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, u8p, u8p, C.POINTER(C.c_int)]
    lib.pm_host_pairing_product_is_one.argtypes = [u8p, u8p, C.c_int, C.POINTER(C.c_int)]
    lib.pm_timer_start.argtypes = []
    lib.pm_timer_stop.argtypes = [C.POINTER(C.c_double)]
    lib.pm_bench_set_kernel_timing.argtypes = [C.c_int]
    lib.pm_bench_last_kernel_ms.argtypes = [C.POINTER(C.c_double)]
    lib._api_bound = True


def _lib():
    lib = require_device()
    _bind(lib)
    return lib


class StdRng:
    """rand 0.8 `StdRng::seed_from_u64` (the reference's RNG: benches/bench.rs:65, tests/mimc.rs:153)."""

    def __init__(self, seed_u64=None, seed_bytes=None):
        lib = load()
        _bind(lib)
        self._lib = lib
        if seed_bytes is not None:
            self._h = lib.pm_rng_from_seed(bytes(seed_bytes))
        else:
            self._h = lib.pm_rng_seed_from_u64(int(seed_u64))

    @classmethod
    def seed_from_u64(cls, s):
        return cls(seed_u64=s)

    def next_u64(self):
        return self._lib.pm_rng_next_u64(self._h)

    def fr_rand(self):
        out = C.create_string_buffer(32)
        self._lib.pm_rng_fr_rand(self._h, out)
        return codec.fr_from_wire(out.raw)

    def __del__(self):
        try:
            if self._h:
                self._lib.pm_rng_free(self._h)
                self._h = None
        except Exception:
            pass


class R1CS:
    """R1CS matrices as `cs.to_matrices()` yields them: rows of (coeff, column) with canonical-int coefficients."""

    def __init__(self, num_instance_variables, num_witness_variables, a, b, c):
        assert len(a) == len(b) == len(c)
        self.m0, self.mw, self.nr = num_instance_variables, num_witness_variables, len(a)
        self.a, self.b, self.c = a, b, c
        self._keep = []
        self.view = self._build_view()

    def _flatten(self, mat):
        row_ptr = (C.c_uint64 * (self.nr + 1))()
        nnz = sum(len(r) for r in mat)
        col = (C.c_uint32 * max(nnz, 1))()
        val = bytearray()
        k = 0
        cache = {}
        for i, row in enumerate(mat):
            row_ptr[i] = k
            for coeff, j in row:
                col[k] = j
                w = cache.get(coeff)
                if w is None:
                    w = codec.fr_to_wire(coeff)
                    if len(cache) < 4096:
                        cache[coeff] = w
                val += w
                k += 1
        row_ptr[self.nr] = k
        vbuf = C.create_string_buffer(bytes(val), max(len(val), 1))
        self._keep += [row_ptr, col, vbuf]
        return C.addressof(row_ptr), C.addressof(col), C.addressof(vbuf)

    def _build_view(self):
        v = R1CSView()
        v.num_instance_variables, v.num_r1cs_witness_variables, v.num_r1cs_constraints = self.m0, self.mw, self.nr
        v.a_row_ptr, v.a_col, v.a_val = self._flatten(self.a)
        v.b_row_ptr, v.b_col, v.b_val = self._flatten(self.b)
        v.c_row_ptr, v.c_col, v.c_val = self._flatten(self.c)
        return v


class ProvingKey:
    """Device-resident proving key (opaque `pm_ctx`) + the compressed verifying key bytes."""

    def __init__(self, handle, vk_bytes=None):
        self._lib = _lib()
        self._h = handle
        self.vk_bytes = vk_bytes
        n, s, cols = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self._lib.pm_ctx_dims(self._h, C.byref(n), C.byref(s), C.byref(cols)))
        self.n, self.sigma, self.num_columns = n.value, s.value, cols.value

    def key_len(self, which):
        ln = C.c_uint64()
        check(self._lib.pm_ctx_key_len(self._h, which, C.byref(ln)))
        return ln.value

    def export_key(self, which, stride=96):
        """One of the six G1 vectors (index into KEY_NAMES) as a list of affine points / None."""
        ln = self.key_len(which)
        buf = C.create_string_buffer(max(ln * stride, 1))
        check(self._lib.pm_ctx_export_key(self._h, which, buf, stride))
        raw = buf.raw[:ln * stride]
        if stride == 48:
            return raw                       # `serialize_compressed` bytes of the vector, without the length prefix
        if stride == 96:
            return codec.g1s_from_wire(raw)
        out = []
        for i in range(ln):
            rec = raw[i * stride:(i + 1) * stride]
            out.append(None if rec[96] else codec.g1_from_wire(rec[:96]))
        return out

    def debug_read(self, which):
        ln = C.c_uint64()
        check(self._lib.pm_ctx_debug_read(self._h, which, None, 0, C.byref(ln)))
        buf = C.create_string_buffer(max(ln.value * 32, 1))
        check(self._lib.pm_ctx_debug_read(self._h, which, buf, ln.value, C.byref(ln)))
        return codec.frs_from_wire(buf.raw[:ln.value * 32])

    def phase_ms(self):
        ms = (C.c_double * 3)()
        check(self._lib.pm_ctx_phase_ms(self._h, ms))
        return list(ms)

    def close(self):
        if self._h:
            self._lib.pm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _checked_frs(vals, what):
    """Field elements as Montgomery wire bytes; values outside [0, r) are refused, never reduced (the reference
    takes typed `F` values, lib.rs:72-86, and so cannot alias x with x + r)."""
    vals = list(vals)
    for v in vals:
        if not 0 <= v < codec.R_MOD:
            raise ValueError("%s: %d is not a reduced BLS12-381 scalar (0 <= v < r)" % (what, v))
    return codec.frs_to_wire(vals)


class Polymath:
    """`Polymath::<Bls12_381, MerlinFieldTranscript<Fr>>` (src/lib.rs:44-98), proving side."""

    @staticmethod
    def setup(r1cs: R1CS, rng: StdRng):
        """`circuit_specific_setup` (lib.rs:63-70): returns (pk, vk_bytes)."""
        lib = _lib()
        h = C.c_void_p()
        vk = C.create_string_buffer(392)
        check(lib.pm_polymath_setup(C.byref(r1cs.view), rng._h, C.byref(h), vk))
        pk = ProvingKey(h, vk.raw)
        pk._r1cs = r1cs
        return pk, vk.raw

    @staticmethod
    def setup_with_trapdoors(r1cs: R1CS, x: int, z: int):
        """`pm_setup` with explicit trapdoors (tests): returns (pk, x_g2, z_g2) with G2 points as Fq2 pairs."""
        lib = _lib()
        h = C.c_void_p()
        xg2, zg2 = C.create_string_buffer(192), C.create_string_buffer(192)
        check(lib.pm_setup(C.byref(r1cs.view), codec.fr_to_wire(x), codec.fr_to_wire(z), C.byref(h), xg2, zg2))

        def g2(raw):
            v = [codec.fq_from_wire(raw[i:i + 48]) for i in range(0, 192, 48)]
            return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))
        pk = ProvingKey(h)
        pk._r1cs = r1cs
        return pk, g2(xg2.raw), g2(zg2.raw)

    @staticmethod
    def load_key(r1cs: R1CS, n, sigma, vectors, stride=96):
        """`pm_ctx_create`: upload a host ProvingKey (dict KEY_NAMES -> list of affine points)."""
        lib = _lib()
        v = PKView()
        v.r1cs = r1cs.view
        v.n, v.sigma, v.point_stride = n, sigma, stride
        keep = []
        for name in KEY_NAMES:
            pts = vectors[name]
            if stride == 48:
                raw = pts if isinstance(pts, (bytes, bytearray)) else b"".join(pts)   # compressed encodings
                pts = range(len(raw) // 48)
            elif stride == 96:
                raw = codec.g1s_to_wire(pts)
            else:
                raw = b"".join(codec.g1_to_wire(p) + (b"\x01" if p is None else b"\x00") + bytes(stride - 97) for p in pts)
            buf = C.create_string_buffer(raw, max(len(raw), 1))
            keep.append(buf)
            setattr(v, name, C.addressof(buf))
            setattr(v, name + "_len", len(pts))
        h = C.c_void_p()
        check(lib.pm_ctx_create(C.byref(v), C.byref(h)))
        pk = ProvingKey(h)
        pk._r1cs = r1cs
        return pk

    @staticmethod
    def prove(pk: ProvingKey, instance, witness, rng: StdRng, transcript="merlin") -> bytes:
        """`prove` (lib.rs:72-78) below synthesis: instance = [1, public...], witness values; returns the
        176-byte compressed Proof.  `transcript`: "merlin" (the north-star type), "keccak256" or "blake3"."""
        lib = _lib()
        out = C.create_string_buffer(176)
        check(lib.pm_polymath_prove_transcript(pk._h, _checked_frs(instance, "instance"), _checked_frs(witness, "witness"),
                                               rng._h, TRANSCRIPTS[transcript], out))
        return out.raw

    @staticmethod
    def prove_phases(pk: ProvingKey, instance, witness, r_a, challenge_fn):
        """Drive the three phase calls directly (what the Rust integration does); `challenge_fn` supplies
        (x1, y1_alpha) from (a, c) and (x2, c_at_x1) from a_at_x1.  Returns (a, c, a_at_x1, d)."""
        lib = _lib()
        a, c = C.create_string_buffer(96), C.create_string_buffer(96)
        check(lib.pm_prove_phase1(pk._h, codec.frs_to_wire(instance), codec.frs_to_wire(witness),
                                  codec.frs_to_wire(r_a), a, c))
        a_pt, c_pt = codec.g1_from_wire(a.raw), codec.g1_from_wire(c.raw)
        x1, y1_alpha = challenge_fn.first(a_pt, c_pt)
        ev = C.create_string_buffer(32)
        check(lib.pm_prove_phase2(pk._h, codec.fr_to_wire(x1), codec.fr_to_wire(y1_alpha), ev))
        a_at_x1 = codec.fr_from_wire(ev.raw)
        x2, c_at_x1 = challenge_fn.second(a_at_x1)
        d = C.create_string_buffer(96)
        check(lib.pm_prove_phase3(pk._h, codec.fr_to_wire(x2), codec.fr_to_wire(c_at_x1), d))
        return a_pt, c_pt, a_at_x1, codec.g1_from_wire(d.raw)

    @staticmethod
    def verify(vk_bytes: bytes, public_inputs, proof_bytes: bytes, transcript="merlin") -> bool:
        """`verify` (lib.rs:80-91 -> verifier.rs:19-62) on the host: compressed VerifyingKey (392 B), the public
        inputs WITHOUT the leading one, compressed Proof (176 B).  No GPU involved."""
        lib = load()
        _bind(lib)
        if len(vk_bytes) != 392 or len(proof_bytes) != 176:
            raise ValueError("vk must be 392 bytes and the proof 176 bytes")
        ok = C.c_int(0)
        pub = _checked_frs(public_inputs, "public input")
        check(lib.pm_polymath_verify_transcript(vk_bytes, pub, len(public_inputs), proof_bytes, TRANSCRIPTS[transcript],
                                                C.byref(ok)))
        return bool(ok.value)

    @staticmethod
    def verify_batch(vk_bytes: bytes, public_inputs_list, proofs, seed: bytes) -> bool:
        """Many proofs under one key with ONE product of three pairings (random linear combination)."""
        lib = load()
        _bind(lib)
        if len(public_inputs_list) != len(proofs):
            raise ValueError("one public-input list per proof")
        k = len(public_inputs_list[0]) if proofs else 0
        if any(len(p) != k for p in public_inputs_list) or any(len(p) != 176 for p in proofs) or len(seed) != 32:
            raise ValueError("ragged public inputs / bad proof or seed length")
        ok = C.c_int(0)
        pub = b"".join(_checked_frs(p, "public input") for p in public_inputs_list)
        check(lib.pm_polymath_verify_batch(vk_bytes, len(proofs), pub, k, b"".join(proofs), seed, C.byref(ok)))
        return bool(ok.value)
