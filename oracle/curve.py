"""BLS12-381 G1/G2 group law and arkworks/zcash point encodings (oracle; test infrastructure).

Restates ark-ec 0.4 short-Weierstrass groups as used by the reference at
`src/generator.rs:169-177` (`g * scalar` then `.into()` affine),
`src/prover.rs:380-384` (`msm_unchecked`) and `src/verifier.rs:44-61`.
Only affine images are observable, so the coordinate system is free.

Affine points are `(x, y)` tuples of canonical ints or `None` for infinity.
Jacobian points are `(X, Y, Z)` with Z == 0 for infinity.
"""
from .fields import (
    Q_MOD, R_MOD, G1_GEN_X, G1_GEN_Y, G2_GEN_X, G2_GEN_Y,
    fq_inv, fq2_add, fq2_sub, fq2_mul, fq2_neg, fq2_inv, fq2_scalar, FQ2_ZERO, FQ2_ONE,
)

P = Q_MOD
G1_GEN = (G1_GEN_X, G1_GEN_Y)
G2_GEN = (G2_GEN_X, G2_GEN_Y)
JINF = (1, 1, 0)


def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - 4) % P == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def jac_double(p1):
    X1, Y1, Z1 = p1
    if Z1 == 0:
        return JINF
    # dbl-2009-l (a = 0)
    A = X1 * X1 % P
    B = Y1 * Y1 % P
    C = B * B % P
    D = 2 * ((X1 + B) * (X1 + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y1 * Z1 % P
    return (X3, Y3, Z3)


def jac_add(p1, p2):
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == 0:
        return p2
    if Z2 == 0:
        return p1
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        if S1 == S2:
            return jac_double(p1)
        return JINF
    H = (U2 - U1) % P
    Rr = (S2 - S1) % P
    HH = H * H % P
    HHH = H * HH % P
    V = U1 * HH % P
    X3 = (Rr * Rr - HHH - 2 * V) % P
    Y3 = (Rr * (V - X3) - S1 * HHH) % P
    Z3 = Z1 * Z2 * H % P
    return (X3, Y3, Z3)


def jac_add_affine(p1, q):
    """Mixed addition p1 (Jacobian) + q (affine or None)."""
    if q is None:
        return p1
    X1, Y1, Z1 = p1
    if Z1 == 0:
        return (q[0], q[1], 1)
    x2, y2 = q
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    if U2 == X1:
        if S2 == Y1:
            return jac_double(p1)
        return JINF
    H = (U2 - X1) % P
    Rr = (S2 - Y1) % P
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (Rr * Rr - HHH - 2 * V) % P
    Y3 = (Rr * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def jac_to_affine(p1):
    X, Y, Z = p1
    if Z == 0:
        return None
    zi = fq_inv(Z)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def jac_batch_to_affine(pts):
    """Montgomery batch inversion; returns list of affine points / None."""
    zs = [p[2] for p in pts]
    prefix = []
    acc = 1
    for z in zs:
        prefix.append(acc)
        if z:
            acc = acc * z % P
    inv = fq_inv(acc)
    out = [None] * len(pts)
    for i in range(len(pts) - 1, -1, -1):
        z = zs[i]
        if z == 0:
            continue
        zi = inv * prefix[i] % P
        inv = inv * z % P
        zi2 = zi * zi % P
        out[i] = (pts[i][0] * zi2 % P, pts[i][1] * zi2 * zi % P)
    return out


def g1_mul(pt, k: int):
    """Double-and-add `pt * k`, the literal `g * scalar` of generator.rs:175; returns affine."""
    k %= R_MOD
    if pt is None or k == 0:
        return None
    acc = JINF
    for bit in bin(k)[2:]:
        acc = jac_double(acc)
        if bit == "1":
            acc = jac_add_affine(acc, pt)
    return jac_to_affine(acc)


def g1_add(a, b):
    if a is None:
        return b
    return jac_to_affine(jac_add_affine((a[0], a[1], 1), b))


class FixedBaseTable:
    """Windowed fixed-base table of G multiples for fast [s]G (oracle speed-up; result is canonical)."""

    def __init__(self, base=G1_GEN, window=8, bits=255):
        self.window = window
        self.nwin = (bits + window - 1) // window
        self.tables = []
        cur = (base[0], base[1], 1)
        for _ in range(self.nwin):
            row_j = [JINF]
            acc = JINF
            for _ in range((1 << window) - 1):
                acc = jac_add(acc, cur)
                row_j.append(acc)
            self.tables.append(jac_batch_to_affine(row_j))
            for _ in range(window):
                cur = jac_double(cur)

    def mul_jac(self, k: int):
        k %= R_MOD
        acc = JINF
        mask = (1 << self.window) - 1
        for w in range(self.nwin):
            d = (k >> (w * self.window)) & mask
            if d:
                acc = jac_add_affine(acc, self.tables[w][d])
        return acc

    def mul_many(self, ks):
        return jac_batch_to_affine([self.mul_jac(k) for k in ks])


# ---------------------------------------------------------------------------
# G2 (affine over Fq2), only what setup's vk and the verifier need
# ---------------------------------------------------------------------------

G2_B = (4, 4)


def g2_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    lhs = fq2_mul(y, y)
    rhs = fq2_add(fq2_mul(fq2_mul(x, x), x), G2_B)
    return lhs == rhs


def g2_neg(pt):
    if pt is None:
        return None
    return (pt[0], fq2_neg(pt[1]))


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if y1 == y2:
            if y1 == FQ2_ZERO:
                return None
            lam = fq2_mul(fq2_scalar(fq2_mul(x1, x1), 3), fq2_inv(fq2_scalar(y1, 2)))
        else:
            return None
    else:
        lam = fq2_mul(fq2_sub(y2, y1), fq2_inv(fq2_sub(x2, x1)))
    x3 = fq2_sub(fq2_sub(fq2_mul(lam, lam), x1), x2)
    y3 = fq2_sub(fq2_mul(lam, fq2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(pt, k: int):
    k %= R_MOD
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


# ---------------------------------------------------------------------------
# Encodings.  ark-bls12-381 0.4 serialises G1/G2 in the zcash format
# (SURVEY.md A.4): big-endian x, flag bits in the top byte:
# 0x80 compressed, 0x40 infinity, 0x20 y is the lexicographically larger root.
# ---------------------------------------------------------------------------

def g1_compress(pt) -> bytes:
    if pt is None:
        b = bytearray(48)
        b[0] = 0xC0
        return bytes(b)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def g1_decompress(b: bytes, validate: bool = True):
    """ark-bls12-381 `read_g1_compressed` + `get_point_from_x_unchecked` [mem] (third-party crate, not under
    /root/reference; call sites: derive(CanonicalDeserialize) on src/data_structures.rs:10-73).  Raises ValueError
    where arkworks returns a SerializationError.  validate=True adds `deserialize_compressed`'s subgroup check."""
    assert len(b) == 48
    flags = b[0]
    if not flags & 0x80:
        raise ValueError("UnexpectedFlags: not a compressed encoding")
    if flags & 0x40:
        # arkworks algebra master (the snapshot the reference patches to, Cargo.toml:84-101) rejects an infinity
        # encoding that carries the sort flag or a non-zero x; ark-bls12-381 0.4.0 returned zero right away
        if flags != 0xC0 or any(b[1:]):
            raise ValueError("InvalidData: infinity flag on a non-zero encoding")
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P:
        raise ValueError("InvalidData: x >= q")
    rhs = (x * x * x + 4) % P
    y = pow(rhs, (P + 1) // 4, P)
    if y * y % P != rhs:
        raise ValueError("InvalidData: not on the curve")
    small, large = (y, P - y) if y < P - y else (P - y, y)
    pt = (x, large if flags & 0x20 else small)
    if validate and not g1_in_subgroup(pt):
        raise ValueError("InvalidData: not in the prime-order subgroup")
    return pt


def g1_in_subgroup(pt) -> bool:
    """[r]P == O (what `is_in_correct_subgroup_assuming_on_curve` decides), by the plain ladder (no reduction of r)."""
    if pt is None:
        return True
    acc = JINF
    for bit in bin(R_MOD)[2:]:
        acc = jac_double(acc)
        if bit == "1":
            acc = jac_add_affine(acc, pt)
    return acc[2] == 0


def _fq2_lex_largest(y) -> bool:
    # zcash: compare c1 first, then c0
    half = (P - 1) // 2
    if y[1] != 0:
        return y[1] > half
    return y[0] > half


def g2_compress(pt) -> bytes:
    if pt is None:
        b = bytearray(96)
        b[0] = 0xC0
        return bytes(b)
    x, y = pt
    b = bytearray(x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if _fq2_lex_largest(y):
        b[0] |= 0x20
    return bytes(b)


def fr_to_bytes(v: int) -> bytes:
    """ark-serialize Fp: 32 bytes little-endian canonical integer."""
    return (v % R_MOD).to_bytes(32, "little")
