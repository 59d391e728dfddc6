"""BLS12-381 field constants and tower arithmetic on Python integers (oracle; test infrastructure).

Restates what the reference gets from ark-bls12-381 / ark-ff 0.4 (un-vendored
dependency; `/root/reference/Cargo.toml:14,35`).  Field elements are plain
canonical integers in [0, p); the Montgomery form arkworks keeps in memory
(R = 2^256 for Fr, 2^384 for Fq, 64-bit LE limbs) is produced by `to_mont` /
`from_mont` at the byte boundary only.
"""

# Scalar field Fr (255 bits) and base field Fq (381 bits).
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB

FR_BITS = 255
FQ_BITS = 381
FR_MONT_R = (1 << 256) % R_MOD
FQ_MONT_R = (1 << 384) % Q_MOD
FR_MONT_RINV = pow(FR_MONT_R, -1, R_MOD)
FQ_MONT_RINV = pow(FQ_MONT_R, -1, Q_MOD)

# ark-bls12-381 FrConfig: GENERATOR = 7, TWO_ADICITY = 32.
FR_GENERATOR = 7
FR_TWO_ADICITY = 32
FR_TWO_ADIC_ROOT = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)

# BLS parameter x (negative): the Miller loop runs over |x|.
BLS_X_ABS = 0xD201000000010000

# G1 generator (affine), curve y^2 = x^3 + 4.
G1_GEN_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_GEN_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G1_B = 4

# G2 generator (affine over Fq2 = Fq[u]/(u^2+1)), twist y^2 = x^3 + 4(1+u).
G2_GEN_X = (
    0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
    0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
)
G2_GEN_Y = (
    0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
    0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
)


def to_mont(v: int, p: int, rbits: int) -> int:
    return (v << rbits) % p


def from_mont(v: int, p: int, rbits: int) -> int:
    return (v * pow(1 << rbits, -1, p)) % p


def fr_inv(a: int) -> int:
    return pow(a, -1, R_MOD)


def fq_inv(a: int) -> int:
    return pow(a, -1, Q_MOD)


# ---------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2 + 1), as tuples (c0, c1)
# ---------------------------------------------------------------------------

def fq2_add(a, b):
    return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)


def fq2_neg(a):
    return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)


def fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)


def fq2_scalar(a, k):
    return ((a[0] * k) % Q_MOD, (a[1] * k) % Q_MOD)


def fq2_inv(a):
    d = fq_inv((a[0] * a[0] + a[1] * a[1]) % Q_MOD)
    return ((a[0] * d) % Q_MOD, (-a[1] * d) % Q_MOD)


FQ2_ZERO = (0, 0)
FQ2_ONE = (1, 0)

# ---------------------------------------------------------------------------
# Fq12 = Fq[w]/(w^12 - 2 w^6 + 2): with w^6 = xi = 1 + u we have u = w^6 - 1 and
# u^2 = -1  <=>  w^12 - 2 w^6 + 2 = 0.  Elements are 12-coefficient lists.
# ---------------------------------------------------------------------------

FQ12_ONE = [1] + [0] * 11


def fq12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    # reduce: w^12 = 2 w^6 - 2
    for k in range(22, 11, -1):
        c = t[k]
        if c:
            t[k - 6] += 2 * c
            t[k - 12] -= 2 * c
    return [x % Q_MOD for x in t[:12]]


def fq12_pow(a, e: int):
    res = list(FQ12_ONE)
    base = list(a)
    while e:
        if e & 1:
            res = fq12_mul(res, base)
        base = fq12_mul(base, base)
        e >>= 1
    return res


def fq2_to_fq12(a):
    """a0 + a1*u -> (a0 - a1) + a1*w^6."""
    out = [0] * 12
    out[0] = (a[0] - a[1]) % Q_MOD
    out[6] = a[1] % Q_MOD
    return out
