"""Byte codecs between Python integers and the library's wire forms (Montgomery LE limbs)."""
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_FR_R = (1 << 256) % R_MOD
_FQ_R = (1 << 384) % Q_MOD
_FR_RI = pow(_FR_R, -1, R_MOD)
_FQ_RI = pow(_FQ_R, -1, Q_MOD)


def fr_to_wire(v):
    return ((v % R_MOD) * _FR_R % R_MOD).to_bytes(32, "little")


def fr_from_wire(b):
    return int.from_bytes(b, "little") * _FR_RI % R_MOD


def fq_to_wire(v):
    return ((v % Q_MOD) * _FQ_R % Q_MOD).to_bytes(48, "little")


def fq_from_wire(b):
    return int.from_bytes(b, "little") * _FQ_RI % Q_MOD


def frs_to_wire(vals):
    return b"".join(fr_to_wire(v) for v in vals)


def frs_from_wire(buf):
    return [fr_from_wire(buf[i:i + 32]) for i in range(0, len(buf), 32)]


def g1_to_wire(pt):
    """Affine (x, y) ints or None -> 96 bytes ((0,0) = infinity)."""
    if pt is None:
        return bytes(96)
    return fq_to_wire(pt[0]) + fq_to_wire(pt[1])


def g1_from_wire(b):
    if b == bytes(96):
        return None
    return (fq_from_wire(b[:48]), fq_from_wire(b[48:96]))


def g1s_to_wire(pts):
    return b"".join(g1_to_wire(p) for p in pts)


def g1s_from_wire(buf):
    return [g1_from_wire(buf[i:i + 96]) for i in range(0, len(buf), 96)]
