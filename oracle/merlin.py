"""Merlin 3.0 transcript (STROBE-128 over Keccak-f[1600]) and the reference's field challenge rule.

Restates `/root/reference/src/transcript/merlin.rs:13-36` over the published
merlin 3.0.0 / STROBE v1.0.2 construction.  Oracle; test
infrastructure.  Pinned by the merlin "test protocol" vector and by a
sha3_256 cross-check of the permutation (tests/test_oracle.py).
"""
from .fields import R_MOD

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]
_M = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M if n else v


def keccak_f1600(state: bytearray) -> None:
    a = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")


FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32
STROBE_R = 166


class Strobe128:
    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, STROBE_R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state = st
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[STROBE_R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data: bytes):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool):
        if more:
            assert self.cur_flags == flags
            return
        assert not (flags & FLAG_T)
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if (flags & (FLAG_C | FLAG_K)) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin_op(FLAG_M | FLAG_A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin_op(FLAG_A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(FLAG_I | FLAG_A | FLAG_C, more)
        return self._squeeze(n)


class MerlinTranscript:
    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)


def fr_from_random_bytes(buf: bytes):
    """ark-ff 0.4 `Fp::from_random_bytes`: first 32 bytes LE, clear bit 255, accept iff < r."""
    v = int.from_bytes(buf[:32], "little") & ((1 << 255) - 1)
    return v if v < R_MOD else None


class MerlinFieldTranscript:
    """`MerlinFieldTranscript<Fr>` of transcript/merlin.rs:13-36."""

    def __init__(self, name: bytes):
        self.merlin = MerlinTranscript(name)

    def append_message(self, label: bytes, message: bytes):
        self.merlin.append_message(label, message)

    def challenge(self, label: bytes) -> int:
        while True:
            buf = self.merlin.challenge_bytes(label, 64)
            r = fr_from_random_bytes(buf)
            if r is not None:
                return r


# ---------------------------------------------------------------------------
# The reference's two hash transcripts (transcript/keccak256.rs:16-42, transcript/blake3.rs:16-42)
# ---------------------------------------------------------------------------

def keccak256(data: bytes) -> bytes:
    """Keccak-256 with the ORIGINAL padding (0x01 ... 0x80), the `sha3` crate's `Keccak256` (not NIST SHA3-256)."""
    rate = 136
    st = bytearray(200)
    padded = bytearray(data) + b"\x01"
    padded += bytes((-len(padded)) % rate)
    padded[-1] ^= 0x80
    for off in range(0, len(padded), rate):
        for i in range(rate):
            st[i] ^= padded[off + i]
        keccak_f1600(st)
    return bytes(st[:32])


def blake3_hash(data: bytes) -> bytes:
    import blake3          # third-party module: an implementation independent of the product's C++ one
    return blake3.blake3(data).digest()


class _HashFieldTranscript:
    """`new` ignores the name; messages are appended as label || message; a challenge is H(transcript || label) as a
    big-endian integer mod r (`from_be_bytes_mod_order`), after which the transcript is the digest."""
    hash_fn = None

    def __init__(self, name: bytes):
        self.transcript = b""

    def append_message(self, label: bytes, message: bytes):
        self.transcript += label + message

    def challenge(self, label: bytes) -> int:
        buf = type(self).hash_fn(self.transcript + label)
        self.transcript = buf
        return int.from_bytes(buf, "big") % R_MOD


class Keccak256FieldTranscript(_HashFieldTranscript):
    hash_fn = staticmethod(keccak256)


class Blake3FieldTranscript(_HashFieldTranscript):
    hash_fn = staticmethod(blake3_hash)


TRANSCRIPTS = {"merlin": MerlinFieldTranscript, "keccak256": Keccak256FieldTranscript, "blake3": Blake3FieldTranscript}

