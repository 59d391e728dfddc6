"""Radix-2 domain, NTT and MSM restated on Python integers (oracle; test infrastructure).

Restates ark-poly 0.4 `Radix2EvaluationDomain` as called at
`/root/reference/src/prover.rs:83,241,319,325` and `src/generator.rs:61,72,77,106,113`
and ark-ec 0.4 `VariableBaseMSM::msm_unchecked` (`src/prover.rs:383`), per
SURVEY.md A.2/A.6.  Outputs are canonical, so algorithm choice is free; a naive
O(n^2) DFT and a naive double-and-add MSM are kept beside the fast versions as
cross-checks.
"""
from .fields import R_MOD, FR_TWO_ADIC_ROOT, FR_TWO_ADICITY, fr_inv
from .curve import JINF, jac_add, jac_add_affine, jac_double, jac_to_affine, g1_mul, g1_add

P = R_MOD


def next_pow2(k: int) -> int:
    n = 1
    while n < k:
        n <<= 1
    return n


class Domain:
    """`Radix2EvaluationDomain::new(k)`: size = next power of two, offset 1."""

    def __init__(self, num_coeffs: int):
        self.size = next_pow2(num_coeffs)
        self.log_size = self.size.bit_length() - 1
        assert self.log_size <= FR_TWO_ADICITY
        g = FR_TWO_ADIC_ROOT
        for _ in range(self.log_size, FR_TWO_ADICITY):
            g = g * g % P
        self.group_gen = g
        self.group_gen_inv = fr_inv(g)
        self.size_inv = fr_inv(self.size)

    def evaluate_vanishing_polynomial(self, tau: int) -> int:
        return (pow(tau, self.size, P) - 1) % P

    def elements(self):
        out, w = [], 1
        for _ in range(self.size):
            out.append(w)
            w = w * self.group_gen % P
        return out

    def evaluate_all_lagrange_coefficients(self, tau: int):
        """L_i(tau) = Z_H(tau) * w^i / (n * (tau - w^i)) for tau outside H (generator.rs:113)."""
        zh = self.evaluate_vanishing_polynomial(tau)
        assert zh != 0
        pref = zh * self.size_inv % P
        return [pref * w % P * fr_inv((tau - w) % P) % P for w in self.elements()]

    def sample_element_outside_domain(self, rng, fr_rand):
        t = fr_rand(rng)
        while self.evaluate_vanishing_polynomial(t) == 0:
            t = fr_rand(rng)
        return t

    def fft(self, coeffs):
        a = list(coeffs) + [0] * (self.size - len(coeffs))
        assert len(a) == self.size
        return ntt(a, self.group_gen)

    def ifft(self, evals):
        a = list(evals) + [0] * (self.size - len(evals))
        assert len(a) == self.size
        out = ntt(a, self.group_gen_inv)
        return [v * self.size_inv % P for v in out]


def ntt(a, root):
    """In-order iterative radix-2 Cooley-Tukey; len(a) must be a power of two."""
    n = len(a)
    a = list(a)
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    length = 2
    while length <= n:
        wlen = pow(root, n // length, P)
        half = length >> 1
        tw = [1] * half
        for k in range(1, half):
            tw[k] = tw[k - 1] * wlen % P
        for start in range(0, n, length):
            for k in range(half):
                u = a[start + k]
                v = a[start + k + half] * tw[k] % P
                a[start + k] = (u + v) % P
                a[start + k + half] = (u - v) % P
        length <<= 1
    return a


def naive_dft(a, root):
    n = len(a)
    return [sum(a[j] * pow(root, i * j, P) for j in range(n)) % P for i in range(n)]


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def strip(coeffs):
    c = list(coeffs)
    while c and c[-1] == 0:
        c.pop()
    return c


# ---------------------------------------------------------------------------
# MSM
# ---------------------------------------------------------------------------

def msm_naive(scalars, bases):
    acc = None
    for s, b in zip(scalars, bases):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


def ark_window_bits(n: int) -> int:
    """ark-ec 0.4 window rule: 3 if n < 32 else ln(n)*69/100 + 2 with ln = ceil(log2)."""
    if n < 32:
        return 3
    log2 = (n - 1).bit_length()
    return log2 * 69 // 100 + 2


def msm_pippenger(scalars, bases, c=None):
    """Unsigned-window Pippenger; returns the affine sum.  `msm_unchecked` truncates to the shorter input."""
    k = min(len(scalars), len(bases))
    scalars = [s % P for s in scalars[:k]]
    bases = bases[:k]
    if c is None:
        c = ark_window_bits(k)
    nwin = (255 + c - 1) // c
    window_sums = []
    for w in range(nwin):
        buckets = [JINF] * ((1 << c) - 1)
        sh = w * c
        mask = (1 << c) - 1
        for s, b in zip(scalars, bases):
            d = (s >> sh) & mask
            if d and b is not None:
                buckets[d - 1] = jac_add_affine(buckets[d - 1], b)
        running, acc = JINF, JINF
        for bk in reversed(buckets):
            running = jac_add(running, bk)
            acc = jac_add(acc, running)
        window_sums.append(acc)
    total = JINF
    for ws in reversed(window_sums):
        for _ in range(c):
            total = jac_double(total)
        total = jac_add(total, ws)
    return jac_to_affine(total)
