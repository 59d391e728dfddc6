"""Polymath setup / prove / verify restated (oracle; test infrastructure).

Follows `/root/reference/src/generator.rs:24-177`, `src/prover.rs:27-384`,
`src/verifier.rs:19-62` and `src/common.rs:21-230` line by line.  Two SAP
evaluation modes are kept:

* ``literal`` — the reference's own dense data flow through the virtual
  `u(i,j)` / `w(i,j)` accessors (`common.rs:138-207`, `prover.rs:87-96`), O(n*m),
  usable only at toy sizes;
* ``sparse`` — the closed form of SURVEY.md §8(a3) (three R1CS SpMVs plus
  element-wise terms), which is what the device path computes.

tests/test_oracle.py checks both modes agree and that `verify` accepts.
PARITY UNPINNED (see oracle/__init__.py).
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from .fields import R_MOD, fr_inv
from .curve import (
    G1_GEN, G2_GEN, FixedBaseTable, g1_add, g1_mul, g1_neg, g2_add, g2_mul, g2_neg,
    g1_compress, g2_compress, fr_to_bytes,
)
from .poly import Domain, poly_eval, strip, msm_pippenger, msm_naive
from .merlin import MerlinFieldTranscript
from .rng import fr_rand
from .pairing import pairing_product_is_one
from .r1cs import synthesize

P = R_MOD
B_POLYMATH = b"polymath"      # common.rs:8
MINUS_ALPHA = 3               # common.rs:11
MINUS_GAMMA = 5               # common.rs:14


@dataclass
class SAPMatrices:            # common.rs:113-127
    num_instance_variables: int
    num_r1cs_witness_variables: int
    num_r1cs_constraints: int
    a: list
    b: list
    c: list

    def m0_m_n(self):         # common.rs:224-229
        m0 = self.num_instance_variables
        return m0, m0 + self.num_r1cs_witness_variables, self.num_r1cs_constraints

    def size(self):           # common.rs:131-135
        m0, m, n = self.m0_m_n()
        return (m0 + n) * 2, m0 * 2 + m + n

    def u(self, i, j):        # common.rs:138-173
        m0, m, n = self.m0_m_n()
        dm0, dm0n, dm0dn, m0m = m0 + m0, m0 + m0 + n, m0 + m0 + n + n, m0 + m
        if (i, j) == (0, 0):
            return 2
        if i < m0 and j == 0:
            return 1
        if i < m0 and j == i:
            return 1
        if i < m0:
            return 0
        if i == m0 and j == 0:
            return 0
        if i < dm0 and j == 0:
            return 1
        if i < dm0 and j == i - m0:
            return P - 1
        if i < dm0:
            return 0
        if j < m0:
            return 0
        if i < dm0n and j < m0m:
            return (m_at(self.a, i - dm0, j - m0) + m_at(self.b, i - dm0, j - m0)) % P
        if i < dm0dn and j < m0m:
            return (m_at(self.a, i - dm0n, j - m0) - m_at(self.b, i - dm0n, j - m0)) % P
        return 0

    def w(self, i, j):        # common.rs:176-207
        m0, m, n = self.m0_m_n()
        dm0, dm0n, dm0dn, m0m = m0 + m0, m0 + m0 + n, m0 + m0 + n + n, m0 + m
        if i < m0 and j == i + m0:
            return 4
        if i < m0 and j == i + m0m:
            return 1
        if i < m0:
            return 0
        if i < dm0 and j == i + m:
            return 1
        if i < dm0:
            return 0
        if j < m0:
            return 0
        if i < dm0n and j < m0m:
            return m_at(self.c, i - dm0, j - m0) * 4 % P
        if i < dm0n and j == i + m:
            return 1
        if i < dm0n:
            return 0
        if i < dm0dn and j == i - n + m:
            return 1
        return 0


def m_at(mat, i, j):          # common.rs:100-105 (first match wins)
    for coeff, idx in mat[i]:
        if idx == j:
            return coeff
    return 0


@dataclass
class VerifyingKey:           # data_structures.rs:25-50
    one_g1: tuple
    one_g2: tuple
    x_g2: tuple
    z_g2: tuple
    n: int
    m0: int
    sigma: int
    omega: int

    def serialize_compressed(self) -> bytes:
        return (g1_compress(self.one_g1) + g2_compress(self.one_g2) + g2_compress(self.x_g2)
                + g2_compress(self.z_g2) + self.n.to_bytes(8, "little") + self.m0.to_bytes(8, "little")
                + self.sigma.to_bytes(8, "little") + fr_to_bytes(self.omega))


@dataclass
class ProvingKey:             # data_structures.rs:56-73
    vk: VerifyingKey
    sap_matrices: SAPMatrices
    x_powers_g1: list
    x_powers_y_alpha_g1: list
    x_powers_zh_by_y_alpha_g1: list
    x_powers_y_gamma_g1: list
    x_powers_y_gamma_z_g1: list
    uj_wj_lcs_by_y_alpha_g1: list
    trapdoor: dict = field(default_factory=dict)   # oracle-only debugging aid (x, z); never serialised


@dataclass
class Proof:                  # data_structures.rs:10-19
    a_g1: Optional[tuple]
    c_g1: Optional[tuple]
    a_at_x1: int
    d_g1: Optional[tuple]

    def serialize_compressed(self) -> bytes:
        return g1_compress(self.a_g1) + g1_compress(self.c_g1) + fr_to_bytes(self.a_at_x1) + g1_compress(self.d_g1)


# ---------------------------------------------------------------------------
# common.rs Fiat-Shamir helpers
# ---------------------------------------------------------------------------

def ser_fr_slice(vals) -> bytes:          # ark-serialize &[F]: u64-LE length prefix + elements
    return len(vals).to_bytes(8, "little") + b"".join(fr_to_bytes(v) for v in vals)


def ser_g1_slice(pts) -> bytes:
    return len(pts).to_bytes(8, "little") + b"".join(g1_compress(p) for p in pts)


def compute_x1(t, public_inputs, commitments):        # common.rs:21-30
    t.append_message(b"public_inputs", ser_fr_slice(public_inputs))
    t.append_message(b"commitments", ser_g1_slice(commitments))
    return t.challenge(b"x1")


def compute_x2(t, x1, values):                        # common.rs:32-37
    t.append_message(b"x1", fr_to_bytes(x1))
    t.append_message(b"values", ser_fr_slice(values))
    return t.challenge(b"x2")


def compute_y1(x1, sigma):                            # common.rs:40-42
    return pow(x1, sigma, P)


def neg_power(y, minus_exp):                          # common.rs:45-47
    return pow(fr_inv(y), minus_exp, P)


def z_tilde_i(public_inputs, i):                      # common.rs:77-97
    m0 = len(public_inputs)
    if i == 0:
        return 2
    if i < m0:
        return (1 + public_inputs[i]) % P
    if i == m0:
        return 0
    return (1 - public_inputs[i - m0]) % P


def compute_pi_at_x1(vk, public_inputs, x1, y1_gamma):  # common.rs:49-71
    total = 0
    num = (pow(x1, vk.n, P) - 1) * fr_inv(vk.n % P) % P
    omega_i = 1
    m0 = len(public_inputs)
    for i in range(2 * m0):
        lagrange = num * fr_inv((x1 - omega_i) % P) % P
        total = (total + z_tilde_i(public_inputs, i) * lagrange) % P
        num = num * vk.omega % P
        omega_i = omega_i * vk.omega % P
    return total * y1_gamma % P


def compute_c_at_x1(y1_gamma, y1_alpha, a_at_x1, pi_at_x1):  # common.rs:73-75
    return ((a_at_x1 + y1_gamma) * a_at_x1 - pi_at_x1) % P * fr_inv(y1_alpha) % P


# ---------------------------------------------------------------------------
# SAP evaluation: literal (dense accessors) and sparse (closed form)
# ---------------------------------------------------------------------------

def spmv(mat, vec):
    return [sum(c * vec[j] for c, j in row) % P for row in mat]


def compute_y_vec(sap, x, w, literal=False):          # prover.rs:279-302
    m0 = sap.num_instance_variables
    y_m0 = [pow(1 - x[j], 2, P) for j in range(1, m0)]
    zp = list(x) + list(w)
    if literal:
        y_n = []
        for i in range(sap.num_r1cs_constraints):
            v = sum((m_at(sap.a, i, j) - m_at(sap.b, i, j)) * zp[j] for j in range(len(zp))) % P
            y_n.append(v * v % P)
    else:
        az, bz = spmv(sap.a, zp), spmv(sap.b, zp)
        y_n = [pow(a_ - b_, 2, P) for a_, b_ in zip(az, bz)]
    return [0] + y_m0 + y_n


def sap_evals_sparse(sap, x, w, y, n):
    """u_evals = U.z, w_evals = W.z by the closed form of SURVEY.md §8(a3)."""
    m0, nr = sap.num_instance_variables, sap.num_r1cs_constraints
    zp = list(x) + list(w)
    az, bz, cz = spmv(sap.a, zp), spmv(sap.b, zp), spmv(sap.c, zp)
    u = [0] * n
    wv = [0] * n
    for i in range(m0):
        u[i] = (1 + x[i]) % P
        wv[i] = (4 * x[i] + y[i]) % P
        u[m0 + i] = (1 - x[i]) % P
        wv[m0 + i] = y[i]
    for r in range(nr):
        u[2 * m0 + r] = (az[r] + bz[r]) % P
        wv[2 * m0 + r] = (4 * cz[r] + y[m0 + r]) % P
        u[2 * m0 + nr + r] = (az[r] - bz[r]) % P
        wv[2 * m0 + nr + r] = y[m0 + r]
    wu = [0] * (2 * m0) + u[2 * m0:]      # witness-column part of U.z (prover.rs:156-161)
    return u, wv, wu, list(wv)


def sap_evals_literal(sap, z, n, m0):
    """prover.rs:87-96,156-166 verbatim: dense U, W, column scaling, vector sums."""
    rows, cols = sap.size()
    zc = [v for part in z for v in part]
    assert len(zc) == cols
    uj = [[sap.u(i, j) for i in range(n)] for j in range(cols)]
    wj = [[sap.w(i, j) for i in range(n)] for j in range(cols)]
    ujz = [[c * zc[j] % P for c in col] for j, col in enumerate(uj)]
    wjz = [[c * zc[j] % P for c in col] for j, col in enumerate(wj)]

    def sum_vectors(vs):
        return [sum(col[i] for col in vs) % P for i in range(n)]

    return sum_vectors(ujz), sum_vectors(wjz), sum_vectors(ujz[m0:]), sum_vectors(wjz[m0:])


# ---------------------------------------------------------------------------
# setup  (generator.rs:24-177)
# ---------------------------------------------------------------------------

def lcs_scalars_sparse(sap, lag, y_gamma, y_to_minus_alpha):
    """(u_j(x)*y^gamma + w_j(x)) * y^-alpha for j >= m0 via transposed SpMV."""
    m0, m, nr = sap.m0_m_n()
    ncols_r1cs = m
    ua = [0] * ncols_r1cs
    wa = [0] * ncols_r1cs
    for r in range(nr):
        l1, l2 = lag[2 * m0 + r], lag[2 * m0 + nr + r]
        seen = set()
        for cf, k in sap.a[r]:
            if k in seen:
                continue
            seen.add(k)
            ua[k] = (ua[k] + cf * (l1 + l2)) % P
        seen = set()
        for cf, k in sap.b[r]:
            if k in seen:
                continue
            seen.add(k)
            ua[k] = (ua[k] + cf * (l1 - l2)) % P
        seen = set()
        for cf, k in sap.c[r]:
            if k in seen:
                continue
            seen.add(k)
            wa[k] = (wa[k] + 4 * cf * l1) % P
    for k in range(m0):
        wa[k] = (wa[k] + 4 * lag[k]) % P
    out = [(ua[k] * y_gamma + wa[k]) % P * y_to_minus_alpha % P for k in range(ncols_r1cs)]
    for t in range(m0):
        out.append((lag[t] + lag[m0 + t]) % P * y_to_minus_alpha % P)
    for r in range(nr):
        out.append((lag[2 * m0 + r] + lag[2 * m0 + nr + r]) % P * y_to_minus_alpha % P)
    return out


def lcs_scalars_literal(sap, lag, n, m, m0, y_gamma, y_to_minus_alpha):   # generator.rs:115-135
    out = []
    for j in range(m - m0):
        uj_x = sum(lag[i] * sap.u(i, j + m0) for i in range(n)) % P
        wj_x = sum(lag[i] * sap.w(i, j + m0) for i in range(n)) % P
        out.append((uj_x * y_gamma + wj_x) % P * y_to_minus_alpha % P)
    return out


def setup_scalars(sap, x, z, literal=False):
    """All scalar vectors whose [.]G the generator publishes, in ProvingKey field order of generation."""
    rows, cols = sap.size()
    domain = Domain(rows)
    n, m, m0 = domain.size, cols, sap.num_instance_variables
    sigma = n + 3
    y = pow(x, sigma, P)
    y_alpha = pow(fr_inv(y), MINUS_ALPHA, P)
    y_to_minus_alpha = pow(y, MINUS_ALPHA, P)
    y_gamma = pow(fr_inv(y), MINUS_GAMMA, P)
    zh = domain.evaluate_vanishing_polynomial(x)

    def powers(count, mult):
        out, cur = [], mult % P
        for _ in range(count):
            out.append(cur)
            cur = cur * x % P
        return out

    d_max = 2 * (n - 1) + sigma * (MINUS_ALPHA + MINUS_GAMMA)
    lag = domain.evaluate_all_lagrange_coefficients(x)
    if literal:
        lcs = lcs_scalars_literal(sap, lag, n, m, m0, y_gamma, y_to_minus_alpha)
    else:
        lcs = lcs_scalars_sparse(sap, lag, y_gamma, y_to_minus_alpha)
    return dict(
        n=n, m=m, m0=m0, sigma=sigma, omega=domain.group_gen,
        x_powers=powers(n + 1, 1),                                   # generator.rs:82
        x_powers_y_alpha=powers(3, y_alpha),                         # :86
        x_powers_y_gamma=powers(2, y_gamma),                         # :90
        x_powers_y_gamma_z=powers(d_max + 1, y_gamma * z),           # :94-100
        x_powers_zh_by_y_alpha=powers(n - 1, zh * y_to_minus_alpha),  # :105-108
        uj_wj_lcs_by_y_alpha=lcs,                                    # :112-135
    )


_FB_TABLE = None


def _fixed_base():
    global _FB_TABLE
    if _FB_TABLE is None:
        _FB_TABLE = FixedBaseTable(G1_GEN, window=8)
    return _FB_TABLE


def generate_proving_key(circuit, rng, literal=False) -> ProvingKey:
    cs = synthesize(circuit, setup_mode=True)
    a, b, c = cs.to_matrices()
    sap = SAPMatrices(cs.num_instance_variables, cs.num_witness_variables, cs.num_constraints, a, b, c)
    rows, _ = sap.size()
    domain = Domain(rows)
    x = domain.sample_element_outside_domain(rng, fr_rand)   # generator.rs:72
    # generator.rs:73-76 consume no randomness
    z = domain.sample_element_outside_domain(rng, fr_rand)   # generator.rs:77
    sc = setup_scalars(sap, x, z, literal=literal)
    fb = _fixed_base()
    vk = VerifyingKey(
        one_g1=G1_GEN, one_g2=G2_GEN, x_g2=g2_mul(G2_GEN, x), z_g2=g2_mul(G2_GEN, z),
        n=sc["n"], m0=sc["m0"], sigma=sc["sigma"], omega=sc["omega"],
    )
    return ProvingKey(
        vk=vk, sap_matrices=sap,
        x_powers_g1=fb.mul_many(sc["x_powers"]),
        x_powers_y_alpha_g1=fb.mul_many(sc["x_powers_y_alpha"]),
        x_powers_zh_by_y_alpha_g1=fb.mul_many(sc["x_powers_zh_by_y_alpha"]),
        x_powers_y_gamma_g1=fb.mul_many(sc["x_powers_y_gamma"]),
        x_powers_y_gamma_z_g1=fb.mul_many(sc["x_powers_y_gamma_z"]),
        uj_wj_lcs_by_y_alpha_g1=fb.mul_many(sc["uj_wj_lcs_by_y_alpha"]),
        trapdoor=dict(x=x, z=z),
    )


# ---------------------------------------------------------------------------
# prove  (prover.rs:27-237)
# ---------------------------------------------------------------------------

def _msm(scalars, bases):                 # prover.rs:380-384
    assert len(scalars) <= len(bases)
    if len(scalars) <= 8:
        return msm_naive(scalars, bases)
    return msm_pippenger(scalars, bases)


def prove_polys(pk: ProvingKey, x, w, literal=False):
    """Phase-1 polynomial work: returns u, w, witness-u coefficient vectors and h (prover.rs:73-108)."""
    sap = pk.sap_matrices
    m0 = len(x)
    y = compute_y_vec(sap, x, w, literal=literal)
    rows, cols = sap.size()
    domain = Domain(rows)
    n = domain.size
    if literal:
        u_ev, w_ev, wu_ev, ww_ev = sap_evals_literal(sap, [x, x, w, y], n, m0)
    else:
        u_ev, w_ev, wu_ev, ww_ev = sap_evals_sparse(sap, x, w, y, n)
    u_co = domain.ifft(u_ev)
    w_co = domain.ifft(w_ev)
    sq = Domain(2 * n)                                        # prover.rs:315-328
    ue = sq.fft(u_co)
    u2_co = sq.ifft([v * v % P for v in ue])
    h_num = [(u2_co[k] - (w_co[k] if k < n else 0)) % P for k in range(2 * n)]
    h_num = strip(h_num)
    # divide_by_vanishing_poly (ark-poly)
    if len(h_num) < n:
        h, rem = [], h_num
    else:
        h = list(h_num[n:])
        rem = strip([(h_num[k] + (h[k] if k < len(h) else 0)) % P for k in range(n)])
        h = strip(h)
    assert h and len(h) - 1 <= n - 2, "h is zero or too large (prover.rs:107)"
    assert not rem, "witness does not satisfy the SAP (prover.rs:108)"
    wu_co = domain.ifft(wu_ev)
    ww_co = domain.ifft(ww_ev)
    return dict(n=n, y=y, u_evals=u_ev, w_evals=w_ev, wu_evals=wu_ev, u=u_co, w=w_co, u2=u2_co,
                h=h, h_num=h_num, wu=wu_co, ww=ww_co)


def d_numerator(n, sigma, u, ra, wu, ww, h_num, x2, a_at_x1, c_at_x1):
    """Dense coefficients of A*Y^-g + x2*C*Y^-g - (a(x1)+x2*c(x1))*Y^-g  (prover.rs:142-209)."""
    size = 2 * (n - 1) + sigma * (MINUS_ALPHA + MINUS_GAMMA) + 1
    num = [0] * size

    def add(shift, coeffs, k=1):
        for i, cf in enumerate(coeffs):
            if cf:
                num[shift + i] = (num[shift + i] + k * cf) % P

    s_g = sigma * MINUS_GAMMA
    s_ga = sigma * (MINUS_GAMMA - MINUS_ALPHA)
    s_a = sigma * MINUS_ALPHA
    s_ag = sigma * (MINUS_ALPHA + MINUS_GAMMA)
    # A(X)*Y^-gamma = u*X^{5s} + r_a*X^{2s}                                   prover.rs:145-152
    add(s_g, u)
    add(s_ga, ra)
    # C(X)*Y^-gamma                                                           prover.rs:154-185
    two_ra_u = [0] * (len(u) + 1)
    for i, cf in enumerate(u):
        two_ra_u[i] = (two_ra_u[i] + 2 * ra[0] * cf) % P
        two_ra_u[i + 1] = (two_ra_u[i + 1] + 2 * ra[1] * cf) % P
    ra_sq = [ra[0] * ra[0] % P, 2 * ra[0] * ra[1] % P, ra[1] * ra[1] % P]
    add(s_a, wu, x2)
    add(s_ag, ww, x2)
    add(s_ag, h_num, x2)
    add(s_g, two_ra_u, x2)
    add(s_ga, ra_sq, x2)
    add(0, ra, x2)
    # evaluations                                                             prover.rs:191-209
    num[s_g] = (num[s_g] - a_at_x1 - x2 * c_at_x1) % P
    return num


def divide_by_linear(num, x1):
    """divide_with_q_and_r by (X - x1): q_{k-1} = p_k + x1*q_k  (prover.rs:211-220)."""
    q = [0] * (len(num) - 1)
    carry = 0
    for k in range(len(num) - 1, 0, -1):
        carry = (num[k] + x1 * carry) % P
        q[k - 1] = carry
    rem = (num[0] + x1 * carry) % P
    return q, rem


def create_proof_with_assignment(pk: ProvingKey, x, w, rng, literal=False, trace=None,
                                 transcript_cls=MerlinFieldTranscript) -> Proof:
    polys = prove_polys(pk, x, w, literal=literal)
    n, u, h = polys["n"], polys["u"], polys["h"]
    sigma = pk.vk.sigma
    ra = [fr_rand(rng), fr_rand(rng)]                                        # prover.rs:110
    u_s = strip(u)
    # compute_a_g1  prover.rs:330-338
    a_g1 = g1_add(_msm(u_s, pk.x_powers_g1), _msm(ra, pk.x_powers_y_alpha_g1))
    # compute_r_g1  prover.rs:340-357
    two_ra_u = [0] * (len(u_s) + 1)
    for i, cf in enumerate(u_s):
        two_ra_u[i] = (two_ra_u[i] + 2 * ra[0] * cf) % P
        two_ra_u[i + 1] = (two_ra_u[i + 1] + 2 * ra[1] * cf) % P
    ra_sq = [ra[0] * ra[0] % P, 2 * ra[0] * ra[1] % P, ra[1] * ra[1] % P]
    r_g1 = g1_add(g1_add(_msm(strip(two_ra_u), pk.x_powers_g1), _msm(ra_sq, pk.x_powers_y_alpha_g1)),
                  _msm(ra, pk.x_powers_y_gamma_g1))
    h_g1 = _msm(h, pk.x_powers_zh_by_y_alpha_g1)                              # prover.rs:118
    z_tail = list(x) + list(w) + polys["y"]                                   # z[1..].concat()  prover.rs:120-121
    lcs_g1 = _msm(z_tail, pk.uj_wj_lcs_by_y_alpha_g1)
    c_g1 = g1_add(g1_add(lcs_g1, h_g1), r_g1)                                 # prover.rs:123

    t = transcript_cls(B_POLYMATH)
    x1 = compute_x1(t, list(x), [a_g1, c_g1])                                 # prover.rs:125-126
    y1 = compute_y1(x1, sigma)
    y1_alpha = neg_power(y1, MINUS_ALPHA)
    a_at_x1 = (poly_eval(u, x1) + poly_eval(ra, x1) * y1_alpha) % P           # prover.rs:132
    y1_gamma = neg_power(y1, MINUS_GAMMA)
    pi_at_x1 = compute_pi_at_x1(pk.vk, list(x), x1, y1_gamma)
    c_at_x1 = compute_c_at_x1(y1_gamma, y1_alpha, a_at_x1, pi_at_x1)
    x2 = compute_x2(t, x1, [a_at_x1, c_at_x1])                                # prover.rs:189

    num = d_numerator(n, sigma, u, ra, polys["wu"], polys["ww"], polys["h_num"], x2, a_at_x1, c_at_x1)
    d_coeffs, rem = divide_by_linear(num, x1)
    assert rem == 0, "opening remainder non-zero (prover.rs:221)"
    d_s = strip(d_coeffs)
    assert len(d_s) - 1 <= 2 * (n - 1) + sigma * (MINUS_ALPHA + MINUS_GAMMA)  # prover.rs:222-225
    d_g1 = _msm(d_s, pk.x_powers_y_gamma_z_g1)                                # prover.rs:229
    if trace is not None:
        trace.update(polys)
        trace.update(ra=ra, a_g1=a_g1, c_g1=c_g1, x1=x1, x2=x2, y1_alpha=y1_alpha, c_at_x1=c_at_x1,
                     pi_at_x1=pi_at_x1, numerator=num, d_coeffs=d_coeffs, d_g1=d_g1, a_at_x1=a_at_x1)
    return Proof(a_g1=a_g1, c_g1=c_g1, a_at_x1=a_at_x1, d_g1=d_g1)


def create_proof(circuit, pk, rng, literal=False, trace=None):               # prover.rs:27-64
    cs = synthesize(circuit, setup_mode=False)
    proof = create_proof_with_assignment(pk, cs.instance_assignment, cs.witness_assignment, rng,
                                         literal=literal, trace=trace)
    return proof


# ---------------------------------------------------------------------------
# verify  (verifier.rs:19-62)
# ---------------------------------------------------------------------------

def verify_proof(vk: VerifyingKey, proof: Proof, public_inputs, transcript_cls=MerlinFieldTranscript) -> bool:
    t = transcript_cls(B_POLYMATH)
    pub = [1] + [v % P for v in public_inputs]
    x1 = compute_x1(t, pub, [proof.a_g1, proof.c_g1])
    y1 = compute_y1(x1, vk.sigma)
    y1_gamma = neg_power(y1, MINUS_GAMMA)
    pi_at_x1 = compute_pi_at_x1(vk, pub, x1, y1_gamma)
    y1_alpha = neg_power(y1, MINUS_ALPHA)
    c_at_x1 = compute_c_at_x1(y1_gamma, y1_alpha, proof.a_at_x1, pi_at_x1)
    x2 = compute_x2(t, x1, [proof.a_at_x1, c_at_x1])
    lhs_g1 = g1_add(g1_add(proof.a_g1, g1_mul(proof.c_g1, x2)),
                    g1_mul(vk.one_g1, (-(proof.a_at_x1 + x2 * c_at_x1)) % P))
    x_minus_x1_g2 = g2_add(vk.x_g2, g2_mul(vk.one_g2, (-x1) % P))
    return pairing_product_is_one([(lhs_g1, vk.z_g2), (g1_neg(proof.d_g1), x_minus_x1_g2)])
