"""BLS12-381 pairing-product check for the oracle verifier (test infrastructure).

The reference verifier ends in `E::multi_pairing(..).0.is_one()`
(`/root/reference/src/verifier.rs:50-61`).  Only the *predicate* "product of
pairings == 1" is observable, so this restatement uses the simplest correct
construction: an affine Miller loop on the twist with lines embedded in
Fq12 = Fq[w]/(w^12 - 2w^6 + 2), and a direct exponentiation by (q^12-1)/r.
(The sign of the BLS parameter only inverts every factor and cannot change
whether the product is one, so it is ignored.)
"""
from .fields import (
    Q_MOD, R_MOD, BLS_X_ABS, FQ12_ONE, fq12_mul, fq12_pow, fq2_to_fq12,
    fq2_add, fq2_sub, fq2_mul, fq2_inv, fq2_scalar, fq2_neg,
)

FINAL_EXP = (Q_MOD ** 12 - 1) // R_MOD


def _line(lam, xt, yt, xp, yp):
    """Line through T=(xt,yt) with twist-slope lam, evaluated at P=(xp,yp), times w^3.

    Untwist (x', y') -> (x'/w^2, y'/w^3); slope on E(Fq12) is lam/w.  Then
    l * w^3 = (lam*x' - y') - lam*xp * w^2 + yp * w^3  (w^3 lies in Fq4 and is
    killed by the final exponentiation).
    """
    c0 = fq2_to_fq12(fq2_sub(fq2_mul(lam, xt), yt))
    c2 = fq2_to_fq12(fq2_neg(fq2_scalar(lam, xp)))
    out = list(c0)
    # multiply c2 by w^2: shift by 2 (coefficients live at positions 0 and 6 only)
    out[2] = (out[2] + c2[0]) % Q_MOD
    out[8] = (out[8] + c2[6]) % Q_MOD
    out[3] = (out[3] + yp) % Q_MOD
    return out


def miller_loop(p_aff, q_aff):
    """f_{|x|,Q}(P) for P in G1 (affine ints), Q in G2 (affine Fq2)."""
    if p_aff is None or q_aff is None:
        return list(FQ12_ONE)
    xp, yp = p_aff
    xq, yq = q_aff
    xt, yt = xq, yq
    f = list(FQ12_ONE)
    for bit in bin(BLS_X_ABS)[3:]:
        lam = fq2_mul(fq2_scalar(fq2_mul(xt, xt), 3), fq2_inv(fq2_scalar(yt, 2)))
        f = fq12_mul(fq12_mul(f, f), _line(lam, xt, yt, xp, yp))
        x3 = fq2_sub(fq2_mul(lam, lam), fq2_scalar(xt, 2))
        y3 = fq2_sub(fq2_mul(lam, fq2_sub(xt, x3)), yt)
        xt, yt = x3, y3
        if bit == "1":
            lam = fq2_mul(fq2_sub(yq, yt), fq2_inv(fq2_sub(xq, xt)))
            f = fq12_mul(f, _line(lam, xt, yt, xp, yp))
            x3 = fq2_sub(fq2_sub(fq2_mul(lam, lam), xt), xq)
            y3 = fq2_sub(fq2_mul(lam, fq2_sub(xt, x3)), yt)
            xt, yt = x3, y3
    return f


def pairing(p_aff, q_aff):
    return fq12_pow(miller_loop(p_aff, q_aff), FINAL_EXP)


def pairing_product_is_one(pairs) -> bool:
    f = list(FQ12_ONE)
    for p_aff, q_aff in pairs:
        f = fq12_mul(f, miller_loop(p_aff, q_aff))
    return fq12_pow(f, FINAL_EXP) == FQ12_ONE
