// Host-side G1 finishing arithmetic (product code, independent of oracle/).
//
// Each device MSM ends with one XYZZ sum per window (<= 32 points).  Their Horner combination
// sum_w 2^(c*w) * W_w is an inherently serial chain of ~255 doublings; a single GPU thread needs
// milliseconds for it, a host core ~0.1 ms, and the phase has to synchronise with the host anyway to
// hand the commitment to the Fiat-Shamir transcript.  The same routines add the per-GPU partial sums
// of a sharded MSM (group addition is not an NCCL reduction) and produce the canonical affine point.
#pragma once
#include "fp_host.hpp"

namespace pm { namespace host {

struct XyzzH {
    FqH x, y, zz, zzz;
    bool is_inf() const { return zz.is_zero(); }
    static XyzzH inf() { return {FqH::zero(), FqH::zero(), FqH::zero(), FqH::zero()}; }
    static XyzzH from_wire(const uint8_t* b) { return {FqH::from_wire(b), FqH::from_wire(b + 48), FqH::from_wire(b + 96), FqH::from_wire(b + 144)}; }
    void to_wire(uint8_t* b) const { x.to_wire(b); y.to_wire(b + 48); zz.to_wire(b + 96); zzz.to_wire(b + 144); }
};

inline void xyzz_dbl(XyzzH& a) {   // EFD dbl-2008-s-1, a = 0
    if (a.is_inf()) return;
    FqH u = a.y + a.y, v = u.sqr(), w = u * v, s = a.x * v;
    FqH xx = a.x.sqr(), m = xx + xx + xx;
    FqH x3 = m.sqr() - s - s;
    a.y = m * (s - x3) - w * a.y;
    a.x = x3;
    a.zz = v * a.zz;
    a.zzz = w * a.zzz;
}

inline void xyzz_add(XyzzH& a, const XyzzH& b) {   // EFD add-2008-s
    if (b.is_inf()) return;
    if (a.is_inf()) { a = b; return; }
    FqH u1 = a.x * b.zz, u2 = b.x * a.zz, s1 = a.y * b.zzz, s2 = b.y * a.zzz;
    FqH p = u2 - u1, r = s2 - s1;
    if (p.is_zero()) {
        if (r.is_zero()) xyzz_dbl(a); else a = XyzzH::inf();
        return;
    }
    FqH pp = p.sqr(), ppp = p * pp, q = u1 * pp;
    FqH x3 = r.sqr() - ppp - q - q;
    a.y = r * (q - x3) - s1 * ppp;
    a.x = x3;
    a.zz = a.zz * b.zz * pp;
    a.zzz = a.zzz * b.zzz * ppp;
}

// sum_w 2^(c*w) * winsums[w]; winsums are 192-byte XYZZ wire records
inline XyzzH combine_windows(const uint8_t* winsums, int nwin, int c) {
    XyzzH acc = XyzzH::inf();
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) xyzz_dbl(acc);
        xyzz_add(acc, XyzzH::from_wire(winsums + (size_t)w * 192));
    }
    return acc;
}

// sum_g 2^(c*g) * sum_s 2^(shift[s]) * sums[g*nsum + s]   (MsmEngine::Shape)
inline XyzzH combine_shifted(const uint8_t* sums, int nwin, int c, int nsum, const uint8_t* shift) {
    int max_shift = 0;
    for (int s = 0; s < nsum; s++) if (shift[s] > max_shift) max_shift = shift[s];
    XyzzH acc = XyzzH::inf();
    for (int g = nwin - 1; g >= 0; g--) {
        for (int k = 0; k < c; k++) xyzz_dbl(acc);
        XyzzH grp = XyzzH::inf();
        for (int sh = max_shift; sh >= 0; sh--) {
            xyzz_dbl(grp);
            for (int s = 0; s < nsum; s++)
                if (shift[s] == sh) xyzz_add(grp, XyzzH::from_wire(sums + ((size_t)g * nsum + s) * 192));
        }
        xyzz_add(acc, grp);
    }
    return acc;
}

// canonical affine image as 96 wire bytes ((0,0) = infinity)
inline void xyzz_to_affine_wire(const XyzzH& a, uint8_t out[96]) {
    if (a.is_inf()) { memset(out, 0, 96); return; }
    FqH izzz = a.zzz.inv();
    FqH t = a.zz * izzz;
    FqH izz = t.sqr();
    (a.x * izz).to_wire(out);
    (a.y * izzz).to_wire(out + 48);
}

// sum of `count` XYZZ wire records (sharded MSM partials)
inline XyzzH sum_partials(const uint8_t* parts, int count, size_t stride = 192) {
    XyzzH acc = XyzzH::inf();
    for (int i = 0; i < count; i++) xyzz_add(acc, XyzzH::from_wire(parts + (size_t)i * stride));
    return acc;
}

}}  // namespace pm::host
