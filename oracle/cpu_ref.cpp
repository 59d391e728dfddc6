// CPU restatement (C++/OpenMP) of the heavy steps of the Polymath proving path: NTT, MSM, the sparse SAP
// evaluation, the opening-numerator assembly and the division by (X - x1) — everything
// `create_proof_with_assignment` (/root/reference/src/prover.rs:66-237) spends time in.  oracle/fast.py drives
// these into a complete CPU prove (transcript and serialisation stay in oracle/*.py).
// TEST INFRASTRUCTURE / CPU BASELINE ONLY — never linked into libpolymath_b200.so.
//
// What it restates (un-vendored arkworks 0.4.x dependencies of /root/reference):
//   orc_ntt  — ark-poly `Radix2EvaluationDomain::{fft, ifft}` (reference call sites
//              src/prover.rs:241,319,325): in-order radix-2, natural order in/out, n^-1 on inverse;
//   orc_msm  — ark-ec `VariableBaseMSM::msm_unchecked` (src/prover.rs:380-384): window rule
//              c = 3 if n < 32 else ceil(log2 n)*69/100 + 2, signed digits, one bucket set per
//              window, windows processed in parallel (rayon in arkworks, OpenMP here), running-sum
//              bucket reduction, Horner combine.  Output is the canonical affine point.
// Data forms are those of the device ABI (Montgomery little-endian limbs), so the same buffers feed
// both sides.  PARITY UNPINNED against real arkworks; pinned against oracle/*.py in tests/test_oracle_cpp.py.
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

typedef unsigned __int128 u128;

template <int N>
struct Mod {
    uint64_t p[N], r2[N], one[N], ninv;
};

static const Mod<4> FR = {
    {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
    {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull},
    {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
    0xfffffffeffffffffull};
static const Mod<6> FQ = {
    {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
    {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull},
    {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
    0x89f3fffcfffcfffdull};

template <int N, const Mod<N>& M>
struct El {
    uint64_t l[N];
    static El zero() { El e; memset(e.l, 0, sizeof e.l); return e; }
    static El one() { El e; memcpy(e.l, M.one, sizeof e.l); return e; }
    bool is_zero() const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= l[i]; return !o; }
    bool eq(const El& b) const { return !memcmp(l, b.l, sizeof l); }
    static bool ge_p(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) if (a[i] != M.p[i]) return a[i] > M.p[i];
        return true;
    }
    static void sub_p(uint64_t* a) {
        uint64_t br = 0;
        for (int i = 0; i < N; i++) { u128 t = (u128)a[i] - M.p[i] - br; a[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
    }
    El add(const El& b) const {
        El r; uint64_t c = 0;
        for (int i = 0; i < N; i++) { u128 t = (u128)l[i] + b.l[i] + c; r.l[i] = (uint64_t)t; c = (uint64_t)(t >> 64); }
        if (c || ge_p(r.l)) sub_p(r.l);
        return r;
    }
    El sub(const El& b) const {
        El r; uint64_t br = 0;
        for (int i = 0; i < N; i++) { u128 t = (u128)l[i] - b.l[i] - br; r.l[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
        if (br) { uint64_t c = 0; for (int i = 0; i < N; i++) { u128 t = (u128)r.l[i] + M.p[i] + c; r.l[i] = (uint64_t)t; c = (uint64_t)(t >> 64); } }
        return r;
    }
    El neg() const { return zero().sub(*this); }
    El dbl() const { return add(*this); }
    // separated product + Montgomery reduction (SOS form).  A word-serial CIOS product ("no-carry" variant, as in ark-ff)
    // was measured in round 2: faster as a lone dependent chain (39 vs 49 ns per Fq product on the GPU boxes' host) but
    // SLOWER inside the 16-thread MSM (one 2^20 prove 16.8-17.8 s against 14.7-15.2 s), so the baseline keeps this form.
    El mul(const El& b) const {
        uint64_t t[2 * N + 1];
        memset(t, 0, sizeof t);
        for (int i = 0; i < N; i++) {
            uint64_t c = 0;
            for (int j = 0; j < N; j++) { u128 s = (u128)l[i] * b.l[j] + t[i + j] + c; t[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
            t[i + N] = c;
        }
        for (int i = 0; i < N; i++) {
            uint64_t m = t[i] * M.ninv, c = 0;
            for (int j = 0; j < N; j++) { u128 s = (u128)m * M.p[j] + t[i + j] + c; t[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
            for (int k = i + N; c && k <= 2 * N; k++) { u128 s = (u128)t[k] + c; t[k] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        }
        El r; memcpy(r.l, t + N, sizeof r.l);
        if (t[2 * N] || ge_p(r.l)) sub_p(r.l);
        return r;
    }
    El sqr() const { return mul(*this); }
    El from_mont() const { El o = zero(); o.l[0] = 1; return mul(o); }
    El to_mont() const { El r; memcpy(r.l, M.r2, sizeof r.l); return mul(r); }
    El pow(const uint64_t* e, int words) const {
        El acc = one();
        for (int w = words - 1; w >= 0; w--)
            for (int b = 63; b >= 0; b--) { acc = acc.sqr(); if ((e[w] >> b) & 1) acc = acc.mul(*this); }
        return acc;
    }
    El inv() const {
        uint64_t e[N]; memcpy(e, M.p, sizeof e);
        uint64_t br = 2;
        for (int i = 0; i < N && br; i++) { u128 t = (u128)e[i] - br; e[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
        return pow(e, N);
    }
};
typedef El<4, FR> Fr;
typedef El<6, FQ> Fq;

// ------------------------------------------------------------------------------------------
// NTT
// ------------------------------------------------------------------------------------------
static Fr fr_root(int log_n, bool inverse) {
    // 7^((r-1)/2^32) in Montgomery form, squared down to order 2^log_n
    Fr w;
    const uint64_t root[4] = {0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull};
    memcpy(w.l, root, sizeof root);
    for (int i = log_n; i < 32; i++) w = w.sqr();
    return inverse ? w.inv() : w;
}

extern "C" int orc_ntt(uint64_t* data, int log_n, int inverse) {
    const size_t n = (size_t)1 << log_n;
    Fr* a = reinterpret_cast<Fr*>(data);
    // bit reversal (each swap is owned by its smaller index)
    if (log_n > 0) {
#pragma omp parallel for schedule(static)
        for (size_t i = 1; i < n; i++) {
            size_t j = 0;
            for (int b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
            if (i < j) { Fr t = a[i]; a[i] = a[j]; a[j] = t; }
        }
    }
    Fr w_n = fr_root(log_n, inverse != 0);
    std::vector<Fr> tw(n / 2 ? n / 2 : 1);
    {
        // w^k for k < n/2: every thread starts its slice with one pow, then a running product
        const size_t half_n = n / 2 ? n / 2 : 1;
        const int nt = omp_get_max_threads();
        const size_t per = (half_n + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            const size_t lo = (size_t)t * per, hi = lo + per < half_n ? lo + per : half_n;
            if (lo >= hi) continue;
            uint64_t e[1] = {lo};
            Fr cur = w_n.pow(e, 1);
            for (size_t k = lo; k < hi; k++) { tw[k] = cur; cur = cur.mul(w_n); }
        }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, step = n / len;
        if (len <= 1024 || n / len >= 8) {
#pragma omp parallel for schedule(static)
            for (size_t start = 0; start < n; start += len)
                for (size_t k = 0; k < half; k++) {
                    Fr u = a[start + k], v = a[start + k + half].mul(tw[k * step]);
                    a[start + k] = u.add(v);
                    a[start + k + half] = u.sub(v);
                }
        } else {
            for (size_t start = 0; start < n; start += len) {
#pragma omp parallel for schedule(static)
                for (size_t k = 0; k < half; k++) {
                    Fr u = a[start + k], v = a[start + k + half].mul(tw[k * step]);
                    a[start + k] = u.add(v);
                    a[start + k + half] = u.sub(v);
                }
            }
        }
    }
    if (inverse) {
        uint64_t nn[4] = {n, 0, 0, 0};
        Fr nf; memcpy(nf.l, nn, sizeof nn);
        Fr ninv = nf.to_mont().inv();
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) a[i] = a[i].mul(ninv);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// G1 (Jacobian) and MSM
// ------------------------------------------------------------------------------------------
struct Aff { Fq x, y; bool inf() const { return x.is_zero() && y.is_zero(); } };
struct Jac { Fq x, y, z; bool inf() const { return z.is_zero(); } };
static Jac jac_inf() { return {Fq::one(), Fq::one(), Fq::zero()}; }

static Jac jac_dbl(const Jac& p) {
    if (p.inf()) return p;
    Fq a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
    Fq d = p.x.add(b).sqr().sub(a).sub(c).dbl();
    Fq e = a.dbl().add(a), f = e.sqr();
    Jac r;
    r.x = f.sub(d.dbl());
    r.y = e.mul(d.sub(r.x)).sub(c.dbl().dbl().dbl());
    r.z = p.y.mul(p.z).dbl();
    return r;
}
static Jac jac_add(const Jac& p, const Jac& q) {
    if (p.inf()) return q;
    if (q.inf()) return p;
    Fq z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
    Fq u1 = p.x.mul(z2z2), u2 = q.x.mul(z1z1);
    Fq s1 = p.y.mul(q.z).mul(z2z2), s2 = q.y.mul(p.z).mul(z1z1);
    if (u1.eq(u2)) return s1.eq(s2) ? jac_dbl(p) : jac_inf();
    Fq h = u2.sub(u1), rr = s2.sub(s1);
    Fq hh = h.sqr(), hhh = h.mul(hh), v = u1.mul(hh);
    Jac r;
    r.x = rr.sqr().sub(hhh).sub(v.dbl());
    r.y = rr.mul(v.sub(r.x)).sub(s1.mul(hhh));
    r.z = p.z.mul(q.z).mul(h);
    return r;
}
static Jac jac_madd(const Jac& p, const Aff& q, bool neg) {
    if (q.inf()) return p;
    Fq qy = neg ? q.y.neg() : q.y;
    if (p.inf()) return {q.x, qy, Fq::one()};
    Fq z1z1 = p.z.sqr();
    Fq u2 = q.x.mul(z1z1), s2 = qy.mul(p.z).mul(z1z1);
    if (u2.eq(p.x)) return s2.eq(p.y) ? jac_dbl(p) : jac_inf();
    Fq h = u2.sub(p.x), rr = s2.sub(p.y);
    Fq hh = h.sqr(), hhh = h.mul(hh), v = p.x.mul(hh);
    Jac r;
    r.x = rr.sqr().sub(hhh).sub(v.dbl());
    r.y = rr.mul(v.sub(r.x)).sub(p.y.mul(hhh));
    r.z = p.z.mul(h);
    return r;
}

static int ark_window(size_t n) {
    if (n < 32) return 3;
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    return lg * 69 / 100 + 2;
}

extern "C" int orc_msm_window_bits(size_t n) { return ark_window(n); }

extern "C" int orc_msm(const uint64_t* bases_raw, const uint64_t* scalars_raw, size_t n, uint64_t* out_affine) {
    const Aff* bases = reinterpret_cast<const Aff*>(bases_raw);
    const Fr* sc = reinterpret_cast<const Fr*>(scalars_raw);
    const int c = ark_window(n);
    const int nwin = (255 + c - 1) / c + 1;   // one extra window absorbs the last signed-digit carry
    const size_t nb = (size_t)1 << (c - 1);
    // signed digits
    std::vector<int32_t> digits((size_t)nwin * n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Fr s = sc[i].from_mont();
        int carry = 0;
        for (int w = 0; w < nwin; w++) {
            int pos = w * c, limb = pos >> 6, off = pos & 63;
            uint64_t v = limb < 4 ? s.l[limb] >> off : 0;
            if (off + c > 64 && limb + 1 < 4) v |= s.l[limb + 1] << (64 - off);
            int d = (int)(v & (((uint64_t)1 << c) - 1)) + carry;
            if (d > (int)nb) { d -= 1 << c; carry = 1; } else carry = 0;
            digits[(size_t)w * n + i] = d;
        }
    }
    // arkworks 0.4 runs one rayon task per window (15-17 tasks at these sizes); on a 16-32 thread host that leaves a
    // third of the cores idle in the last round, so the points of a window are also cut into chunks: every
    // (window, chunk) task fills and reduces its own bucket set, the chunk sums of a window are added afterwards.
    int nchunk = 1;
    {
        const int nt = omp_get_max_threads();
        if (n >= ((size_t)1 << 14)) nchunk = (2 * nt + nwin - 1) / nwin;
        if (nchunk < 1) nchunk = 1;
        // keep the running-sum reduction (2 * nb additions per task) below ~15 % of the task's mixed additions
        while (nchunk > 1 && (n / nchunk) < 12 * nb) nchunk--;
    }
    std::vector<Jac> part((size_t)nwin * nchunk);
#pragma omp parallel for schedule(dynamic, 1)
    for (int task = 0; task < nwin * nchunk; task++) {
        const int w = task / nchunk, ch = task % nchunk;
        const size_t lo = n * (size_t)ch / nchunk, hi = n * (size_t)(ch + 1) / nchunk;
        std::vector<Jac> buckets(nb, jac_inf());
        const int32_t* dw = &digits[(size_t)w * n];
        for (size_t i = lo; i < hi; i++) {
            int d = dw[i];
            if (d > 0) buckets[d - 1] = jac_madd(buckets[d - 1], bases[i], false);
            else if (d < 0) buckets[-d - 1] = jac_madd(buckets[-d - 1], bases[i], true);
        }
        Jac running = jac_inf(), acc = jac_inf();
        for (size_t b = nb; b-- > 0;) { running = jac_add(running, buckets[b]); acc = jac_add(acc, running); }
        part[task] = acc;
    }
    std::vector<Jac> wsum(nwin);
    for (int w = 0; w < nwin; w++) {
        Jac acc = jac_inf();
        for (int ch = 0; ch < nchunk; ch++) acc = jac_add(acc, part[(size_t)w * nchunk + ch]);
        wsum[w] = acc;
    }
    Jac total = jac_inf();
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) total = jac_dbl(total);
        total = jac_add(total, wsum[w]);
    }
    Aff* out = reinterpret_cast<Aff*>(out_affine);
    if (total.inf()) { out->x = Fq::zero(); out->y = Fq::zero(); return 0; }
    Fq zi = total.z.inv(), zi2 = zi.sqr();
    out->x = total.x.mul(zi2);
    out->y = total.y.mul(zi2).mul(zi);
    return 0;
}

// Synthetic bases for the CPU baseline: out[i] = (i+1)*G, canonical affine (chain of mixed additions,
// one batched inversion).  Any valid points do for a throughput sample.
extern "C" int orc_make_bases(size_t n, uint64_t* out_affine) {
    static const uint64_t gx[6] = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull, 0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull};
    static const uint64_t gy[6] = {0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull, 0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};
    Aff g;
    memcpy(g.x.l, gx, sizeof gx);
    memcpy(g.y.l, gy, sizeof gy);
    std::vector<Jac> pts(n);
    Jac cur = jac_inf();
    for (size_t i = 0; i < n; i++) { cur = jac_madd(cur, g, false); pts[i] = cur; }
    std::vector<Fq> prefix(n);
    Fq acc = Fq::one();
    for (size_t i = 0; i < n; i++) { prefix[i] = acc; acc = acc.mul(pts[i].z); }
    Fq inv = acc.inv();
    Aff* out = reinterpret_cast<Aff*>(out_affine);
    for (size_t i = n; i-- > 0;) {
        Fq zi = inv.mul(prefix[i]);
        inv = inv.mul(pts[i].z);
        Fq zi2 = zi.sqr();
        out[i].x = pts[i].x.mul(zi2);
        out[i].y = pts[i].y.mul(zi2).mul(zi);
    }
    return 0;
}

extern "C" int orc_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    const Fr* x = reinterpret_cast<const Fr*>(a);
    const Fr* y = reinterpret_cast<const Fr*>(b);
    Fr* o = reinterpret_cast<Fr*>(out);
    for (size_t i = 0; i < n; i++) o[i] = x[i].mul(y[i]);
    return 0;
}
extern "C" int orc_fq_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    const Fq* x = reinterpret_cast<const Fq*>(a);
    const Fq* y = reinterpret_cast<const Fq*>(b);
    Fq* o = reinterpret_cast<Fq*>(out);
    for (size_t i = 0; i < n; i++) o[i] = x[i].mul(y[i]);
    return 0;
}

// ------------------------------------------------------------------------------------------
// the remaining steps of create_proof_with_assignment (src/prover.rs:66-237), sparse closed form of SURVEY.md 8(a3)
// ------------------------------------------------------------------------------------------
// out[r] = sum_k val[k] * vec[col[k]] over row r of a CSR matrix (first-match duplicates are the caller's business:
// ark-relations compacts rows, the product dedups at upload)
extern "C" int orc_spmv(const uint64_t* row_ptr, const uint32_t* col, const uint64_t* val_raw, size_t nr, const uint64_t* vec_raw,
                        uint64_t* out_raw) {
    const Fr* val = reinterpret_cast<const Fr*>(val_raw);
    const Fr* vec = reinterpret_cast<const Fr*>(vec_raw);
    Fr* out = reinterpret_cast<Fr*>(out_raw);
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nr; r++) {
        Fr acc = Fr::zero();
        for (uint64_t k = row_ptr[r]; k < row_ptr[r + 1]; k++) acc = acc.add(val[k].mul(vec[col[k]]));
        out[r] = acc;
    }
    return 0;
}

// compute_y_vec (prover.rs:279-302) and U.z, W.z, witness-U.z (prover.rs:87-96,156-166) from A.z', B.z', C.z':
//   y = [0] | (1 - x_j)^2 | (az - bz)^2;  rows i < m0: (1 + x_i, 4 x_i + y_i);  m0 + i: (1 - x_i, y_i);
//   2 m0 + r: (az + bz, 4 cz + y_{m0+r});  2 m0 + nr + r: (az - bz, y_{m0+r});  rest 0;  wu = u with rows < 2 m0 zeroed
extern "C" int orc_sap_evals(size_t m0, size_t nr, size_t n, const uint64_t* x_raw, const uint64_t* az_raw, const uint64_t* bz_raw,
                             const uint64_t* cz_raw, uint64_t* y_raw, uint64_t* u_raw, uint64_t* w_raw, uint64_t* wu_raw) {
    const Fr *x = reinterpret_cast<const Fr*>(x_raw), *az = reinterpret_cast<const Fr*>(az_raw),
             *bz = reinterpret_cast<const Fr*>(bz_raw), *cz = reinterpret_cast<const Fr*>(cz_raw);
    Fr *y = reinterpret_cast<Fr*>(y_raw), *u = reinterpret_cast<Fr*>(u_raw), *w = reinterpret_cast<Fr*>(w_raw),
       *wu = reinterpret_cast<Fr*>(wu_raw);
    if (2 * (m0 + nr) > n) return 1;
    const Fr one = Fr::one();
    y[0] = Fr::zero();
    for (size_t j = 1; j < m0; j++) y[j] = one.sub(x[j]).sqr();
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nr; r++) y[m0 + r] = az[r].sub(bz[r]).sqr();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { u[i] = Fr::zero(); w[i] = Fr::zero(); wu[i] = Fr::zero(); }
    for (size_t i = 0; i < m0; i++) {
        u[i] = one.add(x[i]);
        w[i] = x[i].dbl().dbl().add(y[i]);
        u[m0 + i] = one.sub(x[i]);
        w[m0 + i] = y[i];
    }
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nr; r++) {
        u[2 * m0 + r] = az[r].add(bz[r]);
        w[2 * m0 + r] = cz[r].dbl().dbl().add(y[m0 + r]);
        u[2 * m0 + nr + r] = az[r].sub(bz[r]);
        w[2 * m0 + nr + r] = y[m0 + r];
        wu[2 * m0 + r] = u[2 * m0 + r];
        wu[2 * m0 + nr + r] = u[2 * m0 + nr + r];
    }
    return 0;
}

// pointwise square (square_polynomial, prover.rs:321-323)
extern "C" int orc_fr_square(uint64_t* a_raw, size_t n) {
    Fr* a = reinterpret_cast<Fr*>(a_raw);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) a[i] = a[i].sqr();
    return 0;
}

// h_num = u2 - w (2n coefficients), h = h_num[n..2n), remainder h_num[k] + h[k] (k < n) must vanish
// (divide_by_vanishing_poly + the asserts of prover.rs:105-108).  Returns 0 ok, 3 remainder non-zero, 4 h zero / degree.
extern "C" int orc_quotient(const uint64_t* u2_raw, const uint64_t* w_raw, size_t n, uint64_t* hnum_raw) {
    const Fr *u2 = reinterpret_cast<const Fr*>(u2_raw), *w = reinterpret_cast<const Fr*>(w_raw);
    Fr* hn = reinterpret_cast<Fr*>(hnum_raw);
    int bad = 0, nonzero = 0;
#pragma omp parallel for schedule(static) reduction(| : bad, nonzero)
    for (size_t k = 0; k < n; k++) {
        hn[k] = u2[k].sub(w[k]);
        hn[n + k] = u2[n + k];
        if (!hn[k].add(hn[n + k]).is_zero()) bad |= 1;
        if (!hn[n + k].is_zero()) nonzero |= 1;
    }
    if (bad) return 3;
    if (!nonzero || !hn[2 * n - 1].is_zero()) return 4;     // deg h <= n - 2
    return 0;
}

// out[i] = 2 (ra0 u[i] + ra1 u[i-1]), i <= n   (compute_r_g1, prover.rs:340-347)
extern "C" int orc_two_ra_u(const uint64_t* u_raw, size_t n, const uint64_t* ra_raw, uint64_t* out_raw) {
    const Fr *u = reinterpret_cast<const Fr*>(u_raw), *ra = reinterpret_cast<const Fr*>(ra_raw);
    Fr* out = reinterpret_cast<Fr*>(out_raw);
    const Fr r0 = ra[0].dbl(), r1 = ra[1].dbl();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i <= n; i++) {
        Fr v = Fr::zero();
        if (i < n) v = v.add(r0.mul(u[i]));
        if (i > 0) v = v.add(r1.mul(u[i - 1]));
        out[i] = v;
    }
    return 0;
}

// Horner evaluation (u_poly.evaluate(&x1), prover.rs:132); sequential like ark-poly's small case, chunked here
extern "C" int orc_horner(const uint64_t* c_raw, size_t n, const uint64_t* x_raw, uint64_t* out_raw) {
    const Fr* c = reinterpret_cast<const Fr*>(c_raw);
    const Fr x = *reinterpret_cast<const Fr*>(x_raw);
    const int nt = omp_get_max_threads();
    const size_t per = (n + nt - 1) / (nt ? nt : 1);
    std::vector<Fr> part(nt, Fr::zero());
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
        Fr acc = Fr::zero();
        for (size_t k = hi; k-- > lo;) acc = acc.mul(x).add(c[k]);
        part[t] = acc;
    }
    uint64_t e[1] = {per};
    const Fr xp = x.pow(e, 1);
    Fr acc = Fr::zero();
    for (int t = nt; t-- > 0;) acc = acc.mul(xp).add(part[t]);
    *reinterpret_cast<Fr*>(out_raw) = acc;
    return 0;
}

// Dense coefficients of A*Y^-gamma + x2*C*Y^-gamma - (a(x1) + x2*c(x1))*X^(5 sigma)  (prover.rs:142-209; the shifts
// 2s, 3s, 5s, 8s with s = sigma), size 2(n-1) + 8 sigma + 1.  `consts` = [ra0, ra1, x2, a_at_x1, c_at_x1].
extern "C" int orc_d_numerator(size_t n, size_t sigma, const uint64_t* u_raw, const uint64_t* wu_raw, const uint64_t* ww_raw,
                               const uint64_t* hnum_raw, const uint64_t* consts_raw, uint64_t* out_raw) {
    const Fr *u = reinterpret_cast<const Fr*>(u_raw), *wu = reinterpret_cast<const Fr*>(wu_raw),
             *ww = reinterpret_cast<const Fr*>(ww_raw), *hn = reinterpret_cast<const Fr*>(hnum_raw),
             *k = reinterpret_cast<const Fr*>(consts_raw);
    Fr* num = reinterpret_cast<Fr*>(out_raw);
    const size_t size = 2 * (n - 1) + 8 * sigma + 1;
    const Fr ra0 = k[0], ra1 = k[1], x2 = k[2], a_at = k[3], c_at = k[4];
    const size_t s_g = 5 * sigma, s_ga = 2 * sigma, s_a = 3 * sigma, s_ag = 8 * sigma;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; i++) num[i] = Fr::zero();
    // the supports [0,2) [2s,2s+3) [3s,3s+n) [5s,5s+n+1) [8s,8s+2n-1) are disjoint (sigma = n + 3): plain parallel loops
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        num[s_g + i] = num[s_g + i].add(u[i]);                    // A: u * X^{5s}
        num[s_a + i] = num[s_a + i].add(x2.mul(wu[i]));           // C: x2 * wu * X^{3s}
    }
    const Fr r0 = ra0.dbl(), r1 = ra1.dbl();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i <= n; i++) {                             // x2 * two_ra_u * X^{5s}
        Fr v = Fr::zero();
        if (i < n) v = v.add(r0.mul(u[i]));
        if (i > 0) v = v.add(r1.mul(u[i - 1]));
        num[s_g + i] = num[s_g + i].add(x2.mul(v));
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < 2 * n; i++) {                          // x2 * (ww + h_num) * X^{8s}
        Fr v = hn[i];
        if (i < n) v = v.add(ww[i]);
        if (s_ag + i < size) num[s_ag + i] = num[s_ag + i].add(x2.mul(v));
    }
    // r_a terms
    num[s_ga] = num[s_ga].add(ra0);
    num[s_ga + 1] = num[s_ga + 1].add(ra1);
    const Fr sq[3] = {ra0.sqr(), ra0.mul(ra1).dbl(), ra1.sqr()};
    for (int i = 0; i < 3; i++) num[s_ga + i] = num[s_ga + i].add(x2.mul(sq[i]));
    num[0] = num[0].add(x2.mul(ra0));
    num[1] = num[1].add(x2.mul(ra1));
    num[s_g] = num[s_g].sub(a_at).sub(x2.mul(c_at));
    return 0;
}

// divide_with_q_and_r by (X - x1): q_{k-1} = p_k + x1 q_k (prover.rs:211-220).  The reference's division is one
// sequential recurrence; here every thread runs the recurrence of its coefficient range from a zero carry and the
// carries are then propagated with powers of x1 (q is linear in the incoming carry).  rem_out = p_0 + x1 q_0.
extern "C" int orc_divide_linear(const uint64_t* num_raw, size_t len, const uint64_t* x1_raw, uint64_t* q_raw, uint64_t* rem_out) {
    const Fr* num = reinterpret_cast<const Fr*>(num_raw);
    const Fr x1 = *reinterpret_cast<const Fr*>(x1_raw);
    Fr* q = reinterpret_cast<Fr*>(q_raw);
    if (len < 2) return 1;
    const size_t m = len - 1;                                    // q has m coefficients: q[k-1] for k = m..1
    const int nt = omp_get_max_threads();
    const size_t per = (m + nt - 1) / nt;
    std::vector<Fr> tail(nt, Fr::zero());
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        const size_t lo = (size_t)t * per, hi = lo + per < m ? lo + per : m;     // q indices [lo, hi)
        Fr carry = Fr::zero();
        for (size_t j = hi; j-- > lo;) { carry = num[j + 1].add(x1.mul(carry)); q[j] = carry; }
        tail[t] = carry;                                          // q[lo] with zero incoming carry
    }
    // incoming carry of range t = true q[hi_t] = q[lo_{t+1}]; adding carry c at hi changes q[j] by c * x1^(hi - j)
    std::vector<Fr> incoming(nt, Fr::zero());
    uint64_t e[1] = {per};
    const Fr xper = x1.pow(e, 1);
    for (int t = nt - 2; t >= 0; t--) {
        const size_t lo_next = (size_t)(t + 1) * per;
        if (lo_next >= m) continue;
        const size_t hi_next = lo_next + per < m ? lo_next + per : m;
        uint64_t e2[1] = {hi_next - lo_next};
        const Fr xlen = (hi_next - lo_next == per) ? xper : x1.pow(e2, 1);
        incoming[t] = tail[t + 1].add(incoming[t + 1].mul(xlen));
    }
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        const size_t lo = (size_t)t * per, hi = lo + per < m ? lo + per : m;
        if (lo >= hi || incoming[t].is_zero()) continue;
        Fr f = incoming[t];
        for (size_t j = hi; j-- > lo;) { f = f.mul(x1); q[j] = q[j].add(f); }
    }
    *reinterpret_cast<Fr*>(rem_out) = num[0].add(x1.mul(q[0]));
    return 0;
}

extern "C" int orc_num_threads(void) { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm of bench.py asks for the host's cores explicitly
extern "C" void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
