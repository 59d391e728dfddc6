// G1 point (de)compression on the device — SURVEY.md 8(f)-1, the proving-key cold-start path.
//
// Replaces what `ProvingKey::deserialize_compressed[_unchecked]` / `serialize_compressed`
// (/root/reference/src/data_structures.rs:55-73, derive(CanonicalSerialize/Deserialize)) spend per point inside
// ark-bls12-381's `read_g1_compressed` / `serialize_with_mode` (the zcash encoding): 48 bytes, big-endian x,
// top bits of byte 0 = compressed (0x80), infinity (0x40), y lexicographically largest (0x20).  Decoding needs a
// square root in Fq (q = 3 mod 4: y = (x^3 + 4)^((q+1)/4), ~570 field products) per point — 13n + m points per key.
// One thread per point; results are the canonical affine point (Montgomery limbs, (0,0) = infinity), so they are
// independent of how the root is computed.
#include "g1_codec.cuh"

namespace pm {

namespace {

// [r]P == O for an affine point on the curve (r = the scalar-field modulus): `is_in_correct_subgroup_assuming_on_curve`
// decides the same predicate with the endomorphism shortcut; the plain ladder is 255 doublings + ~130 additions.
__device__ bool g1_in_subgroup(const G1Affine& p) {
    G1XYZZ acc = G1XYZZ::inf();
    for (int w = 7; w >= 0; w--) {
        const uint32_t limb = FR_MOD[w];
        for (int bit = 31; bit >= 0; bit--) {
            xyzz_dbl(acc);
            if ((limb >> bit) & 1u) xyzz_madd(acc, p, false);
        }
    }
    return acc.is_inf();
}

// y (Montgomery form) is the lexicographically larger of {y, -y}  <=>  2 * y_canonical >= q
__device__ __forceinline__ bool fq_is_largest(const Fq& y_mont) {
    Fq y = y_mont.from_mont();
    uint32_t t[13];
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        t[i] = (y.v[i] << 1) | carry;
        carry = y.v[i] >> 31;
    }
    t[12] = carry;
    // t - q: no borrow <=> t >= q
    uint32_t d = ptx::sub_cc(t[0], FQ_MOD[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) d = ptx::subc_cc(t[i], FQ_MOD[i]);
    d = ptx::subc_cc(t[12], 0u);
    const uint32_t borrow = ptx::subc(0u, 0u);
    (void)d;
    return borrow == 0;
}

// status: 0 = ok, 1 = compression flag missing, 2 = x >= q, 3 = not on the curve, 4 = not in the prime-order subgroup,
// 5 = infinity flag with the sort flag or a non-zero payload (ark-bls12-381 `read_g1_compressed` and the host codec
// accept exactly 0xc0 00 .. 00 as the point at infinity)
__global__ void __launch_bounds__(128) k_g1_decompress(const uint8_t* __restrict__ in, size_t n, int validate,
                                                       G1Affine* __restrict__ out, unsigned long long* __restrict__ first_bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + 48 * i;
    const uint8_t flags = b[0];
    uint32_t status = 0;
    G1Affine p = G1Affine::inf();
    if (!(flags & 0x80)) {
        status = 1;
    } else if (flags & 0x40) {
        uint32_t rest = flags & 0x3fu;
        for (int j = 1; j < 48; j++) rest |= b[j];
        if (rest) status = 5;
    } else {
        Fq x;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const uint8_t* q = b + 44 - 4 * j;
            x.v[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
        }
        x.v[11] &= 0x1fffffffu;
        // canonical: x < q
        uint32_t d = ptx::sub_cc(x.v[0], FQ_MOD[0]);
#pragma unroll
        for (int j = 1; j < 12; j++) d = ptx::subc_cc(x.v[j], FQ_MOD[j]);
        const uint32_t borrow = ptx::subc(0u, 0u);
        (void)d;
        if (borrow == 0) {
            status = 2;
        } else {
            x = x.to_mont();
            Fq two = Fq::one().dbl(), four = two.dbl();
            Fq rhs = x.sqr() * x + four;                       // y^2 = x^3 + 4
            // e = (q + 1) / 4
            uint32_t e[12];
            e[0] = ptx::add_cc(FQ_MOD[0], 1u);
#pragma unroll
            for (int j = 1; j < 12; j++) e[j] = ptx::addc_cc(FQ_MOD[j], 0u);
#pragma unroll
            for (int j = 0; j < 12; j++) e[j] = (e[j] >> 2) | (j + 1 < 12 ? e[j + 1] << 30 : 0u);
            Fq y = rhs.pow(e, 12);
            if (y.sqr() != rhs) {
                status = 3;
            } else {
                const bool want_largest = (flags & 0x20) != 0;
                if (fq_is_largest(y) != want_largest) y = y.neg();
                p.x = x;
                p.y = y;
                if (validate && !g1_in_subgroup(p)) status = 4;
            }
        }
    }
    if (status) {
        p = G1Affine::inf();
        atomicMin(first_bad, ((unsigned long long)i << 3) | status);   // the lowest failing index wins
    }
    out[i] = p;
}

__global__ void __launch_bounds__(128) k_g1_compress(const G1Affine* __restrict__ in, size_t n, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = in[i];
    uint8_t* b = out + 48 * i;
    if (p.is_inf()) {
        for (int k = 0; k < 48; k++) b[k] = 0;
        b[0] = 0xc0;
        return;
    }
    Fq x = p.x.from_mont();
#pragma unroll
    for (int j = 0; j < 12; j++) {
        uint8_t* q = b + 44 - 4 * j;
        q[0] = (uint8_t)(x.v[j] >> 24); q[1] = (uint8_t)(x.v[j] >> 16); q[2] = (uint8_t)(x.v[j] >> 8); q[3] = (uint8_t)x.v[j];
    }
    b[0] |= 0x80 | (fq_is_largest(p.y) ? 0x20 : 0x00);
}

}  // namespace

void launch_g1_decompress(const uint8_t* in_dev, size_t n, bool validate, G1Affine* out_dev, unsigned long long* first_bad_dev,
                          cudaStream_t stream) {
    if (n == 0) return;
    k_g1_decompress<<<ceil_div(n, 128), 128, 0, stream>>>(in_dev, n, validate ? 1 : 0, out_dev, first_bad_dev);
    PM_LAUNCH_CHECK();
}

void launch_g1_compress(const G1Affine* in_dev, size_t n, uint8_t* out_dev, cudaStream_t stream) {
    if (n == 0) return;
    k_g1_compress<<<ceil_div(n, 128), 128, 0, stream>>>(in_dev, n, out_dev);
    PM_LAUNCH_CHECK();
}

const char* g1_decode_status_name(unsigned status) {
    switch (status) {
        case 1: return "compression flag not set";
        case 2: return "x coordinate not below the field modulus";
        case 3: return "x is not the abscissa of a curve point";
        case 4: return "point is not in the prime-order subgroup";
        case 5: return "infinity flag set on a non-zero encoding";
        default: return "ok";
    }
}

}  // namespace pm
