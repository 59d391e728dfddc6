// C++ host mirror of `Polymath::<Bls12_381, MerlinFieldTranscript<Fr>>::setup / prove`
// (/root/reference/src/lib.rs:52-91) above the phase C-ABI.  The reference's own toolchain
// (cargo/rustc) is absent from this image, so the Rust host side of INTEGRATION.md is
// mirrored here: same flow, same RNG consumption, same transcript bytes, same errors.
//
//   pm_polymath_setup  = generate_proving_key           (src/generator.rs:24-167)
//   pm_polymath_prove  = create_proof_with_assignment   (src/prover.rs:66-237)
//
// Circuit synthesis (src/prover.rs:33-52) stays with the caller: these entry points take the
// R1CS matrices / the instance and witness assignments it produces.
#include <memory>
#include <vector>

#include "../../../include/polymath_b200.h"
#include "../common.cuh"
#include "transcript_host.hpp"
#include "pairing_host.hpp"

using namespace pm::host;

struct pm_rng {
    StdRng rng;
    explicit pm_rng(StdRng r) : rng(r) {}
};

namespace {

const char* B_POLYMATH = "polymath";   // common.rs:8
constexpr uint64_t MINUS_ALPHA = 3;     // common.rs:11
constexpr uint64_t MINUS_GAMMA = 5;     // common.rs:14

FrH neg_power(const FrH& y, uint64_t minus_exp) { return y.inv().pow_u64(minus_exp); }   // common.rs:45-47

// common.rs:77-97
FrH z_tilde_i(const std::vector<FrH>& pub, size_t i) {
    const size_t m0 = pub.size();
    FrH one = FrH::one();
    if (i == 0) return one + one;
    if (i < m0) return one + pub[i];
    if (i == m0) return FrH::zero();
    return one - pub[i - m0];
}

// common.rs:49-71
FrH compute_pi_at_x1(uint64_t n, const FrH& omega, const std::vector<FrH>& pub, const FrH& x1, const FrH& y1_gamma) {
    FrH sum = FrH::zero();
    FrH num = (x1.pow_u64(n) - FrH::one()) * FrH::from_u64(n).inv();
    FrH omega_i = FrH::one();
    for (size_t i = 0; i < 2 * pub.size(); i++) {
        FrH lagrange = num * (x1 - omega_i).inv();
        sum = sum + z_tilde_i(pub, i) * lagrange;
        num = num * omega;
        omega_i = omega_i * omega;
    }
    return sum * y1_gamma;
}

// zcash encodings of the BLS12-381 generators (VerifyingKey.e.one_g1 / one_g2, data_structures.rs:27-29)
const uint8_t G1_GEN_COMPRESSED[48] = {
    0x97, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f, 0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05,
    0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58, 0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};
const uint8_t G2_GEN_COMPRESSED[96] = {
    0x93, 0xe0, 0x2b, 0x60, 0x52, 0x71, 0x9f, 0x60, 0x7d, 0xac, 0xd3, 0xa0, 0x88, 0x27, 0x4f, 0x65, 0x59, 0x6b, 0xd0, 0xd0, 0x99, 0x20, 0xb6, 0x1a,
    0xb5, 0xda, 0x61, 0xbb, 0xdc, 0x7f, 0x50, 0x49, 0x33, 0x4c, 0xf1, 0x12, 0x13, 0x94, 0x5d, 0x57, 0xe5, 0xac, 0x7d, 0x05, 0x5d, 0x04, 0x2b, 0x7e,
    0x02, 0x4a, 0xa2, 0xb2, 0xf0, 0x8f, 0x0a, 0x91, 0x26, 0x08, 0x05, 0x27, 0x2d, 0xc5, 0x10, 0x51, 0xc6, 0xe4, 0x7a, 0xd4, 0xfa, 0x40, 0x3b, 0x02,
    0xb4, 0x51, 0x0b, 0x64, 0x7a, 0xe3, 0xd1, 0x77, 0x0b, 0xac, 0x03, 0x26, 0xa8, 0x05, 0xbb, 0xef, 0xd4, 0x80, 0x56, 0xc8, 0xc1, 0x21, 0xbd, 0xb8};

int log2_u64(uint64_t n) { int l = 0; while (((uint64_t)1 << l) < n) l++; return l; }

}  // namespace

extern "C" {

pm_rng* pm_rng_seed_from_u64(uint64_t seed) { return new pm_rng(StdRng::seed_from_u64(seed)); }
pm_rng* pm_rng_from_seed(const uint8_t seed[32]) { return new pm_rng(StdRng(seed)); }
void pm_rng_free(pm_rng* r) { delete r; }
uint64_t pm_rng_next_u64(pm_rng* r) { return r->rng.next_u64(); }
void pm_rng_fr_rand(pm_rng* r, uint8_t out[PM_FR_BYTES]) { r->rng.fr_rand().to_wire(out); }

int pm_merlin_test_vector(uint8_t out[32]) {
    // merlin's "equivalence_simple" transcript: new("test protocol"), append("some label","some data"), challenge("challenge")
    MerlinTranscript t("test protocol");
    const char* data = "some data";
    t.append_message("some label", reinterpret_cast<const uint8_t*>(data), 9);
    t.challenge_bytes("challenge", out, 32);
    return PM_OK;
}

int pm_polymath_setup(const pm_r1cs_view* r1cs, pm_rng* rng, pm_ctx** ctx_out, uint8_t vk_out[392]) {
    return pm_polymath_setup_sharded(r1cs, rng, 0, 1, ctx_out, vk_out);
}

int pm_allgather_selftest(pm_allgather_fn allgather, void* user, int rank, int world) {
    if (!allgather || world < 1 || rank < 0 || rank >= world) { pm::set_last_error("bad argument"); return PM_ERR_ARG; }
    const size_t bytes = 192;
    std::vector<uint8_t> send(bytes), recv(bytes * world, 0xee);
    for (size_t i = 0; i < bytes; i++) send[i] = (uint8_t)(rank * 31 + i);
    if (allgather(user, send.data(), bytes, recv.data()) != 0) { pm::set_last_error("allgather callback failed"); return PM_ERR_STATE; }
    for (int r = 0; r < world; r++)
        for (size_t i = 0; i < bytes; i++)
            if (recv[r * bytes + i] != (uint8_t)(r * 31 + i)) { pm::set_last_error("allgather returned wrong data"); return PM_ERR_STATE; }
    return PM_OK;
}

int pm_polymath_setup_sharded(const pm_r1cs_view* r1cs, pm_rng* rng, int rank, int world, pm_ctx** ctx_out, uint8_t vk_out[392]) {
    if (!r1cs || !rng || !ctx_out || !vk_out) { pm::set_last_error("null argument"); return PM_ERR_ARG; }
    const uint64_t m0 = r1cs->num_instance_variables, nr = r1cs->num_r1cs_constraints;
    uint64_t rows = 2 * (m0 + nr), n = 1;
    while (n < rows) n <<= 1;
    const int log_n = log2_u64(n);
    // sample_element_outside_domain: x (generator.rs:72) then z (generator.rs:77)
    auto sample = [&]() {
        for (;;) {
            FrH t = rng->rng.fr_rand();
            if (!(t.pow_u64(n) == FrH::one())) return t;
        }
    };
    FrH x = sample();
    FrH z = sample();
    uint8_t xb[32], zb[32], x_g2[192], z_g2[192];
    x.to_wire(xb);
    z.to_wire(zb);
    int rc = pm_setup_sharded(r1cs, xb, zb, rank, world, ctx_out, x_g2, z_g2);
    if (rc != PM_OK) return rc;
    // VerifyingKey, compressed (data_structures.rs:25-50): e.one_g1, e.one_g2, e.x_g2, e.z_g2, n, m0, sigma, omega
    std::vector<uint8_t> vk;
    vk.insert(vk.end(), G1_GEN_COMPRESSED, G1_GEN_COMPRESSED + 48);
    vk.insert(vk.end(), G2_GEN_COMPRESSED, G2_GEN_COMPRESSED + 96);
    ser_g2_compressed(vk, x_g2);
    ser_g2_compressed(vk, z_g2);
    ser_u64(vk, n);
    ser_u64(vk, m0);
    ser_u64(vk, n + 3);
    ser_fr(vk, fr_group_gen(log_n));
    memcpy(vk_out, vk.data(), 392);
    return PM_OK;
}

static int prove_impl(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, bool upload, pm_rng* rng,
                      pm_allgather_fn allgather, void* user, uint8_t proof_out[176], int transcript = kTranscriptMerlin);

int pm_polymath_prove(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, pm_rng* rng, uint8_t proof_out[176]) {
    return prove_impl(ctx, instance, witness, true, rng, nullptr, nullptr, proof_out);
}
int pm_polymath_prove_resident(pm_ctx* ctx, const uint8_t* instance, pm_rng* rng, uint8_t proof_out[176]) {
    return prove_impl(ctx, instance, nullptr, false, rng, nullptr, nullptr, proof_out);
}
int pm_polymath_prove_transcript(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, pm_rng* rng, int transcript,
                                 uint8_t proof_out[176]) {
    return prove_impl(ctx, instance, witness, true, rng, nullptr, nullptr, proof_out, transcript);
}
int pm_polymath_prove_sharded(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, int upload, pm_rng* rng,
                              pm_allgather_fn allgather, void* user, uint8_t proof_out[176]) {
    return prove_impl(ctx, instance, witness, upload != 0, rng, allgather, user, proof_out);
}

static int prove_impl(pm_ctx* ctx, const uint8_t* instance, const uint8_t* witness, bool upload, pm_rng* rng,
                      pm_allgather_fn allgather, void* user, uint8_t proof_out[176], int transcript) {
    if (!ctx || !instance || !rng || !proof_out) { pm::set_last_error("null argument"); return PM_ERR_ARG; }
    if (transcript < kTranscriptMerlin || transcript > kTranscriptBlake3) { pm::set_last_error("unknown transcript"); return PM_ERR_ARG; }
    int rank = 0, world = 1;
    if (pm_ctx_shard(ctx, &rank, &world) != PM_OK) return PM_ERR_ARG;
    const bool collective = !allgather && pm_ctx_has_collective(ctx);    // NCCL all-gather inside the phases
    if (world > 1 && !allgather && !collective) {
        pm::set_last_error("sharded context needs an all-gather callback or an attached communicator (pm_ctx_attach_nccl)");
        return PM_ERR_ARG;
    }
    // gather `bytes` from every rank (identity when unsharded)
    auto gather = [&](const uint8_t* send, size_t bytes, std::vector<uint8_t>& recv) -> int {
        recv.resize(bytes * world);
        if (world == 1) { memcpy(recv.data(), send, bytes); return PM_OK; }
        if (allgather(user, send, bytes, recv.data()) != 0) { pm::set_last_error("all-gather callback failed"); return PM_ERR_STATE; }
        return PM_OK;
    };
    std::vector<uint8_t> gathered;
    uint64_t n = 0, sigma = 0, cols = 0;
    int rc = pm_ctx_dims(ctx, &n, &sigma, &cols);
    if (rc != PM_OK) return rc;
    uint64_t m0 = 0;
    {
        uint64_t lcs_len = 0;
        rc = pm_ctx_key_len(ctx, 5, &lcs_len);
        if (rc != PM_OK) return rc;
        m0 = cols - lcs_len;
    }
    (void)rank;
    // The device work of phase 1 needs r_a, which the reference draws after the polynomial work and after
    // its divisibility / degree asserts (prover.rs:107-110).  Nothing else consumes the RNG in between, so
    // the stream is identical when the proof succeeds; when phase 1 fails (the reference would have panicked
    // BEFORE `F::rand`) the caller's generator is put back to where it was, so it has not advanced either.
    for (uint64_t i = 0; i < m0; i++) {
        uint64_t limbs[4];
        memcpy(limbs, instance + 32 * i, 32);
        if (FrH::geq_mod(limbs)) { pm::set_last_error("instance value is not a reduced field element"); return PM_ERR_ARG; }
    }
    if (upload) {
        rc = pm_ctx_set_assignment(ctx, instance, witness);
        if (rc != PM_OK) return rc;
    }
    uint8_t ra[64], a_g1[96], c_g1[96];
    const StdRng rng_before = rng->rng;
    rng->rng.fr_rand().to_wire(ra);         // r_a coefficient 0
    rng->rng.fr_rand().to_wire(ra + 32);    // r_a coefficient 1
    if (collective) {
        rc = pm_prove_phase1_collective(ctx, ra, a_g1, c_g1);
        if (rc != PM_OK) { rng->rng = rng_before; return rc; }
    } else {
        uint8_t part1[2 * PM_XYZZ_BYTES];
        rc = pm_prove_phase1_partial(ctx, ra, part1);
        if (rc != PM_OK) { rng->rng = rng_before; return rc; }
        rc = gather(part1, sizeof part1, gathered);
        if (rc != PM_OK) return rc;
        rc = pm_prove_phase1_finish(ctx, gathered.data(), world, a_g1, c_g1);
        if (rc != PM_OK) return rc;
    }

    std::vector<FrH> pub(m0);
    for (uint64_t i = 0; i < m0; i++) pub[i] = FrH::from_wire(instance + 32 * i);

    // compute_x1, common.rs:21-30
    FieldTranscript t(transcript, B_POLYMATH);
    std::vector<uint8_t> msg;
    ser_u64(msg, m0);
    for (auto& v : pub) ser_fr(msg, v);
    t.append_message("public_inputs", msg);
    msg.clear();
    ser_u64(msg, 2);
    ser_g1_compressed(msg, a_g1);
    ser_g1_compressed(msg, c_g1);
    t.append_message("commitments", msg);
    FrH x1 = t.challenge("x1");

    FrH y1 = x1.pow_u64(sigma);                         // compute_y1, common.rs:40-42
    FrH y1_alpha = neg_power(y1, MINUS_ALPHA);
    uint8_t x1b[32], y1ab[32], a_at_x1b[32];
    x1.to_wire(x1b);
    y1_alpha.to_wire(y1ab);
    rc = pm_prove_phase2(ctx, x1b, y1ab, a_at_x1b);      // prover.rs:132
    if (rc != PM_OK) return rc;
    FrH a_at_x1 = FrH::from_wire(a_at_x1b);

    FrH y1_gamma = neg_power(y1, MINUS_GAMMA);
    FrH omega = fr_group_gen(log2_u64(n));
    FrH pi_at_x1 = compute_pi_at_x1(n, omega, pub, x1, y1_gamma);
    FrH c_at_x1 = ((a_at_x1 + y1_gamma) * a_at_x1 - pi_at_x1) * y1_alpha.inv();   // common.rs:73-75

    // compute_x2, common.rs:32-37
    msg.clear();
    ser_fr(msg, x1);
    t.append_message("x1", msg);
    msg.clear();
    ser_u64(msg, 2);
    ser_fr(msg, a_at_x1);
    ser_fr(msg, c_at_x1);
    t.append_message("values", msg);
    FrH x2 = t.challenge("x2");

    uint8_t x2b[32], cb[32], d_g1[96];
    x2.to_wire(x2b);
    c_at_x1.to_wire(cb);
    if (collective) {
        rc = pm_prove_phase3_collective(ctx, x2b, cb, d_g1);   // prover.rs:142-229
        if (rc != PM_OK) return rc;
    } else {
        uint8_t part3[PM_XYZZ_BYTES];
        rc = pm_prove_phase3_partial(ctx, x2b, cb, part3);   // prover.rs:142-229
        if (rc != PM_OK) return rc;
        rc = gather(part3, sizeof part3, gathered);
        if (rc != PM_OK) return rc;
        rc = pm_prove_phase3_finish(ctx, gathered.data(), world, d_g1);
        if (rc != PM_OK) return rc;
    }

    // Proof, compressed (data_structures.rs:10-19): a_g1, c_g1, a_at_x1, d_g1
    std::vector<uint8_t> proof;
    ser_g1_compressed(proof, a_g1);
    ser_g1_compressed(proof, c_g1);
    ser_fr(proof, a_at_x1);
    ser_g1_compressed(proof, d_g1);
    memcpy(proof_out, proof.data(), 176);
    return PM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// `Polymath::verify` on the host (src/verifier.rs:19-62) — no device work
// ---------------------------------------------------------------------------------------
namespace {

uint64_t rd_u64(const uint8_t* b) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)b[i] << (8 * i); return v; }

struct VkH {       // VerifyingKey, data_structures.rs:25-50
    G1H one_g1;
    G2H one_g2, x_g2, z_g2;
    uint64_t n, m0, sigma;
    FrH omega;
};
struct ProofH {    // Proof, data_structures.rs:10-19
    G1H a, c, d;
    FrH a_at_x1;
};

// `VerifyingKey::deserialize_compressed`: points validated (on curve, prime-order subgroup), scalars canonical
bool parse_vk(const uint8_t vk[392], VkH& out) {
    // the standard generators (what `generate_proving_key` writes, generator.rs:141-143) need no subgroup check
    if (!g1_decompress(vk, out.one_g1, memcmp(vk, G1_GEN_COMPRESSED, 48) != 0)) return false;
    if (!g2_decompress(vk + 48, out.one_g2, memcmp(vk + 48, G2_GEN_COMPRESSED, 96) != 0)) return false;
    if (!g2_decompress(vk + 144, out.x_g2)) return false;
    if (!g2_decompress(vk + 240, out.z_g2)) return false;
    out.n = rd_u64(vk + 336);
    out.m0 = rd_u64(vk + 344);
    out.sigma = rd_u64(vk + 352);
    return fr_from_canonical(vk + 360, out.omega);
}
bool parse_proof(const uint8_t p[176], ProofH& out) {
    return g1_decompress(p, out.a) && g1_decompress(p + 48, out.c) && fr_from_canonical(p + 96, out.a_at_x1) &&
           g1_decompress(p + 128, out.d);
}
void g1_wire(const G1H& p, uint8_t out[96]) {
    if (p.inf) { memset(out, 0, 96); return; }
    p.x.to_wire(out);
    p.y.to_wire(out + 48);
}

struct Challenges { FrH x1, x2, c_at_x1; };

// verifier.rs:24-42: the transcript is fed exactly like the prover's (common.rs:21-37)
Challenges derive_challenges(const VkH& vk, const std::vector<FrH>& pub, const ProofH& pr, int transcript = kTranscriptMerlin) {
    FieldTranscript t(transcript, B_POLYMATH);
    std::vector<uint8_t> msg;
    ser_u64(msg, pub.size());
    for (auto& v : pub) ser_fr(msg, v);
    t.append_message("public_inputs", msg);
    msg.clear();
    uint8_t w[96];
    ser_u64(msg, 2);
    g1_wire(pr.a, w);
    ser_g1_compressed(msg, w);
    g1_wire(pr.c, w);
    ser_g1_compressed(msg, w);
    t.append_message("commitments", msg);
    Challenges ch;
    ch.x1 = t.challenge("x1");
    FrH y1 = ch.x1.pow_u64(vk.sigma);
    FrH y1_gamma = neg_power(y1, MINUS_GAMMA);
    FrH pi_at_x1 = compute_pi_at_x1(vk.n, vk.omega, pub, ch.x1, y1_gamma);
    FrH y1_alpha = neg_power(y1, MINUS_ALPHA);
    ch.c_at_x1 = ((pr.a_at_x1 + y1_gamma) * pr.a_at_x1 - pi_at_x1) * y1_alpha.inv();   // common.rs:73-75
    msg.clear();
    ser_fr(msg, ch.x1);
    t.append_message("x1", msg);
    msg.clear();
    ser_u64(msg, 2);
    ser_fr(msg, pr.a_at_x1);
    ser_fr(msg, ch.c_at_x1);
    t.append_message("values", msg);
    ch.x2 = t.challenge("x2");
    return ch;
}

// [a]_1 + x2 [c]_1 - (a(x1) + x2 c(x1)) [1]_1     (verifier.rs:44-47)
G1H commitments_minus_evals(const VkH& vk, const ProofH& pr, const Challenges& ch) {
    G1H acc = aff_add(pr.a, jac_to_affine(scalar_mul(pr.c, ch.x2)));
    FrH ev = (pr.a_at_x1 + ch.x2 * ch.c_at_x1).neg();
    return aff_add(acc, jac_to_affine(scalar_mul(vk.one_g1, ev)));
}

// The reference takes typed field elements (`&[F]`, lib.rs:80-86) and cannot alias x with x + r; raw limb
// vectors can, so anything that is not a reduced Montgomery representative is refused (no silent reduction).
bool public_with_one(const uint8_t* public_inputs, size_t num_public, std::vector<FrH>& pub) {
    pub.assign(num_public + 1, FrH::one());               // verifier.rs:26
    for (size_t i = 0; i < num_public; i++) {
        uint64_t limbs[4];
        memcpy(limbs, public_inputs + 32 * i, 32);
        if (FrH::geq_mod(limbs)) return false;
        pub[i + 1] = FrH::from_wire(public_inputs + 32 * i);
    }
    return true;
}

// exceptions (std::bad_alloc, ...) must not cross the C ABI
template <class F>
int host_guarded(F&& f) {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        pm::set_last_error("out of host memory");
        return PM_ERR_STATE;
    } catch (const std::exception& e) {
        pm::set_last_error(e.what());
        return PM_ERR_STATE;
    } catch (...) {
        pm::set_last_error("unknown exception");
        return PM_ERR_STATE;
    }
}

}  // namespace

extern "C" {

int pm_polymath_verify(const uint8_t vk[392], const uint8_t* public_inputs, size_t num_public, const uint8_t proof[176],
                       int* accepted) {
    return pm_polymath_verify_transcript(vk, public_inputs, num_public, proof, kTranscriptMerlin, accepted);
}

int pm_host_hash(int kind, const uint8_t* data, size_t len, uint8_t out[32]) {
    if ((len && !data) || !out) { pm::set_last_error("null argument"); return PM_ERR_ARG; }
    static const uint8_t none = 0;
    if (kind == kTranscriptKeccak256) keccak256(data ? data : &none, len, out);
    else if (kind == kTranscriptBlake3) blake3_hash(data ? data : &none, len, out);
    else { pm::set_last_error("unknown hash"); return PM_ERR_ARG; }
    return PM_OK;
}

int pm_polymath_verify_transcript(const uint8_t vk[392], const uint8_t* public_inputs, size_t num_public, const uint8_t proof[176],
                                  int transcript, int* accepted) {
    if (!vk || !proof || !accepted || (num_public && !public_inputs)) { pm::set_last_error("null argument"); return PM_ERR_ARG; }
    if (transcript < kTranscriptMerlin || transcript > kTranscriptBlake3) { pm::set_last_error("unknown transcript"); return PM_ERR_ARG; }
    *accepted = 0;
    return host_guarded([&]() -> int {
        VkH k;
        ProofH pr;
        if (!parse_vk(vk, k)) { pm::set_last_error("verifying key does not deserialise (ark-serialize SerializationError)"); return PM_ERR_ARG; }
        if (!parse_proof(proof, pr)) { pm::set_last_error("proof does not deserialise (ark-serialize SerializationError)"); return PM_ERR_ARG; }
        std::vector<FrH> pub;
        if (!public_with_one(public_inputs, num_public, pub)) { pm::set_last_error("public input is not a reduced field element"); return PM_ERR_ARG; }
        Challenges ch = derive_challenges(k, pub, pr, transcript);
        G1H lhs = commitments_minus_evals(k, pr, ch);
        G2H x_minus_x1 = aff_add(k.x_g2, jac_to_affine(scalar_mul(k.one_g2, ch.x1.neg())));   // verifier.rs:48
        PairingTerm terms[2] = {{lhs, k.z_g2}, {pr.d.neg(), x_minus_x1}};                      // verifier.rs:50-59
        *accepted = pairing_product_is_one(terms, 2) ? 1 : 0;
        return PM_OK;
    });
}

int pm_polymath_verify_batch(const uint8_t vk[392], size_t count, const uint8_t* public_inputs, size_t num_public,
                             const uint8_t* proofs, const uint8_t seed[32], int* accepted) {
    if (!vk || !accepted || (count && (!proofs || !seed)) || (count && num_public && !public_inputs)) {
        pm::set_last_error("null argument");
        return PM_ERR_ARG;
    }
    *accepted = 0;
    return host_guarded([&]() -> int {
        VkH k;
        if (!parse_vk(vk, k)) { pm::set_last_error("verifying key does not deserialise"); return PM_ERR_ARG; }
        if (count == 0) { *accepted = 1; return PM_OK; }      // the empty conjunction; no seed needed
        // sum_i r_i * (check_i): e(sum r_i L_i, [z]_2) * e(-sum r_i d_i, [x]_2) * e(sum r_i x1_i d_i, [1]_2) == 1: three
        // pairings for any number of proofs.  r_0 = 1 and r_i = 128-bit values from StdRng keyed by
        // BLAKE3(seed | vk | proofs | public inputs): the coefficients are bound to the statements, so a party
        // that knows (or fixes) the caller's seed still cannot choose proofs against known coefficients.
        uint8_t key[32];
        {
            std::vector<uint8_t> bind;
            bind.reserve(32 + 392 + count * (176 + 32 * num_public));
            bind.insert(bind.end(), seed, seed + 32);
            bind.insert(bind.end(), vk, vk + 392);
            bind.insert(bind.end(), proofs, proofs + 176 * count);
            if (num_public) bind.insert(bind.end(), public_inputs, public_inputs + 32 * num_public * count);
            blake3_hash(bind.data(), bind.size(), key);
        }
        StdRng rng(key);
        JacH<FqH> sum_l = JacH<FqH>::infinity(), sum_d = JacH<FqH>::infinity(), sum_dx = JacH<FqH>::infinity();
        for (size_t i = 0; i < count; i++) {
            ProofH pr;
            if (!parse_proof(proofs + 176 * i, pr)) { pm::set_last_error("proof " + std::to_string(i) + " does not deserialise"); return PM_ERR_ARG; }
            std::vector<FrH> pub;
            if (!public_with_one(public_inputs ? public_inputs + 32 * num_public * i : nullptr, num_public, pub)) {
                pm::set_last_error("public input of proof " + std::to_string(i) + " is not a reduced field element");
                return PM_ERR_ARG;
            }
            Challenges ch = derive_challenges(k, pub, pr);
            FrH r = FrH::one();
            if (i) {
                FrH c = FrH::zero();
                c.v[0] = rng.next_u64();
                c.v[1] = rng.next_u64();
                r = c.to_mont();
            }
            jac_add_affine(sum_l, jac_to_affine(scalar_mul(commitments_minus_evals(k, pr, ch), r)));
            jac_add_affine(sum_d, jac_to_affine(scalar_mul(pr.d, r)));
            jac_add_affine(sum_dx, jac_to_affine(scalar_mul(pr.d, r * ch.x1)));
        }
        PairingTerm terms[3] = {{jac_to_affine(sum_l), k.z_g2}, {jac_to_affine(sum_d).neg(), k.x_g2}, {jac_to_affine(sum_dx), k.one_g2}};
        *accepted = pairing_product_is_one(terms, 3) ? 1 : 0;
        return PM_OK;
    });
}

int pm_host_pairing_product_is_one(const uint8_t* g1_points, const uint8_t* g2_points, int count, int* is_one) {
    if (count < 0 || !is_one || (count && (!g1_points || !g2_points))) { pm::set_last_error("bad argument"); return PM_ERR_ARG; }
    std::vector<PairingTerm> terms(count);
    static const uint8_t zero[192] = {0};
    for (int i = 0; i < count; i++) {
        const uint8_t* p = g1_points + 96 * (size_t)i;
        const uint8_t* q = g2_points + 192 * (size_t)i;
        terms[i].p = memcmp(p, zero, 96) == 0 ? G1H::infinity() : G1H{FqH::from_wire(p), FqH::from_wire(p + 48), false};
        terms[i].q = memcmp(q, zero, 192) == 0
                         ? G2H::infinity()
                         : G2H{{FqH::from_wire(q), FqH::from_wire(q + 48)}, {FqH::from_wire(q + 96), FqH::from_wire(q + 144)}, false};
        if (!g1_on_curve(terms[i].p) || !g2_on_curve(terms[i].q)) { pm::set_last_error("point not on the curve"); return PM_ERR_ARG; }
    }
    *is_one = pairing_product_is_one(terms.data(), count) ? 1 : 0;
    return PM_OK;
}

}  // extern "C"
