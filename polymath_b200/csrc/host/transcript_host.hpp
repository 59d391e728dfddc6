// Host-side Fiat-Shamir transcript and RNG used by the C++ mirror of Polymath::setup/prove.
//   * MerlinFieldTranscript  — /root/reference/src/transcript/merlin.rs:13-36 over merlin 3.0
//     (STROBE-128 / Keccak-f[1600]); challenge rule: 64 bytes -> Fr::from_random_bytes, retry.
//   * StdRng                 — rand 0.8 StdRng = ChaCha12 (seed_from_u64 via PCG32), and
//     ark-ff `Fr::rand` rejection sampling (reference call sites: prover.rs:110, generator.rs:72,77).
// In a Rust integration these stay the caller's own `merlin` / `rand` crates; they exist here so
// that the C++ host mirror can reproduce the reference flow byte for byte.  Product code.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "fp_host.hpp"

namespace pm { namespace host {

// ---- Keccak-f[1600] -------------------------------------------------------------------
inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
        0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    static const int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = PILN[i];
            uint64_t b = st[j];
            st[j] = rotl64(t, ROTC[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[round];
    }
}

// ---- STROBE-128 subset used by merlin ---------------------------------------------------
class Strobe128 {
public:
    explicit Strobe128(const char* protocol_label) {
        memset(state_, 0, sizeof state_);
        const uint8_t init[6] = {1, kRate + 2, 1, 0, 1, 96};
        memcpy(state_, init, 6);
        memcpy(state_ + 6, "STROBEv1.0.2", 12);
        permute();
        meta_ad(reinterpret_cast<const uint8_t*>(protocol_label), strlen(protocol_label), false);
    }
    void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
    void ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
    void prf(uint8_t* out, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(out, n); }

private:
    static constexpr int kRate = 166;
    enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
    uint8_t state_[200];
    uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;

    void permute() {
        uint64_t lanes[25];
        memcpy(lanes, state_, 200);   // little-endian host
        keccak_f1600(lanes);
        memcpy(state_, lanes, 200);
    }
    void run_f() {
        state_[pos_] ^= pos_begin_;
        state_[pos_ + 1] ^= 0x04;
        state_[kRate + 1] ^= 0x80;
        permute();
        pos_ = 0;
        pos_begin_ = 0;
    }
    void absorb(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            state_[pos_] ^= d[i];
            if (++pos_ == kRate) run_f();
        }
    }
    void squeeze(uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            d[i] = state_[pos_];
            state_[pos_] = 0;
            if (++pos_ == kRate) run_f();
        }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;   // continuation of the current operation
        uint8_t old_begin = pos_begin_;
        pos_begin_ = pos_ + 1;
        cur_flags_ = flags;
        const uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
    }
};

class MerlinTranscript {
public:
    explicit MerlinTranscript(const std::string& label) : strobe_("Merlin v1.0") {
        append_message("dom-sep", reinterpret_cast<const uint8_t*>(label.data()), label.size());
    }
    void append_message(const char* label, const uint8_t* msg, size_t n) {
        uint32_t len = (uint32_t)n;
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        strobe_.meta_ad(reinterpret_cast<const uint8_t*>(label), strlen(label), false);
        strobe_.meta_ad(le, 4, true);
        strobe_.ad(msg, n, false);
    }
    void challenge_bytes(const char* label, uint8_t* out, size_t n) {
        uint32_t len = (uint32_t)n;
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        strobe_.meta_ad(reinterpret_cast<const uint8_t*>(label), strlen(label), false);
        strobe_.meta_ad(le, 4, true);
        strobe_.prf(out, n, false);
    }

private:
    Strobe128 strobe_;
};

// transcript/merlin.rs:27-35: draw 64 bytes, `F::from_random_bytes` (first 32 bytes LE, top bit
// cleared, accept iff < r), retry with the same label otherwise.
class MerlinFieldTranscript {
public:
    explicit MerlinFieldTranscript(const std::string& name) : t_(name) {}
    void append_message(const char* label, const std::vector<uint8_t>& msg) { t_.append_message(label, msg.data(), msg.size()); }
    FrH challenge(const char* label) {
        for (;;) {
            uint8_t buf[64];
            t_.challenge_bytes(label, buf, 64);
            buf[31] &= 0x7f;
            uint64_t limbs[4];
            memcpy(limbs, buf, 32);
            if (!FrH::geq_mod(limbs)) return FrH::from_canonical_le(buf);
        }
    }

private:
    MerlinTranscript t_;
};

// ---- Keccak-256 (original padding, as the `sha3` crate's Keccak256) and BLAKE3 ------------------------------
// for the reference's two hash transcripts (transcript/keccak256.rs:16-42, transcript/blake3.rs:16-42)
inline void keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    constexpr size_t kRate = 136;
    uint64_t st[25];
    memset(st, 0, sizeof st);
    uint8_t block[kRate];
    auto absorb_block = [&](const uint8_t* b) {
        for (size_t i = 0; i < kRate / 8; i++) { uint64_t w; memcpy(&w, b + 8 * i, 8); st[i] ^= w; }   // little-endian host
        keccak_f1600(st);
    };
    while (len >= kRate) { absorb_block(data); data += kRate; len -= kRate; }
    memset(block, 0, sizeof block);
    memcpy(block, data, len);
    block[len] ^= 0x01;
    block[kRate - 1] ^= 0x80;
    absorb_block(block);
    memcpy(out, st, 32);
}

namespace blake3_detail {
constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
constexpr int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };
inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline void g(uint32_t* v, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    v[a] = v[a] + v[b] + mx; v[d] = rotr(v[d] ^ v[a], 16);
    v[c] = v[c] + v[d];      v[b] = rotr(v[b] ^ v[c], 12);
    v[a] = v[a] + v[b] + my; v[d] = rotr(v[d] ^ v[a], 8);
    v[c] = v[c] + v[d];      v[b] = rotr(v[b] ^ v[c], 7);
}
// compression function; out = the first eight words (the chaining value / first half of the root output)
inline void compress(const uint32_t cv[8], const uint8_t block[64], uint64_t counter, uint32_t block_len, uint32_t flags, uint32_t out[8]) {
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++) memcpy(&m[i], block + 4 * i, 4);   // little-endian host
    for (int i = 0; i < 8; i++) v[i] = cv[i];
    for (int i = 0; i < 4; i++) v[8 + i] = IV[i];
    v[12] = (uint32_t)counter; v[13] = (uint32_t)(counter >> 32); v[14] = block_len; v[15] = flags;
    for (int round = 0; round < 7; round++) {
        g(v, 0, 4, 8, 12, m[0], m[1]);   g(v, 1, 5, 9, 13, m[2], m[3]);
        g(v, 2, 6, 10, 14, m[4], m[5]);  g(v, 3, 7, 11, 15, m[6], m[7]);
        g(v, 0, 5, 10, 15, m[8], m[9]);  g(v, 1, 6, 11, 12, m[10], m[11]);
        g(v, 2, 7, 8, 13, m[12], m[13]); g(v, 3, 4, 9, 14, m[14], m[15]);
        uint32_t p[16];
        for (int i = 0; i < 16; i++) p[i] = m[PERM[i]];
        memcpy(m, p, sizeof m);
    }
    for (int i = 0; i < 8; i++) out[i] = v[i] ^ v[i + 8];
}
// one chunk (<= 1024 bytes): chained block compressions; `root` marks a single-chunk input
inline void chunk_cv(const uint8_t* data, size_t len, uint64_t chunk_index, bool root, uint32_t out[8]) {
    uint32_t cv[8];
    memcpy(cv, IV, sizeof cv);
    const size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t block[64];
        memset(block, 0, sizeof block);
        const size_t off = b * 64, take = len - off < 64 ? len - off : 64;
        memcpy(block, data + off, take);
        uint32_t flags = (b == 0 ? CHUNK_START : 0) | (b + 1 == nblocks ? CHUNK_END | (root ? ROOT : 0) : 0);
        uint32_t next[8];
        compress(cv, block, chunk_index, (uint32_t)take, flags, next);
        memcpy(cv, next, sizeof cv);
    }
    memcpy(out, cv, sizeof cv);
}
inline void parent_cv(const uint32_t l[8], const uint32_t r[8], bool root, uint32_t out[8]) {
    uint8_t block[64];
    memcpy(block, l, 32);
    memcpy(block + 32, r, 32);
    compress(IV, block, 0, 64, PARENT | (root ? ROOT : 0), out);
}
// left subtree = the largest power-of-two number of whole chunks that leaves at least one byte on the right
inline size_t left_len(size_t len) {
    size_t chunks = (len - 1) / 1024, p = 1;
    while (p * 2 <= chunks) p *= 2;
    return p * 1024;
}
inline void subtree_cv(const uint8_t* data, size_t len, uint64_t chunk_index, bool root, uint32_t out[8]) {
    if (len <= 1024) { chunk_cv(data, len, chunk_index, root, out); return; }
    const size_t ll = left_len(len);
    uint32_t l[8], r[8];
    subtree_cv(data, ll, chunk_index, false, l);
    subtree_cv(data + ll, len - ll, chunk_index + ll / 1024, false, r);
    parent_cv(l, r, root, out);
}
}  // namespace blake3_detail

inline void blake3_hash(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint32_t cv[8];
    blake3_detail::subtree_cv(data, len, 0, true, cv);
    memcpy(out, cv, 32);   // little-endian words
}

enum TranscriptKind { kTranscriptMerlin = 0, kTranscriptKeccak256 = 1, kTranscriptBlake3 = 2 };

// The reference's `Transcript` trait (transcript/mod.rs:17-29) over its three implementations.  The two hash
// transcripts ignore the name, keep label || message bytes, and answer H(transcript || label) reduced mod r
// (`from_be_bytes_mod_order`), after which the transcript IS the digest (keccak256.rs:26-41, blake3.rs:26-41).
class FieldTranscript {
public:
    FieldTranscript(int kind, const std::string& name) : kind_(kind), merlin_(name) {}
    void append_message(const char* label, const std::vector<uint8_t>& msg) {
        if (kind_ == kTranscriptMerlin) { merlin_.append_message(label, msg); return; }
        buf_.insert(buf_.end(), label, label + strlen(label));
        buf_.insert(buf_.end(), msg.begin(), msg.end());
    }
    FrH challenge(const char* label) {
        if (kind_ == kTranscriptMerlin) return merlin_.challenge(label);
        std::vector<uint8_t> in(buf_);
        in.insert(in.end(), label, label + strlen(label));
        uint8_t digest[32];
        if (kind_ == kTranscriptKeccak256) keccak256(in.data(), in.size(), digest);
        else blake3_hash(in.data(), in.size(), digest);
        buf_.assign(digest, digest + 32);
        // from_be_bytes_mod_order: the 256-bit big-endian integer mod r (a Montgomery product by R^2 reduces it)
        FrH v;
        uint8_t le[32];
        for (int i = 0; i < 32; i++) le[i] = digest[31 - i];
        memcpy(v.v, le, 32);
        return v.to_mont();
    }

private:
    int kind_;
    MerlinFieldTranscript merlin_;
    std::vector<uint8_t> buf_;
};

// ---- rand 0.8 StdRng ------------------------------------------------------------------
class StdRng {
public:
    explicit StdRng(const uint8_t seed[32]) { memcpy(key_, seed, 32); }
    static StdRng seed_from_u64(uint64_t state) {   // rand_core 0.6: PCG32 stream fills the seed
        uint8_t seed[32];
        for (int i = 0; i < 8; i++) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            memcpy(seed + 4 * i, &x, 4);
        }
        return StdRng(seed);
    }
    uint32_t next_u32() {
        if (idx_ == 16) refill();
        return buf_[idx_++];
    }
    uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return lo | (hi << 32);
    }
    // ark-ff 0.4 `Fp::rand`: the accepted limbs are the Montgomery representation itself
    FrH fr_rand() {
        for (;;) {
            FrH r;
            for (int i = 0; i < 4; i++) r.v[i] = next_u64();
            r.v[3] &= 0x7fffffffffffffffull;
            if (!FrH::geq_mod(r.v)) return r;
        }
    }

private:
    uint32_t key_[8];
    uint64_t counter_ = 0;
    uint32_t buf_[16];
    int idx_ = 16;

    static uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
    static void qr(uint32_t* s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
    }
    void refill() {   // one ChaCha12 block; 64-bit block counter in words 12-13, stream id 0
        uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        memcpy(init + 4, key_, 32);
        init[12] = (uint32_t)counter_;
        init[13] = (uint32_t)(counter_ >> 32);
        init[14] = init[15] = 0;
        uint32_t s[16];
        memcpy(s, init, sizeof s);
        for (int r = 0; r < 6; r++) {
            qr(s, 0, 4, 8, 12); qr(s, 1, 5, 9, 13); qr(s, 2, 6, 10, 14); qr(s, 3, 7, 11, 15);
            qr(s, 0, 5, 10, 15); qr(s, 1, 6, 11, 12); qr(s, 2, 7, 8, 13); qr(s, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) buf_[i] = s[i] + init[i];
        counter_++;
        idx_ = 0;
    }
};

// ---- ark-serialize / zcash encodings -----------------------------------------------------
inline void ser_u64(std::vector<uint8_t>& out, uint64_t v) { for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i))); }
inline void ser_fr(std::vector<uint8_t>& out, const FrH& v) { uint8_t b[32]; v.to_canonical_le(b); out.insert(out.end(), b, b + 32); }

// 96-byte Montgomery affine point ((0,0) = infinity) -> 48-byte compressed zcash form
inline void ser_g1_compressed(std::vector<uint8_t>& out, const uint8_t pt[96]) {
    FqH x = FqH::from_wire(pt), y = FqH::from_wire(pt + 48);
    uint8_t b[48];
    if (x.is_zero() && y.is_zero()) {
        memset(b, 0, 48);
        b[0] = 0xc0;
    } else {
        uint8_t le[48];
        x.to_canonical_le(le);
        for (int i = 0; i < 48; i++) b[i] = le[47 - i];
        b[0] |= 0x80;
        if (y.canonical_gt_half()) b[0] |= 0x20;
    }
    out.insert(out.end(), b, b + 48);
}
// 192-byte Montgomery affine G2 point (x.c0, x.c1, y.c0, y.c1) -> 96-byte compressed zcash form
inline void ser_g2_compressed(std::vector<uint8_t>& out, const uint8_t pt[192]) {
    FqH x0 = FqH::from_wire(pt), x1 = FqH::from_wire(pt + 48), y0 = FqH::from_wire(pt + 96), y1 = FqH::from_wire(pt + 144);
    uint8_t b[96];
    if (x0.is_zero() && x1.is_zero() && y0.is_zero() && y1.is_zero()) {
        memset(b, 0, 96);
        b[0] = 0xc0;
    } else {
        uint8_t le[48];
        x1.to_canonical_le(le);
        for (int i = 0; i < 48; i++) b[i] = le[47 - i];
        x0.to_canonical_le(le);
        for (int i = 0; i < 48; i++) b[48 + i] = le[47 - i];
        b[0] |= 0x80;
        bool largest = y1.is_zero() ? y0.canonical_gt_half() : y1.canonical_gt_half();
        if (largest) b[0] |= 0x20;
    }
    out.insert(out.end(), b, b + 96);
}

}}  // namespace pm::host
