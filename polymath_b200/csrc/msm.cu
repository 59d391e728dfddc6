// K4 — G1 Pippenger MSM for sm_100a.
//
// Replaces `VariableBaseMSM::msm_unchecked` as called from /root/reference/src/prover.rs:380-384
// (callers :114-121, :229, :330-357).  Pipeline (no host sync; the pair rounds of the two bucket halves and the level
// sums of the reduction use a second stream):
//   1. k_digits<COUNT>   scalar -> canonical -> signed c-bit digits; histogram of (window, bucket)
//   2. k_scan            exclusive prefix of the histogram -> bucket offsets
//   3. k_digits<SCATTER> counting-sort scatter of (point index | sign) by (window, bucket)
//                        (order inside a bucket is irrelevant: group addition commutes); lanes of a warp that hit the
//                        same bucket share one atomic (repeated scalars: SURVEY.md 8d)
//   4a. pair rounds      (large MSMs) batched-AFFINE tree reduction of every bucket's run: round r adds the
//                        entries of a run pairwise (2i, 2i+1) -> half as many affine points, with ONE shared
//                        inversion per round (Montgomery's trick: k_pairs_forward prefix products,
//                        k_invert_up / top / down, k_pairs_backward) = 6 Fq products per addition instead of the 10 of
//                        an XYZZ mixed addition
//   4. k_accumulate      one thread per bucket, XYZZ mixed additions over its (remaining) run;
//                        buckets longer than a threshold are deferred to
//      k_accumulate_heavy / k_heavy_finish (4096-entry chunks, one CTA each, then a per-bucket
//                        sum of the chunk partials) so narrow top windows and skewed scalar
//                        distributions (SURVEY.md §7 "hard parts") spread over the whole GPU
//   5. k_reduce_level    hierarchical running sums over K-bucket segments (K = 8, see below), k_sum_slices
//                        tree-sums each level: a handful of partial sums per bucket set
//   6. host: Horner over the <= 32 window sums (c doublings each) and the affine normalisation
//      (host/g1_host.hpp) — a serial chain of ~255 doublings that costs a GPU thread milliseconds
// Points at infinity ((0,0)) and zero scalars are skipped in step 1/3.
#include "msm.cuh"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>

namespace pm {

namespace {

constexpr int kMaxWindows = kMaxMsmWindows;
constexpr int kMaxRounds = 10;

__device__ __forceinline__ uint32_t window_bits(const uint32_t* s, int pos, int c) {
    int limb = pos >> 5, off = pos & 31;
    if (limb >= 8) return 0;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> off) & ((1u << c) - 1u);
}

// Only the windows of the bucket sets [set_begin, set_end) are emitted (one pass of a large MSM); the signed-digit
// carry still runs through all lower windows.
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(const G1Affine* __restrict__ bases, const Fr* __restrict__ scalars,
                                                size_t n, size_t sc_stride, size_t sc_offset, int c, int nwin, uint32_t nb,
                                                int levels, uint32_t level_stride, int set_begin, int set_end,
                                                uint32_t* __restrict__ counters, uint32_t* __restrict__ sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // every thread of the warp stays to the end (the atomics below are warp-aggregated): `live` marks the ones with work
    bool live = i < n;
    Fr s = Fr::zero();
    if (live) {
        s = scalars[i * sc_stride + sc_offset];
        live = !s.is_zero();
    }
    // infinity bases carry no weight: test the first limbs cheaply, full test only if they vanish
    if (live) {
        const uint4* bp = reinterpret_cast<const uint4*>(bases + i);
        uint4 q = __ldg(bp);
        if ((q.x | q.y | q.z | q.w) == 0) {
            G1Affine b = bases[i];
            if (b.is_inf()) live = false;
        }
    }
    if (!__any_sync(0xffffffffu, live)) return;
    s = s.from_mont();
    uint32_t limbs[8];
#pragma unroll
    for (int k = 0; k < 8; k++) limbs[k] = s.v[k];
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    // windows in batches of four: the four atomics of a batch are in flight together
    const int w_end = min(nwin, set_end * levels);
    for (int w0 = 0; w0 < w_end; w0 += 4) {
        uint32_t slot[4], val[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int w = w0 + k;
            slot[k] = 0xffffffffu;
            if (w < w_end) {
                uint32_t d = window_bits(limbs, w * c, c) + carry;
                uint32_t neg = 0;
                if (d > half) { d = (1u << c) - d; neg = 1; carry = 1; } else { carry = 0; }
                if (live && d != 0 && w / levels >= set_begin) {
                    slot[k] = (uint32_t)(w / levels - set_begin) * nb + (d - 1);
                    val[k] = ((uint32_t)(w % levels) * level_stride + (uint32_t)i) | (neg << 31);
                }
            }
        }
        // Lanes that target the same bucket (repeated scalars — skewed witnesses, SURVEY.md 8d — put most of a warp into
        // ONE bucket per window) elect a leader that reserves all their positions with one atomic.  The MATCH costs
        // (+0.5 ms per prove on uniform digits when issued for every window), so it only runs when a one-shuffle probe
        // sees two neighbouring lanes on the same bucket: never for uniform digits (2^-19 per pair), always under skew.
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t pos[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t nbr = __shfl_down_sync(0xffffffffu, slot[k], 1);
            const bool dup = lane < 31u && slot[k] != 0xffffffffu && nbr == slot[k];
            if (__any_sync(0xffffffffu, dup)) {
                const unsigned peers = __match_any_sync(0xffffffffu, slot[k]);
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (slot[k] != 0xffffffffu && lane == (uint32_t)leader) base = atomicAdd(&counters[slot[k]], (uint32_t)__popc(peers));
                if (SCATTER) pos[k] = __shfl_sync(0xffffffffu, base, leader) + __popc(peers & ((1u << lane) - 1u));
            } else if (slot[k] != 0xffffffffu) {
                const uint32_t old = atomicAdd(&counters[slot[k]], 1u);
                if (SCATTER) pos[k] = old;
            }
        }
        if (SCATTER) {
#pragma unroll
            for (int k = 0; k < 4; k++) if (slot[k] != 0xffffffffu) sorted[pos[k]] = val[k];
        }
    }
}

__device__ __forceinline__ G1Affine load_point(const G1Affine* __restrict__ bases, uint64_t idx) {
    G1Affine p;
    const uint4* src = reinterpret_cast<const uint4*>(bases + idx);
    uint4* dst = reinterpret_cast<uint4*>(&p);
#pragma unroll
    for (int k = 0; k < 6; k++) dst[k] = __ldg(src + k);
    return p;
}

// ---- bucket order ---------------------------------------------------------------------------
// Buckets are walked in order of decreasing length (counting sort on min(length, kLenBins-1)) so that the
// 32 buckets of a warp have nearly equal lengths: removes the divergence of Poisson-distributed lengths.
constexpr uint32_t kLenBins = 2048;
// (warp-aggregated atomics: after the pair rounds nearly all runs share a few lengths)
__global__ void k_len_hist(const uint32_t* __restrict__ offsets, uint32_t total, int shift, uint32_t* __restrict__ hist) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < total;
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    uint32_t len = (offsets[t + 1] - offsets[t]) >> shift;
    const uint32_t bin = kLenBins - 1 - min(len, kLenBins - 1);   // bin 0 = longest
    const unsigned peers = __match_any_sync(mask, bin);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}
__global__ void __launch_bounds__(1024) k_len_scan(uint32_t* __restrict__ hist) {   // exclusive scan of kLenBins entries, in place
    __shared__ uint32_t sh[kLenBins];
    for (uint32_t i = threadIdx.x; i < kLenBins; i += blockDim.x) sh[i] = hist[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < kLenBins; i++) { uint32_t v = sh[i]; sh[i] = run; run += v; }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kLenBins; i += blockDim.x) hist[i] = sh[i];
}
__global__ void k_len_scatter(const uint32_t* __restrict__ offsets, uint32_t total, int shift, uint32_t* __restrict__ cursor,
                              uint32_t* __restrict__ order) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < total;
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    uint32_t len = (offsets[t + 1] - offsets[t]) >> shift;
    const uint32_t bin = kLenBins - 1 - min(len, kLenBins - 1);
    const unsigned peers = __match_any_sync(mask, bin);
    const int leader = __ffs(peers) - 1;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (uint32_t)leader) base = atomicAdd(&cursor[bin], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + __popc(peers & ((1u << lane) - 1u))] = t;
}

// ---- heavy buckets -------------------------------------------------------------------------
// A bucket longer than `heavy_thr` is cut into chunks of kHeavyChunk sorted entries; every chunk
// becomes a task for one CTA of k_accumulate_heavy, and k_heavy_finish adds the chunk partials of
// each heavy bucket.  Narrow top windows (255 mod c bits) and skewed scalar distributions (repeated
// witness values) put n / 2^k points into single buckets, so this path is hit in
// normal operation and must spread over the whole GPU.
constexpr uint32_t kHeavyChunk = 4096;
struct HeavyLists {
    uint2* tasks;        // (bucket, chunk index)
    uint4* heavy;        // (bucket, first task, task count, 0)
    uint32_t* counters;  // [0] = tasks, [1] = heavy buckets
    G1XYZZ* partials;    // one per task
};

__device__ __forceinline__ void defer_heavy(const HeavyLists& hl, uint32_t t, uint32_t len) {
    uint32_t ntask = (len + kHeavyChunk - 1) / kHeavyChunk;
    uint32_t first = atomicAdd(&hl.counters[0], ntask);
    for (uint32_t i = 0; i < ntask; i++) hl.tasks[first + i] = make_uint2(t, i);
    uint32_t slot = atomicAdd(&hl.counters[1], 1u);
    hl.heavy[slot] = make_uint4(t, first, ntask, 0u);
}

// ---- batched-affine pair rounds ----------------------------------------------------------------
// With R rounds planned, every bucket's run in the sorted list starts at a multiple of A = 2^R and is padded
// with sentinel entries (infinity) to a multiple of A (k_scan_tiles pads the counts, k_pad_runs writes the
// sentinels).  Round r then is one flat, bucket-oblivious pass  X_{r+1}[p] = X_r[2p] + X_r[2p+1]  over all
// S_r / 2 slot pairs (S_0 = offsets[total], S_{r+1} = S_r / 2): pairs never straddle a bucket, after R rounds
// bucket t owns X_R[offsets[t] >> R .. offsets[t+1] >> R), and heavy buckets need no special care.
// X_0 is the sorted (index | sign) list into the base table; X_{r>=1} are scratch arrays stored as six planes
// of 16-byte chunks (plane k, element e -> planes[k * cap + e]) so that the loads of a warp coalesce.
// All additions of a round are independent, so their denominators are inverted together (Montgomery's trick,
// two levels):
//   k_pairs_forward   thread = kPairsPerThread slot pairs (strided by the CTA); stores the exclusive prefix
//                     product of its denominators x2 - x1 (planes, like the points), thread total -> T
//   k_batch_invert    T -> T^-1 element-wise (same trick over T, one Fermat inverse per thread)
//   k_pairs_backward  walks the same pairs backwards: 1/d_i = prefix_i * (running inverse), then
//                     lambda = (y2 - y1)/d, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1
// Per slot pair: 1 + 5 Fq products (+ 3/kPairsPerThread for the second level), against 10 for an XYZZ
// mixed addition.  Exceptional pairs (an operand at infinity — all padding —, P + P, P + (-P)) are classified
// identically in both passes (pair_kind) and contribute no denominator, or 2y for a doubling.
// Slot pairs per thread: chosen per round on the host (16..128, a power of two) so that the round still launches a few
// waves of CTAs while the number of thread totals — the input of the latency-bound inversion tree — stays small.
constexpr int kMinPairsPerThread = 16, kMaxPairsPerThread = 128;
__host__ __device__ constexpr uint32_t pair_tile(int ppt) { return 128u * (uint32_t)ppt; }   // slot pairs per CTA
constexpr uint32_t kPadEntry = 0xffffffffu;             // sorted-list sentinel: the point at infinity

struct FqPlanes {      // n field elements as three planes of 16-byte chunks
    uint4* base;
    size_t cap;
    __device__ __forceinline__ Fq load(size_t e) const {
        Fq r;
        uint4* d4 = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int k = 0; k < 3; k++) d4[k] = base[(size_t)k * cap + e];
        return r;
    }
    __device__ __forceinline__ void store(size_t e, const Fq& v) const {
        const uint4* s4 = reinterpret_cast<const uint4*>(&v);
#pragma unroll
        for (int k = 0; k < 3; k++) base[(size_t)k * cap + e] = s4[k];
    }
};
struct PointPlanes {   // n affine points as six planes (x: planes 0-2, y: planes 3-5)
    uint4* base;
    size_t cap;
    __device__ __forceinline__ Fq load_x(size_t e) const { return FqPlanes{base, cap}.load(e); }
    __device__ __forceinline__ Fq load_y(size_t e) const { return FqPlanes{base + 3 * cap, cap}.load(e); }
    __device__ __forceinline__ G1Affine load(size_t e) const { return {load_x(e), load_y(e)}; }
    __device__ __forceinline__ void store(size_t e, const G1Affine& v) const {
        FqPlanes{base, cap}.store(e, v.x);
        FqPlanes{base + 3 * cap, cap}.store(e, v.y);
    }
};

__device__ __forceinline__ Fq load_fq(const Fq* src) {
    Fq r;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int k = 0; k < 3; k++) d4[k] = __ldg(s4 + k);
    return r;
}
__device__ __forceinline__ void store_fq(Fq* dst, const Fq& v) {
    const uint4* s4 = reinterpret_cast<const uint4*>(&v);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int k = 0; k < 3; k++) d4[k] = s4[k];
}

enum PairKind : int { kPairAdd = 0, kPairDouble = 1, kPairFirst = 2, kPairSecond = 3, kPairInfinity = 4 };
__device__ __forceinline__ int pair_kind(const G1Affine& a, const G1Affine& b) {
    if (a.is_inf()) return kPairSecond;     // also both infinite: the result is b = infinity
    if (b.is_inf()) return kPairFirst;
    if (a.x == b.x) return (a.y == b.y && !a.y.is_zero()) ? kPairDouble : kPairInfinity;
    return kPairAdd;
}

// sentinel entries stay at the end of every padded run
__global__ void k_pad_runs(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t total,
                           uint32_t* __restrict__ sorted) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const uint32_t end = offsets[t + 1];
    for (uint32_t k = offsets[t] + counts[t]; k < end; k++) sorted[k] = kPadEntry;
}

// The slot pairs of round r that one launch handles: slots [*s_begin, *s_end) of round 0 (bucket boundaries read from the
// offsets array on the device; s_begin == nullptr means 0) shrink to pairs [(*s_begin >> r) / 2, (*s_end >> r) / 2).
// A whole pass is ONE span; with two spans (the lower and the upper half of the buckets) the rounds of the halves are
// independent pipelines that run on two streams, so the latency-bound inversion of one half overlaps the
// multiplier-bound forward / backward pass of the other.
// Storage: the two spans advance through their rounds at their own pace, and the pair indices of a later round of the
// upper span fall into the range an earlier round of the lower span may still be reading.  So every span addresses
// its round storage (points, prefix products) by the LOCAL pair / slot index, the lower span upwards from 0 and the
// upper span (`rev`) downwards from the end of the same arrays: [0, a_r) and (cap - b_r, cap] never meet
// (a_r + b_r = pairs of the round <= cap), whatever the rounds the two spans are in.
struct PairSpan {
    const uint32_t* s_begin;
    const uint32_t* s_end;
    int rev;
    __device__ __forceinline__ size_t at(size_t cap, size_t local) const { return rev ? cap - 1 - local : local; }
    __device__ __forceinline__ void get(int r, size_t& begin, size_t& count) const {
        begin = s_begin ? ((size_t)(*s_begin >> r) >> 1) : 0;
        count = ((size_t)(*s_end >> r) >> 1) - begin;
    }
};

// operands of slot pair p in round storage: FIRST reads two (index | sign) entries and gathers from the table
template <bool FIRST>
struct PairSource {
    const G1Affine* table;
    const uint32_t* sorted;
    PointPlanes in;
    // the two (index | sign) entries of slot pair p (FIRST only; later rounds have none)
    __device__ __forceinline__ uint2 load_e(size_t p) const {
        if (FIRST) return __ldg(reinterpret_cast<const uint2*>(sorted + 2 * p));
        return make_uint2(0u, 0u);
    }
    // x coordinates only; returns false when an operand is padding (FIRST only)
    // (later rounds read the previous round's output at the span's LOCAL slots 2 pl, 2 pl + 1: PairSpan::at)
    __device__ __forceinline__ bool gather_x(const PairSpan& sp, size_t pl, const uint2& e, Fq& x1, Fq& x2) const {
        if (FIRST) {
            if (e.x == kPadEntry || e.y == kPadEntry) return false;
            x1 = load_fq(&table[e.x & 0x7fffffffu].x);
            x2 = load_fq(&table[e.y & 0x7fffffffu].x);
        } else {
            x1 = in.load_x(sp.at(in.cap, 2 * pl));
            x2 = in.load_x(sp.at(in.cap, 2 * pl + 1));
        }
        return true;
    }
    __device__ __forceinline__ void load_y(const PairSpan& sp, size_t pl, const uint2& e, Fq& y1, Fq& y2) const {
        if (FIRST) {
            y1 = load_fq(&table[e.x & 0x7fffffffu].y);
            y2 = load_fq(&table[e.y & 0x7fffffffu].y);
            if (e.x >> 31) y1 = y1.neg();
            if (e.y >> 31) y2 = y2.neg();
        } else {
            y1 = in.load_y(sp.at(in.cap, 2 * pl));
            y2 = in.load_y(sp.at(in.cap, 2 * pl + 1));
        }
    }
    __device__ __forceinline__ void load_pair(const PairSpan& sp, size_t p, size_t pl, G1Affine& a, G1Affine& b) const {
        if (FIRST) {
            uint2 e = *reinterpret_cast<const uint2*>(sorted + 2 * p);
            a = G1Affine::inf();
            b = G1Affine::inf();
            if (e.x != kPadEntry) {
                a = load_point(table, e.x & 0x7fffffffu);
                if (e.x >> 31) a.y = a.y.neg();
            }
            if (e.y != kPadEntry) {
                b = load_point(table, e.y & 0x7fffffffu);
                if (e.y >> 31) b.y = b.y.neg();
            }
        } else {
            a = in.load(sp.at(in.cap, 2 * pl));
            b = in.load(sp.at(in.cap, 2 * pl + 1));
        }
    }
};


// Software-pipelined: the (index | sign) entries are fetched two pairs ahead and the x coordinates one pair ahead of
// the product chain, so the two dependent memory latencies of the first round (entry -> table gather, random 96-byte
// records of a multi-GB table) overlap the multiplication of the previous pair instead of serialising with it.
template <bool FIRST, int MINB = 4>
__global__ void __launch_bounds__(128, MINB) k_pairs_forward(PairSource<FIRST> src, PairSpan span, int r, int ppt,
                                                             FqPlanes prefix, Fq* __restrict__ T) {
    size_t first, count;
    span.get(r, first, count);
    const size_t base = (size_t)blockIdx.x * pair_tile(ppt);
    if (base >= count) return;
    const size_t npairs = first + count;                 // pairs are addressed globally, [first, npairs)
    const size_t p0 = first + base + threadIdx.x;
    auto pair_at = [&](int i) { return p0 + (size_t)i * 128; };
    auto live = [&](int i) { return i < ppt && pair_at(i) < npairs; };
    Fq acc = Fq::one();
    uint2 e_cur = make_uint2(0u, 0u), e_nxt = make_uint2(0u, 0u);
    Fq x1, x2;
    bool has = false;
    if (live(0)) e_cur = src.load_e(pair_at(0));
    if (live(1)) e_nxt = src.load_e(pair_at(1));
    if (live(0)) has = src.gather_x(span, pair_at(0) - first, e_cur, x1, x2);
    for (int i = 0; i < ppt; i++) {
        const size_t p = pair_at(i);
        if (p >= npairs) break;
        // issue the loads of the following pairs before touching the multiplier
        uint2 e_nn = make_uint2(0u, 0u);
        if (live(i + 2)) e_nn = src.load_e(pair_at(i + 2));
        Fq nx1, nx2;
        bool nhas = false;
        if (live(i + 1)) nhas = src.gather_x(span, pair_at(i + 1) - first, e_nxt, nx1, nx2);
        Fq d;
        if (has) {
            d = x2 - x1;
            if (d.is_zero() || x1.is_zero() || x2.is_zero()) {
                // rare: decide with the full points, exactly as the backward pass will
                G1Affine a, b;
                a.x = x1;
                b.x = x2;
                src.load_y(span, p - first, e_cur, a.y, b.y);
                const int kind = pair_kind(a, b);
                if (kind == kPairDouble) d = a.y.dbl();
                else if (kind != kPairAdd) has = false;
            }
        }
        prefix.store(span.at(prefix.cap, p - first), acc);
        if (has) acc = fq_mul_call(acc, d);
        e_cur = e_nxt;
        e_nxt = e_nn;
        x1 = nx1;
        x2 = nx2;
        has = nhas;
    }
    store_fq(T + (size_t)blockIdx.x * 128 + threadIdx.x, acc);
}

// ---- inversion of the thread totals T of round r (all non-zero), Montgomery's trick over a small tree ----
// Level sizes: n_0 = thread totals of k_pairs_forward, n_{l+1} = ceil(n_l / fan_l).  k_invert_up multiplies
// fan_l strided elements of level l into one element of level l+1 (exclusive prefixes kept), k_invert_top
// takes the binary-Euclid inverse of the top elements, k_invert_down walks back.  3 products per element and
// level.  A dependent Fq product costs ~1.4 us in a lone warp (its carry chains serialise ~500 instructions), so the
// tree is LATENCY-bound: 3 x (fan_0 + fan_1) product latencies + the inverse (~90 us), whatever the size.  Measured
// and rejected in round 2 (profiles/r2_summary.md): block-wide shared-memory trees with the tile held in registers
// (one launch per level, but 1100 CTAs of 512 threads x 123 us = 0.9 ms).  The latency is hidden instead: the bucket
// range is cut in two halves whose rounds run on two streams (MsmEngine::run), so one half's inversion overlaps
// the other half's forward / backward pass.
constexpr uint32_t kInvFan = 16;       // level 0: many short chains side by side
constexpr int kInvLevels = 2;
// Fan-in of level l.  Level 1 works on ~1/32 of the totals with a few hundred threads: there a SHORT chain (fan 4) and
// more elements for the parallel inverses of the top (thousands instead of hundreds, still under one wave) cut the
// critical path from 32 + 64 dependent products to 4 + 8 (PM_INV_FAN1 overrides; profiles/r2_summary.md).
// Measured, S-mimc(2^20) phases 1 + 3: fans 32 / 32: 65.38 ms, 32 / 4: 64.83, 16 / 4: 64.74, 8 / 4: 64.80 (as rank 0 of 8: 15.44 / 15.24 / 15.17 / 15.11).
// fan1 packs both: low 16 bits = fan of level >= 1, high 16 bits = fan of level 0 (0 = kInvFan)
__host__ __device__ __forceinline__ uint32_t inv_fan(int level, uint32_t fan1) {
    return level == 0 ? ((fan1 >> 16) ? (fan1 >> 16) : kInvFan) : (fan1 & 0xffffu);
}
__device__ __forceinline__ size_t invert_level_size(const PairSpan& span, int r, int ppt, int level, uint32_t fan1) {
    size_t first, count;
    span.get(r, first, count);
    size_t n = (count + pair_tile(ppt) - 1) / pair_tile(ppt) * 128;
    for (int l = 0; l < level; l++) n = (n + inv_fan(l, fan1) - 1) / inv_fan(l, fan1);
    return n;
}
__global__ void __launch_bounds__(128) k_invert_up(const Fq* __restrict__ lo, Fq* __restrict__ pre, Fq* __restrict__ hi,
                                                   PairSpan span, int r, int ppt, int level, uint32_t fan1) {
    const size_t n = invert_level_size(span, r, ppt, level, fan1), m = (n + inv_fan(level, fan1) - 1) / inv_fan(level, fan1);
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    Fq acc = Fq::one();
    for (size_t idx = j; idx < n; idx += m) {
        store_fq(pre + idx, acc);
        acc = fq_mul_call(acc, load_fq(lo + idx));
    }
    store_fq(hi + j, acc);
}
__global__ void __launch_bounds__(128) k_invert_top(Fq* __restrict__ top, PairSpan span, int r, int ppt, int level, uint32_t fan1) {
    const size_t n = invert_level_size(span, r, ppt, level, fan1);
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    store_fq(top + j, load_fq(top + j).inv());
}
__global__ void __launch_bounds__(128) k_invert_down(Fq* __restrict__ lo, const Fq* __restrict__ pre, const Fq* __restrict__ hi,
                                                     PairSpan span, int r, int ppt, int level, uint32_t fan1) {
    const size_t n = invert_level_size(span, r, ppt, level, fan1), m = (n + inv_fan(level, fan1) - 1) / inv_fan(level, fan1);
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    Fq inv = load_fq(hi + j);
    const size_t cnt = (n - j + m - 1) / m;
    for (size_t k = cnt; k-- > 0;) {
        const size_t idx = j + k * m;
        Fq t = load_fq(lo + idx);
        store_fq(lo + idx, fq_mul_call(inv, load_fq(pre + idx)));
        inv = fq_mul_call(inv, t);
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(128, 4) k_pairs_backward(PairSource<FIRST> src, PairSpan span, int r, int ppt,
                                                           FqPlanes prefix, const Fq* __restrict__ Tinv, PointPlanes next) {
    size_t first, count;
    span.get(r, first, count);
    if ((size_t)blockIdx.x * pair_tile(ppt) >= count) return;
    const size_t npairs = first + count;
    const size_t base = first + (size_t)blockIdx.x * pair_tile(ppt);
    // threads of a partial last tile that own no pair hold T = 1
    Fq inv = load_fq(Tinv + (size_t)blockIdx.x * 128 + threadIdx.x);
    for (int i = ppt; i-- > 0;) {
        const size_t p = base + (size_t)i * 128 + threadIdx.x;
        if (p >= npairs) continue;
        G1Affine a, b;
        src.load_pair(span, p, p - first, a, b);
        G1Affine res;
        Fq d = b.x - a.x, num;
        int kind = kPairAdd;
        if (d.is_zero() || a.x.is_zero() || b.x.is_zero()) kind = pair_kind(a, b);
        if (kind == kPairAdd) {
            num = b.y - a.y;
        } else if (kind == kPairDouble) {
            d = a.y.dbl();
            Fq xx = fq_sqr_call(a.x);
            num = xx.dbl() + xx;
        }
        if (kind <= kPairDouble) {
            Fq inv_d = fq_mul_call(inv, prefix.load(span.at(prefix.cap, p - first)));
            inv = fq_mul_call(inv, d);
            Fq lam = fq_mul_call(num, inv_d);
            res.x = fq_sqr_call(lam) - a.x - b.x;
            res.y = fq_mul_call(lam, a.x - res.x) - a.y;
        } else if (kind == kPairFirst) {
            res = a;
        } else if (kind == kPairSecond) {
            res = b;
        } else {
            res = G1Affine::inf();
        }
        next.store(span.at(next.cap, p - first), res);
    }
}

// One thread per (window, bucket): walk the bucket's sorted run with XYZZ mixed additions.
// M selects inlined Fq products or one shared out-of-line copy (smaller loop body: the fully
// inlined body is ~140 KB of SASS and stalls on instruction fetch).  Tuning hook:
// PM_ACC_VARIANT = 3 -> <3, MulInline>, 24 (default) -> <4, MulCall>.
template <int MINB, class M>
__global__ void __launch_bounds__(128, MINB) k_accumulate(const G1Affine* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                          const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ order,
                                                          G1XYZZ* __restrict__ buckets,
                                                          uint32_t total_buckets, uint32_t heavy_thr, HeavyLists hl) {
    uint32_t tix = blockIdx.x * blockDim.x + threadIdx.x;
    if (tix >= total_buckets) return;
    const uint32_t t = order[tix];
    uint32_t beg = offsets[t], end = offsets[t + 1];
    if (end - beg > heavy_thr) {
        defer_heavy(hl, t, end - beg);
        return;
    }
    G1XYZZ acc = G1XYZZ::inf();
    if (beg < end) {
        uint32_t e = sorted[beg];
        G1Affine p = load_point(bases, e & 0x7fffffffu);
        for (uint32_t k = beg; k < end; k++) {
            uint32_t e_next = 0;
            G1Affine p_next;
            if (k + 1 < end) {
                e_next = sorted[k + 1];
                p_next = load_point(bases, e_next & 0x7fffffffu);
            }
            xyzz_madd_t<M>(acc, p, (e >> 31) != 0);
            p = p_next;
            e = e_next;
        }
    }
    buckets[t] = acc;
}

// After R pair rounds: bucket t owns X_R[offsets[t] >> R .. offsets[t+1] >> R); one thread per bucket finishes the
// few remaining points in XYZZ.  Runs still longer than heavy_thr (hot buckets) go to the chunked path.
// Where slot k of X_R lives: the lower span stores its slots upwards from 0, the upper span (slots >= *mid >> R) downwards
// from the end of the planes (PairSpan).  mid == nullptr: one span, identity.
struct RoundSlots {
    PointPlanes pts;
    const uint32_t* mid;
    int R;
    __device__ __forceinline__ G1Affine load(uint32_t k) const {
        if (mid) {
            const uint32_t m = *mid >> R;
            if (k >= m) return pts.load(pts.cap - 1 - (size_t)(k - m));
        }
        return pts.load(k);
    }
};
__global__ void __launch_bounds__(128, 4) k_accumulate_rounds(RoundSlots slots, const uint32_t* __restrict__ offsets,
                                                              const uint32_t* __restrict__ order, int R,
                                                              G1XYZZ* __restrict__ buckets, uint32_t total_buckets,
                                                              uint32_t heavy_thr, HeavyLists hl) {
    uint32_t tix = blockIdx.x * blockDim.x + threadIdx.x;
    if (tix >= total_buckets) return;
    const uint32_t t = order[tix];
    const uint32_t beg = offsets[t] >> R, end = offsets[t + 1] >> R;
    if (end - beg > heavy_thr) {
        defer_heavy(hl, t, end - beg);
        return;
    }
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t k = beg; k < end; k++) xyzz_madd_t<MulCall>(acc, slots.load(k), false);
    buckets[t] = acc;
}

// One CTA per task: strided per-thread sums over one chunk, then a shared-memory tree.
// ROUNDS: the chunk is a range of X_R (after R pair rounds) instead of the sorted list.
template <bool ROUNDS>
__global__ void __launch_bounds__(256) k_accumulate_heavy(const G1Affine* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                          RoundSlots slots, int R,
                                                          const uint32_t* __restrict__ offsets, HeavyLists hl) {
    extern __shared__ uint4 smem_raw[];
    G1XYZZ* sh = reinterpret_cast<G1XYZZ*>(smem_raw);
    const uint32_t ntasks = hl.counters[0];
    for (uint32_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
        uint2 tk = hl.tasks[task];
        const uint32_t run_beg = ROUNDS ? offsets[tk.x] >> R : offsets[tk.x];
        const uint32_t run_end = ROUNDS ? offsets[tk.x + 1] >> R : offsets[tk.x + 1];
        uint32_t beg = run_beg + tk.y * kHeavyChunk;
        uint32_t end = min(run_end, beg + kHeavyChunk);
        G1XYZZ acc = G1XYZZ::inf();
        for (uint32_t k = beg + threadIdx.x; k < end; k += blockDim.x) {
            if (ROUNDS) {
                xyzz_madd_t<MulCall>(acc, slots.load(k), false);
            } else {
                uint32_t e = sorted[k];
                G1Affine p = load_point(bases, e & 0x7fffffffu);
                xyzz_madd_t<MulCall>(acc, p, (e >> 31) != 0);
            }
        }
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (uint32_t stride = blockDim.x / 2; stride > 0; stride >>= 1) {
            if (threadIdx.x < stride) {
                G1XYZZ a = sh[threadIdx.x];
                xyzz_add(a, sh[threadIdx.x + stride]);
                sh[threadIdx.x] = a;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) hl.partials[task] = sh[0];
        __syncthreads();
    }
}

// One CTA per heavy bucket: add its chunk partials.
__global__ void __launch_bounds__(128) k_heavy_finish(G1XYZZ* __restrict__ buckets, HeavyLists hl) {
    __shared__ uint4 smem_raw[128 * sizeof(G1XYZZ) / sizeof(uint4)];
    G1XYZZ* sh = reinterpret_cast<G1XYZZ*>(smem_raw);
    const uint32_t nheavy = hl.counters[1];
    for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
        uint4 hb = hl.heavy[h];
        G1XYZZ acc = G1XYZZ::inf();
        for (uint32_t k = threadIdx.x; k < hb.z; k += blockDim.x) xyzz_add(acc, hl.partials[hb.y + k]);
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (uint32_t stride = blockDim.x / 2; stride > 0; stride >>= 1) {
            if (threadIdx.x < stride) {
                G1XYZZ a = sh[threadIdx.x];
                xyzz_add(a, sh[threadIdx.x + stride]);
                sh[threadIdx.x] = a;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) buckets[hb.x] = sh[0];
        __syncthreads();
    }
}

// ---- bucket reduction: sum_b (b + 1) * B_b per bucket set, hierarchically ---------------------
// With b = hi*K + lo:  sum_b (b + off) X_b = sum_hi T_hi + K * sum_hi hi * R_hi,  where for segment hi
// T_hi = sum_lo (lo + off) X_{hi,lo} and R_hi = sum_lo X_{hi,lo} (one running-sum pass, 2K additions).
// The second term is the same problem on the K-times shorter array R with off = 0, so level j
// contributes 2^(kbits*j) * sum(T_j); the plain sums of the T arrays are tree-reduced and the few
// level sums are combined on the host (combine_levels).  No data-dependent scalar multiplications.
// Level radix K = 2^red_bits.  One thread walks a segment with 2 K dependent XYZZ additions, the array shrinks K-fold per
// level; the first level is throughput-bound (2 additions per bucket = 0.5 ms of Fq products for 2^19 buckets), the later
// ones latency-bound, so a smaller K shortens the tail (sum_levels 2 K + top additions) at the price of more launches.
// Measured on S-mimc(2^20), one GPU / as rank 0 of 8 (phase 1 + 3): K = 16: 66.5 / 16.2 ms, K = 8: 66.6 / 15.7 ms,
// K = 4: 67.1 / 16.1 ms (profiles/r2_summary.md).  The plain sums of the T arrays run on a side stream.
// PM_RED_BITS overrides (tuning hook).
inline int red_bits() {
    static int bits = -1;
    if (bits < 0) {
        const char* v = getenv("PM_RED_BITS");
        bits = v ? atoi(v) : 3;
        if (bits < 1 || bits > 6) bits = 3;
    }
    return bits;
}

// X: [ngroups][m];  T, R: [ngroups][mseg]
template <bool FIRST>
__global__ void __launch_bounds__(128) k_reduce_level(const G1XYZZ* __restrict__ X, uint32_t m, uint32_t mseg, uint32_t ngroups,
                                                      uint32_t seg_len, G1XYZZ* __restrict__ T, G1XYZZ* __restrict__ R) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= mseg * ngroups) return;
    uint32_t g = t / mseg, sgm = t % mseg;
    uint32_t lo = sgm * seg_len;
    uint32_t hi = min(lo + seg_len, m);
    const G1XYZZ* xs = X + (size_t)g * m;
    G1XYZZ running = G1XYZZ::inf(), acc = G1XYZZ::inf();
    for (uint32_t b = hi; b-- > lo;) {
        xyzz_add(running, xs[b]);
        if (FIRST || b > lo) xyzz_add(acc, running);
    }
    T[t] = acc;
    R[t] = running;
}

// out[g * out_per_group + s] = sum of in[g][s*slice .. (s+1)*slice)
__global__ void __launch_bounds__(128) k_sum_slices(const G1XYZZ* __restrict__ in, uint32_t pitch, uint32_t count, uint32_t slice,
                                                    uint32_t out_per_group, uint32_t out_stride, G1XYZZ* __restrict__ out) {
    __shared__ uint4 smem_raw[128 * sizeof(G1XYZZ) / sizeof(uint4)];
    G1XYZZ* sh = reinterpret_cast<G1XYZZ*>(smem_raw);
    const uint32_t g = blockIdx.x / out_per_group, sl = blockIdx.x % out_per_group;
    const G1XYZZ* src = in + (size_t)g * pitch;
    const uint32_t beg = sl * slice, end = min(beg + slice, count);
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t k = beg + threadIdx.x; k < end; k += blockDim.x) xyzz_add(acc, src[k]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t stride = blockDim.x / 2; stride > 0; stride >>= 1) {
        if (threadIdx.x < stride) {
            G1XYZZ a = sh[threadIdx.x];
            xyzz_add(a, sh[threadIdx.x + stride]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)g * out_stride + sl] = sh[0];
}

// Top of the reduction (arrays of a few thousand elements at most): sum_i (i + off) X_i = [off] * sum_i X_i +
// sum_bit 2^bit * sum_{i : bit set} X_i.  One CTA per (bucket set, masked sum); the masked sums go to the host.
constexpr uint32_t kTopTarget = 2048;   // 8 serial + 8 tree additions per thread of k_reduce_top
__global__ void __launch_bounds__(256) k_reduce_top(const G1XYZZ* __restrict__ X, uint32_t m, uint32_t nsums, int with_ones,
                                                    uint32_t out_stride, G1XYZZ* __restrict__ out) {
    __shared__ uint4 smem_raw[256 * sizeof(G1XYZZ) / sizeof(uint4)];
    G1XYZZ* sh = reinterpret_cast<G1XYZZ*>(smem_raw);
    const uint32_t g = blockIdx.x / nsums, s = blockIdx.x % nsums;
    const G1XYZZ* xs = X + (size_t)g * m;
    const bool all = with_ones && s == 0;
    const uint32_t bit = s - (with_ones ? 1u : 0u);
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
        if (all || ((i >> bit) & 1u)) xyzz_add(acc, xs[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t stride = blockDim.x / 2; stride > 0; stride >>= 1) {
        if (threadIdx.x < stride) {
            G1XYZZ a = sh[threadIdx.x];
            xyzz_add(a, sh[threadIdx.x + stride]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)g * out_stride + s] = sh[0];
}

// ---- parallel exclusive scan of the histogram (three small kernels) ---------------------------
constexpr uint32_t kScanTile = 4096;   // elements per CTA (1024 threads x 4)
// pad_mask = A - 1: every count is rounded up to a multiple of A (pair rounds), 0 = plain scan
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t* __restrict__ counts, uint32_t total, uint32_t pad_mask,
                                                     uint32_t* __restrict__ offsets, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t idx = blockIdx.x * kScanTile + tid * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (idx + k < total) ? ((counts[idx + k] + pad_mask) & ~pad_mask) : 0;
    uint32_t local = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        warp_sums[lane] = wi - ws;
        if (lane == 31) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t excl = warp_sums[wid] + incl - local;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (idx + k < total) offsets[idx + k] = excl;
        excl += v[k];
    }
}
// exclusive scan of the tile sums by one CTA (ntiles <= a few thousand); tile_sums[ntiles] = grand total
__global__ void __launch_bounds__(1024) k_scan_tile_sums(uint32_t* __restrict__ tile_sums, uint32_t ntiles) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        uint32_t i = base + tid;
        uint32_t v = i < ntiles ? tile_sums[i] : 0, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += t;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        uint32_t excl = carry_s + warp_sums[wid] + incl - v;
        if (i < ntiles) tile_sums[i] = excl;
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) tile_sums[ntiles] = carry_s;
}
__global__ void __launch_bounds__(1024) k_scan_apply(uint32_t total, const uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                                                     uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursors) {
    const uint32_t idx = blockIdx.x * kScanTile + threadIdx.x * 4;
    const uint32_t base = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (idx + k < total) {
            uint32_t o = offsets[idx + k] + base;
            offsets[idx + k] = o;
            cursors[idx + k] = o;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[total] = tile_sums[ntiles];
}

// out[i] = 2^c * in[i] in XYZZ (levels table construction)
__global__ void __launch_bounds__(128) k_level_up(const G1Affine* __restrict__ in, size_t count, int c, G1XYZZ* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    G1Affine p = in[i];
    G1XYZZ acc = G1XYZZ::from_affine(p);
    for (int k = 0; k < c; k++) xyzz_dbl(acc);
    out[i] = acc;
}

}  // namespace

void launch_batch_to_affine(const G1XYZZ* in, size_t n, G1Affine* out, cudaStream_t stream);   // fixed_base.cu

void launch_build_levels(G1Affine* bases, size_t count, int levels, size_t stride, int c, cudaStream_t stream) {
    if (levels <= 1 || count == 0) return;
    const size_t chunk = (size_t)1 << 22;
    DevBuf scratch;
    G1XYZZ* tmp = scratch.as<G1XYZZ>(count < chunk ? count : chunk);
    for (int l = 1; l < levels; l++) {
        for (size_t lo = 0; lo < count; lo += chunk) {
            size_t cnt = (count - lo) < chunk ? (count - lo) : chunk;
            k_level_up<<<ceil_div(cnt, 128), 128, 0, stream>>>(bases + (size_t)(l - 1) * stride + lo, cnt, c, tmp);
            PM_LAUNCH_CHECK();
            launch_batch_to_affine(tmp, cnt, bases + (size_t)l * stride + lo, stream);
        }
    }
    PM_CUDA(cudaStreamSynchronize(stream));   // scratch dies here
}

static int g_tuning_rounds = -1;
void MsmEngine::set_tuning(int rounds) { g_tuning_rounds = rounds; }

int MsmEngine::choose_window(size_t n) {
    // Empirical optimum on B200 (profiles/msm_window_sweep_r1.jsonl): the bucket-reduction tail grows
    // with 2^(c-1) * ceil(256/c) while the additions shrink with ceil(256/c); c = 16 also makes the 16
    // windows tile the 256-bit scalar exactly (no narrow top window).
    if (n < 64) return 4;
    if (n < 1024) return 8;
    if (n < ((size_t)1 << 14)) return 10;
    if (n < (size_t)92682) return 12;      // 2^16.5
    if (n < (size_t)741455) return 14;     // 2^19.5
    return 16;
}

MsmEngine::~MsmEngine() {
    if (side_stream_) cudaStreamDestroy(side_stream_);
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_join_) cudaEventDestroy(ev_join_);
    if (ev_acc_begin) cudaEventDestroy(ev_acc_begin);
    if (ev_acc_end) cudaEventDestroy(ev_acc_end);
    if (ev_bwd_begin) cudaEventDestroy(ev_bwd_begin);
    if (ev_bwd_end) cudaEventDestroy(ev_bwd_end);
}

static inline size_t entries_of(size_t n, int nwin) { return n * (size_t)nwin; }

void MsmEngine::ensure_side_stream() {
    if (side_stream_) return;
    PM_CUDA(cudaStreamCreateWithFlags(&side_stream_, cudaStreamNonBlocking));
    PM_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    PM_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
}

MsmEngine::Shape MsmEngine::run(const G1Affine* bases, const Fr* scalars, size_t n, G1XYZZ* winsums, cudaStream_t stream,
                                MsmConfig cfg, size_t scalar_stride, size_t scalar_offset) {
    if (n == 0) {
        PM_CUDA(cudaMemsetAsync(winsums, 0, sizeof(G1XYZZ), stream));
        return Shape();
    }
    if (n >= ((size_t)1 << 31)) throw CudaError("msm: n must be < 2^31");
    const int c = cfg.c ? cfg.c : choose_window(n);
    const int nwin = (256 + c - 1) / c;
    if (nwin > kMaxWindows) throw CudaError("msm: too many windows");
    const int levels = cfg.levels > 1 ? cfg.levels : 1;
    const int ngroups = (nwin + levels - 1) / levels;          // bucket sets
    if (levels > 1 && ((size_t)levels * cfg.level_stride >= ((size_t)1 << 31) || cfg.level_stride < n))
        throw CudaError("msm: bad precomputed-level layout");
    const uint32_t nb = 1u << (c - 1);
    const uint32_t total = (uint32_t)ngroups * nb;
    // hierarchical reduction geometry: K-ary levels (K = 2^red_bits) down to a top of <= kTopTarget elements, then bit sums
    const int kRedBits = red_bits();
    const uint32_t kRedK = 1u << kRedBits;
    uint32_t lev_m[24];
    int nlev = 0;
    uint32_t m_top = nb;
    for (;;) {
        if (m_top <= 512) break;
        if (m_top <= kTopTarget) {
            // The masked sums make log2(m) half-empty passes over the array (~16 k pipe cycles per warp addition on
            // 592 schedulers); with many bucket sets (the shards of a multi-GPU prove, table-less MSMs) another level
            // (2 K additions of latency, ~17 us each) is cheaper than those passes: descend while it is
            int bits = 0;
            while ((1u << bits) < m_top) bits++;
            const double masked_ms = (double)ngroups * (bits + 1) * (m_top / 32.0 + 12.0) * 16e3 / 592.0 / 1.9e6;
            if (masked_ms <= 2.0 * kRedK * 0.017) break;
        }
        lev_m[nlev++] = m_top;
        m_top = (m_top + kRedK - 1) / kRedK;
    }
    int top_bits = 0;
    while ((1u << top_bits) < m_top) top_bits++;
    const int with_ones = nlev == 0 ? 1 : 0;       // weights start at 1 only when no level ran (off = 1)
    const int ntop = top_bits + with_ones;
    Shape shape;
    shape.c = c * levels;
    shape.nwin = ngroups;
    shape.nsum = nlev + ntop;
    if (shape.nsum > 32 || shape.count() > kMaxMsmSums) throw CudaError("msm: too many partial sums");
    for (int j = 0; j < nlev; j++) shape.shift[j] = (uint8_t)(kRedBits * j);
    for (int s2 = 0; s2 < ntop; s2++)
        shape.shift[nlev + s2] = (uint8_t)(kRedBits * nlev + (with_ones ? (s2 == 0 ? 0 : s2 - 1) : s2));
    size_t seg_total = 0;   // T and R arrays of all levels
    for (int j = 0; j < nlev; j++) seg_total += (size_t)((lev_m[j] + kRedK - 1) / kRedK) * ngroups;
    const uint32_t ntiles = (total + kScanTile - 1) / kScanTile;
    // A run is cut into chunked tasks when walking it serially would approach the whole kernel's duration: one
    // thread adds a point every ~6.4 us, the whole GPU ~2.8 G points/s, so a run of E / 18000 entries already takes
    // as long as everything else together; split from E / 32768 (skewed witnesses: repeated values).
    size_t share = n * (size_t)nwin / 32768;
    // Small MSMs are latency-bound by their LONGEST run (one thread, ~12-25 us per dependent mixed addition): the floor of
    // the threshold follows the mean run length (4x, between 24 and 128 entries) — narrow top windows of a 2k-point MSM put
    // 64 entries into each of 16 buckets, 0.8 ms as serial walks against 0.15 ms as chunk tasks (profiles/r2_summary.md).
    const size_t mean_run = entries_of(n, nwin) / total;
    const size_t thr_floor = mean_run * 4 < 24 ? 24 : mean_run * 4 > 128 ? 128 : mean_run * 4;
    uint32_t heavy_thr = cfg.heavy ? (uint32_t)cfg.heavy : (uint32_t)(share > thr_floor ? share : thr_floor);
    const size_t max_tasks = n * (size_t)nwin / kHeavyChunk + total + 16;
    const size_t entries = n * (size_t)nwin;   // upper bound of the sorted list

    // ---- batched-affine pair rounds (see k_pairs_forward): choose R from the mean run length ----
    const int forced_rounds = cfg.rounds >= 0 ? cfg.rounds : g_tuning_rounds;
    int rounds = 0;
    // The plan (pair rounds, bucket sets per pass) of a shape is computed once per engine: a proving context issues the
    // same few MSM shapes for every proof, its buffers only grow, and planning may query the free device memory.
    int forced_sets_env = 0;
    {
        const char* v = getenv("PM_MSM_SETS_PER_PASS");      // test hook, read on every call
        if (v) forced_sets_env = atoi(v);
    }
    static int rounds_bias_env = -100;
    if (rounds_bias_env == -100) {
        const char* v = getenv("PM_MSM_ROUNDS_BIAS");
        rounds_bias_env = v ? atoi(v) : 0;
    }
    const PlanKey plan_key{n, c, levels, forced_rounds, cfg.rounds_bias + rounds_bias_env,
                           cfg.sets_per_pass > 0 ? cfg.sets_per_pass : forced_sets_env};
    const auto cached_plan = plans_.find(plan_key);
    int sets_per_pass = ngroups;
    if (cached_plan != plans_.end()) {
        rounds = cached_plan->second.first;
        sets_per_pass = cached_plan->second.second;
    } else {
        {
            const double lambda = (double)entries / (double)total;
            if (forced_rounds >= 0) {
                rounds = forced_rounds;
            } else if (entries >= ((size_t)1 << 20)) {
                // Cost per bucket in ns, constants measured on B200 (profiles/r1_e_summary.md): a slot pair of the padded
                // run costs 0.205 (backward) + 0.083 / 0.044 (forward: first round gathers, later rounds stream), an XYZZ
                // addition of what is left 0.38, and every round a fixed ~0.55 ms (inversion tree, launches)
                // (round 2: the two halves of the buckets run their rounds on two streams, which hides most of the ~0.45 ms of
                // latency-bound inversion per round: ~0.2 ms of launches and tails remain)
                static double round_fixed_ns = -1;
                if (round_fixed_ns < 0) {
                    const char* v = getenv("PM_MSM_ROUND_FIXED_NS");
                    round_fixed_ns = v ? atof(v) : 0.2e6;
                }
                double best = 1e300;
                for (int r = 0; r <= 6; r++) {
                    const double a = (double)(1u << r), lp = lambda + (a - 1) / 2, rest = lp / a;
                    double cost = 0.38 * (rest > 1 ? rest - 1 : 0) + r * round_fixed_ns / (double)total;
                    for (int k = 0; k < r; k++) cost += lp / (double)(2u << k) * (0.205 + (k ? 0.044 : 0.083));
                    if (r == 0) cost = 0.365 * (lambda > 1 ? lambda - 1 : 0);
                    if (cost < best) { best = cost; rounds = r; }
                }
                rounds += cfg.rounds_bias + rounds_bias_env;
                if (rounds < 0) rounds = 0;
            }
            if (rounds > kMaxRounds) rounds = kMaxRounds;
            // The bucket sets are independent: a large MSM runs them in passes of `sets_per_pass`, so that the padded list
            // stays addressable (32-bit offsets) and the pair-round workspace fits into the device memory that is still
            // free (plus what this engine already holds for the purpose).  If not even one set fits: XYZZ walk only.
        }
        {
            const size_t per_set = n * (size_t)levels;      // entries of one bucket set (upper bound)
            auto workspace = [&](int sets, int r) {
                const size_t slots = per_set * sets + (size_t)nb * sets * ((1u << r) - 1);
                return r ? (slots / 2 + 2) * (sizeof(G1Affine) + sizeof(Fq)) + (slots / 4 + 2) * sizeof(G1Affine) + slots * 4 : slots * 4;
            };
            auto addressable = [&](int sets, int r) {
                return per_set * sets + (size_t)nb * sets * ((1u << r) - 1) < ((size_t)1 << 32) - 4096;
            };
            const size_t held = pairs_a_.cap + pairs_b_.cap + prefix_.cap + sorted_.cap;
            // does the plan fit the buffers this engine already holds?  (per buffer: one round needs no second ping-pong
            // array.)  Then nothing is queried: cudaMemGetInfo on a busy device blocks the host for up to tens of
            // milliseconds (measured: 5-60 ms in 1 of 5 proofs while the NTT kernels of the phase were running).
            auto fits_held = [&](int sets, int r) {
                const size_t slots = per_set * sets + (size_t)nb * sets * ((1u << r) - 1);
                const size_t smax = (slots + 1) & ~(size_t)1, cap_a = smax / 2 + 2, cap_b = smax / 4 + 2;
                return sorted_.cap >= (smax + 2) * 4 &&
                       (r == 0 || (pairs_a_.cap >= cap_a * sizeof(G1Affine) && prefix_.cap >= cap_a * sizeof(Fq))) &&
                       (r <= 1 || pairs_b_.cap >= cap_b * sizeof(G1Affine));
            };
            size_t budget = held;
            if (forced_rounds >= 0) {
                budget = ~(size_t)0;
            } else if (fits_held(ngroups, rounds)) {
                const size_t w = workspace(ngroups, rounds);
                budget = w > held ? w : held;
            } else {
                size_t free_b = 0, total_b = 0;
                PM_CUDA(cudaMemGetInfo(&free_b, &total_b));
                budget = held + free_b / 10 * 8;
            }
            int forced_sets = cfg.sets_per_pass;
            if (forced_sets <= 0) forced_sets = forced_sets_env;
            if (forced_sets > 0) sets_per_pass = forced_sets < ngroups ? forced_sets : ngroups;
            while (sets_per_pass > 1 && forced_sets <= 0 &&
                   (!addressable(sets_per_pass, rounds) || workspace(sets_per_pass, rounds) > budget))
                sets_per_pass = (sets_per_pass + 1) / 2;
            if (rounds > 0 && (!addressable(sets_per_pass, rounds) || workspace(sets_per_pass, rounds) > budget)) {
                rounds = 0;                                   // not even one set fits: one XYZZ walk over as many sets as possible
                if (forced_sets <= 0) sets_per_pass = ngroups;
            }
            while (sets_per_pass > 1 && !addressable(sets_per_pass, rounds)) sets_per_pass = (sets_per_pass + 1) / 2;
            if (!addressable(sets_per_pass, rounds)) throw CudaError("msm: too many points per bucket set");
        }
        plans_[plan_key] = std::make_pair(rounds, sets_per_pass);
    }   // plan not cached
    {
        static int debug = -1;
        if (debug < 0) { const char* v = getenv("PM_MSM_DEBUG"); debug = v ? atoi(v) : 0; }
        if (debug) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            fprintf(stderr, "[msm] n=%zu c=%d windows=%d levels=%d sets=%d sets_per_pass=%d rounds=%d free=%.1f GB\n", n, c, nwin,
                    levels, ngroups, sets_per_pass, rounds, free_b / 1e9);
        }
    }
    const uint32_t pad_mask = (1u << rounds) - 1u;
    const uint32_t pass_total = (uint32_t)sets_per_pass * nb;           // buckets of one pass
    const size_t pass_entries = n * (size_t)levels * sets_per_pass;
    const size_t slots_max = (pass_entries + (size_t)pass_total * pad_mask + 1) & ~(size_t)1;   // upper bound of S_0
    const uint32_t pass_tiles = (pass_total + kScanTile - 1) / kScanTile;

    uint32_t* counts = counts_.as<uint32_t>(pass_total + 1);
    uint32_t* offsets = offsets_.as<uint32_t>(pass_total + 1);
    uint32_t* cursors = cursors_.as<uint32_t>((size_t)pass_total + 1 + pass_tiles + 1);
    uint32_t* tile_sums = cursors + pass_total + 1;
    uint32_t* sorted = sorted_.as<uint32_t>(slots_max + 2);
    G1XYZZ* all_buckets = buckets_.as<G1XYZZ>(total);
    G1XYZZ* segs = segs_.as<G1XYZZ>(2 * seg_total + (size_t)ngroups * 256 + 64);
    uint32_t* order = order_.as<uint32_t>((size_t)pass_total + kLenBins);
    uint32_t* len_hist = order + pass_total;
    HeavyLists hl;
    {
        // one allocation: [tasks uint2 | heavy uint4 | partials]
        size_t bytes = max_tasks * sizeof(uint2) + (size_t)pass_total * sizeof(uint4) + max_tasks * sizeof(G1XYZZ) + 64;
        uint8_t* base = heavy_list_.as<uint8_t>(bytes);
        hl.heavy = reinterpret_cast<uint4*>(base);
        hl.partials = reinterpret_cast<G1XYZZ*>(base + (size_t)pass_total * sizeof(uint4));
        hl.tasks = reinterpret_cast<uint2*>(base + (size_t)pass_total * sizeof(uint4) + max_tasks * sizeof(G1XYZZ));
        hl.counters = heavy_count_.as<uint32_t>(2);
    }
    if (time_accumulate && !ev_acc_begin) {
        PM_CUDA(cudaEventCreate(&ev_acc_begin)); PM_CUDA(cudaEventCreate(&ev_acc_end));
        PM_CUDA(cudaEventCreate(&ev_bwd_begin)); PM_CUDA(cudaEventCreate(&ev_bwd_end));
    }
    last_rounds = rounds;
    last_entries = pass_entries;

    for (int set_begin = 0; set_begin < ngroups; set_begin += sets_per_pass) {
        const int set_end = set_begin + sets_per_pass < ngroups ? set_begin + sets_per_pass : ngroups;
        const uint32_t total = (uint32_t)(set_end - set_begin) * nb;      // buckets of this pass (shadows the MSM total)
        const uint32_t ntiles = (total + kScanTile - 1) / kScanTile;
        G1XYZZ* buckets = all_buckets + (size_t)set_begin * nb;
        const bool timed = time_accumulate && set_begin == 0;             // the hooks time the first pass
        PM_CUDA(cudaMemsetAsync(counts, 0, (total + 1) * sizeof(uint32_t), stream));
        PM_CUDA(cudaMemsetAsync(hl.counters, 0, 2 * sizeof(uint32_t), stream));
        const unsigned dgrid = ceil_div(n, 256);
        k_digits<false><<<dgrid, 256, 0, stream>>>(bases, scalars, n, scalar_stride, scalar_offset, c, nwin, nb, levels, (uint32_t)cfg.level_stride, set_begin, set_end, counts, nullptr);
        PM_LAUNCH_CHECK();
        k_scan_tiles<<<ntiles, 1024, 0, stream>>>(counts, total, pad_mask, offsets, tile_sums);
        k_scan_tile_sums<<<1, 1024, 0, stream>>>(tile_sums, ntiles);
        k_scan_apply<<<ntiles, 1024, 0, stream>>>(total, tile_sums, ntiles, offsets, cursors);
        PM_LAUNCH_CHECK();
        k_digits<true><<<dgrid, 256, 0, stream>>>(bases, scalars, n, scalar_stride, scalar_offset, c, nwin, nb, levels, (uint32_t)cfg.level_stride, set_begin, set_end, cursors, sorted);
        PM_LAUNCH_CHECK();
        if (rounds > 0) {
            k_pad_runs<<<ceil_div(total, 256), 256, 0, stream>>>(counts, offsets, total, sorted);
            PM_LAUNCH_CHECK();
        }
        // walk order: by the run length the XYZZ walk will see (after the pair rounds)
        PM_CUDA(cudaMemsetAsync(len_hist, 0, kLenBins * sizeof(uint32_t), stream));
        k_len_hist<<<ceil_div(total, 256), 256, 0, stream>>>(offsets, total, rounds, len_hist);
        k_len_scan<<<1, 1024, 0, stream>>>(len_hist);
        k_len_scatter<<<ceil_div(total, 256), 256, 0, stream>>>(offsets, total, rounds, len_hist, order);
        PM_LAUNCH_CHECK();
        if (timed) PM_CUDA(cudaEventRecord(ev_acc_begin, stream));
        PointPlanes run_pts{nullptr, 0};
        const uint32_t* round_mid = nullptr;      // two spans: first slot (round 0) of the upper one
        if (rounds > 0) {
            const size_t cap_a = slots_max / 2 + 2, cap_b = slots_max / 4 + 2;
            PointPlanes ping{pairs_a_.as<uint4>(6 * cap_a), cap_a};
            PointPlanes pong{rounds > 1 ? pairs_b_.as<uint4>(6 * cap_b) : nullptr, cap_b};
            FqPlanes prefix{prefix_.as<uint4>(3 * cap_a), cap_a};
            // Two spans = the lower and the upper half of this pass's buckets (PairSpan): independent pipelines over
            // disjoint ranges of the same arrays, each with its own thread totals and inversion tree, on two streams.
            static int halves_env = -1;
            if (halves_env < 0) {
                const char* v = getenv("PM_MSM_HALVES");      // tuning hook: 0 = one span on the caller's stream
                halves_env = v ? atoi(v) : 1;
            }
            const int nspans = (halves_env && total >= 2) ? 2 : 1;
            // thread totals and the levels of the inversion tree above them: [T_0 | T_1 | T_2], prefixes [pre_0 | pre_1];
            // sized for the smallest pairs-per-thread; one set per span
            size_t lvl[kInvLevels + 1];
            lvl[0] = (slots_max / 2 + pair_tile(kMinPairsPerThread) - 1) / pair_tile(kMinPairsPerThread) * 128 + 256;
            static uint32_t fan1 = 0;
            if (fan1 == 0) {
                const char* v = getenv("PM_INV_FAN1");      // tuning hook
                const int f = v ? atoi(v) : 4;
                fan1 = (f >= 2 && f <= 64) ? (uint32_t)f : 4;
                const char* v0 = getenv("PM_INV_FAN0");     // tuning hook
                const int f0 = v0 ? atoi(v0) : 0;
                if (f0 >= 2 && f0 <= 64) fan1 |= (uint32_t)f0 << 16;
            }
            for (int l = 0; l < kInvLevels; l++) lvl[l + 1] = (lvl[l] + inv_fan(l, fan1) - 1) / inv_fan(l, fan1) + 1;
            const size_t t_per_span = lvl[0] + lvl[1] + lvl[2], p_per_span = lvl[0] + lvl[1];
            Fq* t_all = tvals_.as<Fq>(t_per_span * nspans);
            Fq* p_all = tpre_.as<Fq>(p_per_span * nspans);
            static int forced_ppt = -1;
            if (forced_ppt < 0) {
                const char* v = getenv("PM_MSM_PAIRS_PER_THREAD");      // tuning hook
                forced_ppt = v ? atoi(v) : 0;
            }
            if (nspans == 2) {
                ensure_side_stream();
                PM_CUDA(cudaEventRecord(ev_fork_, stream));
                PM_CUDA(cudaStreamWaitEvent(side_stream_, ev_fork_, 0));
            }
            const uint32_t mid = total / 2;
            for (int sp = 0; sp < nspans; sp++) {
                cudaStream_t st = sp == 0 ? stream : side_stream_;
                PairSpan span;
                span.s_begin = (nspans == 2 && sp == 1) ? offsets + mid : nullptr;
                span.s_end = (nspans == 2 && sp == 0) ? offsets + mid : offsets + total;
                span.rev = sp;
                Fq* tlev[kInvLevels + 1];
                Fq* plev[kInvLevels];
                tlev[0] = t_all + (size_t)sp * t_per_span;
                plev[0] = p_all + (size_t)sp * p_per_span;
                for (int l = 0; l < kInvLevels; l++) tlev[l + 1] = tlev[l] + lvl[l];
                plev[1] = plev[0] + lvl[0];
                Fq* tvals = tlev[0];
                PointPlanes cur{nullptr, 0};
                for (int r = 0; r < rounds; r++) {
                    // upper bound of the span's pairs (the exact count is read on the device): a half cannot hold more
                    // than the whole, and for uniform digits holds about half — CTAs beyond the exact count exit at once
                    const size_t pairs_max = (slots_max >> r) >> 1;
                    // pairs per thread: keep >= ~4 waves of 128-thread CTAs (148 SMs x 4 resident) in the round
                    int ppt = kMinPairsPerThread;
                    const size_t pairs_est = pairs_max / (size_t)nspans;
                    while (ppt < kMaxPairsPerThread && pairs_est / (pair_tile(ppt) * 2) >= (size_t)sm_count() * 16) ppt *= 2;
                    if (forced_ppt >= kMinPairsPerThread && forced_ppt <= kMaxPairsPerThread) ppt = forced_ppt;
                    const unsigned g = ceil_div(pairs_max, pair_tile(ppt));
                    if (g == 0) break;
                    PointPlanes dst = (r & 1) ? pong : ping;
                    // level sizes for the grids (upper bounds; the kernels derive the exact ones from the span)
                    size_t nl[kInvLevels + 1];
                    nl[0] = (size_t)g * 128;
                    for (int l = 0; l < kInvLevels; l++) nl[l + 1] = (nl[l] + inv_fan(l, fan1) - 1) / inv_fan(l, fan1);
                    auto invert = [&]() {
                        for (int l = 0; l < kInvLevels; l++)
                            k_invert_up<<<ceil_div(nl[l + 1], 128), 128, 0, st>>>(tlev[l], plev[l], tlev[l + 1], span, r, ppt, l, fan1);
                        k_invert_top<<<ceil_div(nl[kInvLevels], 128), 128, 0, st>>>(tlev[kInvLevels], span, r, ppt, kInvLevels, fan1);
                        for (int l = kInvLevels; l-- > 0;)
                            k_invert_down<<<ceil_div(nl[l + 1], 128), 128, 0, st>>>(tlev[l], plev[l], tlev[l + 1], span, r, ppt, l, fan1);
                    };
                    const bool time_bwd = timed && sp == 0 && r == 0;
                    if (r == 0) {
                        PairSource<true> src{bases, sorted, cur};
                        static int fwd_ctas = -1;
                        if (fwd_ctas < 0) {
                            const char* v = getenv("PM_FWD_CTAS");        // tuning hook: resident CTAs of the gather-bound pass
                            fwd_ctas = v ? atoi(v) : 4;
                        }
                        if (fwd_ctas == 6) k_pairs_forward<true, 6><<<g, 128, 0, st>>>(src, span, r, ppt, prefix, tvals);
                        else if (fwd_ctas == 5) k_pairs_forward<true, 5><<<g, 128, 0, st>>>(src, span, r, ppt, prefix, tvals);
                        else if (fwd_ctas == 3) k_pairs_forward<true, 3><<<g, 128, 0, st>>>(src, span, r, ppt, prefix, tvals);
                        else k_pairs_forward<true, 4><<<g, 128, 0, st>>>(src, span, r, ppt, prefix, tvals);
                        invert();
                        if (time_bwd) PM_CUDA(cudaEventRecord(ev_bwd_begin, st));
                        k_pairs_backward<true><<<g, 128, 0, st>>>(src, span, r, ppt, prefix, tvals, dst);
                        if (time_bwd) PM_CUDA(cudaEventRecord(ev_bwd_end, st));
                    } else {
                        PairSource<false> src{bases, sorted, cur};
                        k_pairs_forward<false><<<g, 128, 0, st>>>(src, span, r, ppt, prefix, tvals);
                        invert();
                        k_pairs_backward<false><<<g, 128, 0, st>>>(src, span, r, ppt, prefix, tvals, dst);
                    }
                    PM_LAUNCH_CHECK();
                    cur = dst;
                    launches += 3 + 2 * kInvLevels;
                }
                run_pts = cur;
            }
            if (nspans == 2) {
                PM_CUDA(cudaEventRecord(ev_join_, side_stream_));
                PM_CUDA(cudaStreamWaitEvent(stream, ev_join_, 0));
                round_mid = offsets + mid;
            }
        }
        // runs are cut into chunked tasks only when walking them serially would approach the kernel's duration
        const uint32_t walk_heavy_thr = rounds > 0 && !cfg.heavy ? (heavy_thr >> rounds > 32 ? heavy_thr >> rounds : 32) : heavy_thr;
        {
            static int variant = -1;
            if (variant < 0) {
                const char* v = getenv("PM_ACC_VARIANT");
                variant = v ? atoi(v) : 24;
            }
            const unsigned g = ceil_div(total, 128);
            if (rounds > 0) k_accumulate_rounds<<<g, 128, 0, stream>>>(RoundSlots{run_pts, round_mid, rounds}, offsets, order, rounds, buckets, total, walk_heavy_thr, hl);
            else if (variant == 3) k_accumulate<3, MulInline><<<g, 128, 0, stream>>>(bases, sorted, offsets, order, buckets, total, heavy_thr, hl);
            else k_accumulate<4, MulCall><<<g, 128, 0, stream>>>(bases, sorted, offsets, order, buckets, total, heavy_thr, hl);
            PM_LAUNCH_CHECK();
        }
        if (timed) PM_CUDA(cudaEventRecord(ev_acc_end, stream));
        {
            static bool attr_set = false;
            const int smem = 256 * (int)sizeof(G1XYZZ);
            if (!attr_set) {
                PM_CUDA(cudaFuncSetAttribute(k_accumulate_heavy<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                PM_CUDA(cudaFuncSetAttribute(k_accumulate_heavy<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                attr_set = true;
            }
            if (rounds > 0) k_accumulate_heavy<true><<<4 * sm_count(), 256, smem, stream>>>(bases, sorted, RoundSlots{run_pts, round_mid, rounds}, rounds, offsets, hl);
            else k_accumulate_heavy<false><<<4 * sm_count(), 256, smem, stream>>>(bases, sorted, RoundSlots{run_pts, nullptr, 0}, 0, offsets, hl);
            PM_LAUNCH_CHECK();
            k_heavy_finish<<<2 * sm_count(), 128, 0, stream>>>(buckets, hl);
            PM_LAUNCH_CHECK();
        }
        launches += 9;
    }   // passes over the bucket sets
    {
        const G1XYZZ* X = all_buckets;
        G1XYZZ* cursor = segs;
        G1XYZZ* scratch = segs + 2 * seg_total;     // [ngroups][64] slice partials
        const int nsum = shape.nsum;
        uint32_t m_last = nb;
        // The level kernels form the critical path; the plain sums of the T arrays only feed the host, so they run on
        // the side stream behind an event per level and rejoin before the caller reads winsums.
        ensure_side_stream();
        for (int j = 0; j < nlev; j++) {
            const uint32_t m = lev_m[j], mseg = (m + kRedK - 1) / kRedK;
            G1XYZZ* T = cursor;
            G1XYZZ* R = cursor + (size_t)mseg * ngroups;
            cursor = R + (size_t)mseg * ngroups;
            const unsigned grid = ceil_div((size_t)mseg * ngroups, 128);
            if (j == 0) k_reduce_level<true><<<grid, 128, 0, stream>>>(X, m, mseg, ngroups, kRedK, T, R);
            else k_reduce_level<false><<<grid, 128, 0, stream>>>(X, m, mseg, ngroups, kRedK, T, R);
            PM_LAUNCH_CHECK();
            launches++;
            PM_CUDA(cudaEventRecord(ev_fork_, stream));
            PM_CUDA(cudaStreamWaitEvent(side_stream_, ev_fork_, 0));
            // level sum: winsums[g*nsum + j] = sum_seg T[g][seg]
            if (mseg > 512) {
                // two stages: slices of <= 512 elements per 128-thread CTA (4 serial + 7 tree additions), then
                // the <= 256 slice sums by one CTA per bucket set
                uint32_t slice = (mseg + 255) / 256;
                if (slice < 256) slice = 256;
                const uint32_t nsl = (mseg + slice - 1) / slice;
                k_sum_slices<<<ngroups * nsl, 128, 0, side_stream_>>>(T, mseg, mseg, slice, nsl, 256, scratch);
                k_sum_slices<<<ngroups, 128, 0, side_stream_>>>(scratch, 256, nsl, nsl, 1, nsum, winsums + j);
                launches += 2;
            } else {
                k_sum_slices<<<ngroups, 128, 0, side_stream_>>>(T, mseg, mseg, mseg, 1, nsum, winsums + j);
                launches++;
            }
            PM_LAUNCH_CHECK();
            X = R;
            m_last = mseg;
        }
        k_reduce_top<<<ngroups * ntop, 256, 0, stream>>>(X, m_last, (uint32_t)ntop, with_ones, (uint32_t)nsum, winsums + nlev);
        PM_LAUNCH_CHECK();
        launches++;
        if (nlev > 0) {
            PM_CUDA(cudaEventRecord(ev_join_, side_stream_));
            PM_CUDA(cudaStreamWaitEvent(stream, ev_join_, 0));
        }
    }
    return shape;
}

}  // namespace pm
