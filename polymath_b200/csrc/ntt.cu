// K3 — Fr NTT / iNTT for sm_100a (shared-memory staged, four-step layout).
//
// Replaces `Radix2EvaluationDomain::{fft, ifft_in_place}` as called from
// /root/reference/src/prover.rs:239-243 (poly_coeffs) and :315-328 (square_polynomial).
// Natural order in and out; w_N is arkworks' `group_gen` (TWO_ADIC_ROOT squared 32-log_n times).
//
// N = 2^k is split into up to three digits N = N1*N2*N3 (each 2^6..2^11):
//   column pass(es): for every residue of the lower digits, an M-point sub-NTT over a
//       strided digit, done entirely in shared memory by one CTA for a batch of B adjacent
//       columns (B*32-byte contiguous chunks in HBM), then the inter-digit twiddle
//       w^(low*k) is applied on the way out.  The first column pass reads the caller's buffer and
//       writes the engine's scratch buffer, later ones run in place there.
//   row pass: contiguous M-point sub-NTTs; the store performs the digit-reversal transpose
//       (B adjacent outputs per k), so the result lands in natural order — written straight back into
//       the caller's buffer (ping-pong: no copy after the transform).
// Inside a CTA: a tile of 2^9..2^11 elements (2048 from 2^20 on; smaller transforms use smaller tiles so
// that 2^16..2^19 still fill the 148 SMs), tile/8 threads, 8 elements per thread; radix-8 register rounds
// (three butterfly stages between shared-memory exchanges).  Data and twiddles live in shared memory as
// 16-byte limb planes with a multi-level skew (slot i -> i + i/8 + i/64 + i/512), which makes every
// power-of-two stride conflict-free for the 8 lanes of a quarter-warp — the strided twiddle reads of the
// previous layout (a contiguous Fr array) were 17.5 M of 25.3 M shared-memory wavefronts
// (profiles/r1_c_summary.md).  Each distinct twiddle of a round is fetched once per thread (7 per
// radix-8 round instead of 12).
#include "ntt.cuh"
#include "roots.cuh"

namespace pm {

namespace {

constexpr int kCoreBits = 12;        // largest digit: one CTA transforms up to 4096 points in shared memory
constexpr int kMaxTileBits = 12;     // 4096-element tiles (220 KB, one CTA of 512 threads per SM): PM_NTT_BIG_TILE, off by default
constexpr int kStdTileBits = 11;     // 2048-element tiles (110 KB, two CTAs of 256 threads per SM) everywhere else
constexpr int kMinTileBits = 9;
constexpr int kTwBits = 11;          // inter-pass twiddle tables of transforms up to 2^22: 2048 entries per level
// Above 2^22 two levels of 2^ceil(log_n / 2) entries (128 KB .. 2 MB per table, L2-resident) keep the inter-pass twiddle at
// ONE table product: with three 2048-entry levels a 2^24 transform spent 6 of its 18 products per element on rebuilding
// twiddles (3.25 Gelem/s against 4.15 at 2^22, profiles/r2_summary.md).
__host__ __device__ constexpr int tw_bits_for(int log_n) { return log_n <= 2 * kTwBits ? kTwBits : (log_n + 1) / 2; }

__host__ __device__ constexpr int skew(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }
__host__ __device__ constexpr int data_plane(int log_tile) { return skew(1 << log_tile) + 8; }
__host__ __device__ constexpr int tw_plane(int log_tile) { return skew(1 << (log_tile - 1)) + 8; }
constexpr int smem_bytes(int log_tile) { return 2 * (data_plane(log_tile) + tw_plane(log_tile)) * (int)sizeof(uint4); }

// table[k] = base^k for k < count, base = w_{2^log_size}^(2^shift)
__global__ void k_build_table(Fr* table, int count, int log_size, int shift, bool inverse) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    Fr base = root_of_unity(log_size, inverse);
    for (int i = 0; i < shift; i++) base = base.sqr();
    table[k] = base.pow_u64((uint64_t)k);
}

__global__ void k_build_ninv(Fr* out, int log_n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // n^-1 = (2^-1)^log_n
    Fr two = Fr::one() + Fr::one();
    Fr half = two.inv();
    out[0] = half.pow_u64((uint64_t)log_n);
}

struct PassArgs {
    const Fr* in;
    Fr* out;
    int m;        // log2 of the sub-transform size M
    int log_n;    // log2 of the full transform
    int log_s;    // column pass: log2 stride between sub-transform points
    int log_n1;   // row pass: bits of the most significant digit of the row index
    int log_tile; // elements per CTA (blockDim.x = tile / 8)
    const Fr* core;
    const Fr* tw0;
    const Fr* tw1;
    const Fr* tw2;
    const Fr* scale;  // row pass: optional n^-1
};

__device__ __forceinline__ Fr sm_load(const uint4* lo, const uint4* hi, int slot) {
    Fr r;
    uint4 a = lo[slot], b = hi[slot];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sm_store(uint4* lo, uint4* hi, int slot, const Fr& r) {
    lo[slot] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[slot] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// One register round of R butterfly stages (s .. s+R-1) of an in-place M-point DIT on bit-reversed data.
// A thread owns 8 slots: 2^(3-R) groups of 2^R (R < 3 only in the last round of a transform whose m is not a
// multiple of three).  Stage q pairs j and j | 2^q; its twiddle w_M^e, e = (pos & (2^(s+q) - 1)) << (m-s-q-1),
// depends on j only through its low q bits and its group, so each distinct twiddle is loaded once.
template <int R>
__device__ __forceinline__ void ntt_round(uint4* blo, uint4* bhi, const uint4* tlo, const uint4* thi, int m, int s, int tt) {
    int pos[8];
    Fr x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int vt = (tt << (3 - R)) + (j >> R);
        const int jb = j & ((1 << R) - 1);
        pos[j] = ((vt >> s) << (s + R)) | (jb << s) | (vt & ((1 << s) - 1));
        x[j] = sm_load(blo, bhi, skew(pos[j]));
    }
#pragma unroll
    for (int q = 0; q < R; q++) {
        const int hb = s + q;  // half-size = 2^hb
#pragma unroll
        for (int v = 0; v < (1 << (3 - R)); v++) {
#pragma unroll
            for (int g = 0; g < (1 << q); g++) {
                const int j0 = (v << R) | g;
                const int e = (pos[j0] & ((1 << hb) - 1)) << (m - hb - 1);
                Fr w;
                if (e != 0) w = sm_load(tlo, thi, skew(e));
#pragma unroll
                for (int h = 0; h < (1 << (R - 1 - q)); h++) {
                    const int j = j0 | (h << (q + 1));
                    const int jj = j | (1 << q);
                    Fr t = (e == 0) ? x[jj] : x[jj] * w;
                    x[jj] = x[j] - t;
                    x[j] = x[j] + t;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) sm_store(blo, bhi, skew(pos[j]), x[j]);
}

// In-place M-point DIT butterflies on bit-reversed data for every sub-transform of the CTA.
__device__ __forceinline__ void core_ntt(uint4* lo, uint4* hi, const uint4* tlo, const uint4* thi, int m, int tid) {
    const int M = 1 << m;
    const int mpad = skew(M);
    const int per_sub = M >> 3;          // threads per sub-transform
    const int b = tid / per_sub;
    const int tt = tid - b * per_sub;
    uint4* blo = lo + b * mpad;
    uint4* bhi = hi + b * mpad;
    int s = 0;
    for (; s + 3 <= m; s += 3) {
        ntt_round<3>(blo, bhi, tlo, thi, m, s, tt);
        __syncthreads();
    }
    if (m - s == 2) {
        ntt_round<2>(blo, bhi, tlo, thi, m, s, tt);
        __syncthreads();
    } else if (m - s == 1) {
        ntt_round<1>(blo, bhi, tlo, thi, m, s, tt);
        __syncthreads();
    }
}

__device__ __forceinline__ Fr interpass_twiddle(const PassArgs& a, uint64_t e) {
    const int tb = tw_bits_for(a.log_n);
    Fr t = a.tw0[e & ((1u << tb) - 1)];
    if (a.log_n > tb) t = t * a.tw1[(e >> tb) & ((1u << tb) - 1)];
    if (a.log_n > 2 * tb) t = t * a.tw2[e >> (2 * tb)];
    return t;
}

// w_M^k, k < M/2, from the resident table of w_2048^k into the skewed planes
__device__ __forceinline__ void load_core_twiddles(uint4* tlo, uint4* thi, const Fr* core, int m, int tid, int nthreads) {
    const int half = 1 << (m - 1);
    const uint4* c4 = reinterpret_cast<const uint4*>(core);
    for (int u = tid; u < 2 * half; u += nthreads) {
        const int k = u >> 1, hf = u & 1;
        (hf ? thi : tlo)[skew(k)] = c4[(((size_t)k << (kCoreBits - m)) << 1) + hf];
    }
}

// ---- column pass ----------------------------------------------------------------------
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_ntt_columns(PassArgs a) {
    extern __shared__ uint4 smem[];
    const int lt = a.log_tile, tile = 1 << lt;
    uint4* lo = smem;
    uint4* hi = lo + data_plane(lt);
    uint4* tlo = hi + data_plane(lt);
    uint4* thi = tlo + tw_plane(lt);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int m = a.m, M = 1 << m;
    const int log_b = lt - m;            // B = tile / M columns per CTA
    const int B = 1 << log_b;
    const int mpad = skew(M);
    const uint64_t blocks_per_outer = (uint64_t)1 << (a.log_s - log_b);
    const uint64_t outer = blockIdx.x / blocks_per_outer;
    const uint64_t low0 = (blockIdx.x % blocks_per_outer) << log_b;

    load_core_twiddles(tlo, thi, a.core, m, tid, nthreads);
    const uint4* in4 = reinterpret_cast<const uint4*>(a.in);
    for (int u = tid; u < 2 * tile; u += nthreads) {
        int el = u >> 1, half = u & 1;
        int c = el & (B - 1);
        int n1 = el >> log_b;
        uint64_t gi = ((((outer << m) + (uint64_t)n1) << a.log_s) + low0 + (uint64_t)c);
        uint4 val = in4[gi * 2 + half];
        int slot = c * mpad + skew((int)(__brev((unsigned)n1) >> (32 - m)));
        (half ? hi : lo)[slot] = val;
    }
    __syncthreads();
    core_ntt(lo, hi, tlo, thi, m, tid);
    // twiddle by w_{M*S}^(low*k1) = w_N^(low*k1 * N/(M*S)) and store
    const int tw_shift = a.log_n - m - a.log_s;
    for (int el = tid; el < tile; el += nthreads) {
        int c = el & (B - 1);
        int k1 = el >> log_b;
        Fr v = sm_load(lo, hi, c * mpad + skew(k1));
        uint64_t low = low0 + (uint64_t)c;
        uint64_t e = (low * (uint64_t)k1) << tw_shift;
        if (e != 0) v = v * interpass_twiddle(a, e);
        uint64_t gi = ((((outer << m) + (uint64_t)k1) << a.log_s) + low);
        a.out[gi] = v;
    }
}

// ---- row pass (final; digit-reversal transpose on store) --------------------------------
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_ntt_rows(PassArgs a) {
    extern __shared__ uint4 smem[];
    const int lt = a.log_tile;
    uint4* lo = smem;
    uint4* hi = lo + data_plane(lt);
    uint4* tlo = hi + data_plane(lt);
    uint4* thi = tlo + tw_plane(lt);
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int m = a.m, M = 1 << m;
    const int log_rows = a.log_n - m;
    int log_b = lt - m;
    if (log_b > log_rows) log_b = log_rows;
    const int B = 1 << log_b;
    const int mpad = skew(M);
    const int elems = B << m;
    // rows handled: k1 = k1_0 + b (most significant digit), rest = lower digits
    const int log_rest = log_rows - a.log_n1;
    const uint64_t blocks_per_rest = (uint64_t)1 << (a.log_n1 - log_b);
    const uint64_t rest = blockIdx.x / blocks_per_rest;
    const uint64_t k1_0 = (blockIdx.x % blocks_per_rest) << log_b;

    load_core_twiddles(tlo, thi, a.core, m, tid, nthreads);
    const uint4* in4 = reinterpret_cast<const uint4*>(a.in);
    for (int u = tid; u < 2 * elems; u += nthreads) {
        int el = u >> 1, half = u & 1;
        int n = el & (M - 1);
        int b = el >> m;
        uint64_t row = ((k1_0 + (uint64_t)b) << log_rest) + rest;
        uint4 val = in4[((row << m) + (uint64_t)n) * 2 + half];
        int slot = b * mpad + skew((int)(__brev((unsigned)n) >> (32 - m)));
        (half ? hi : lo)[slot] = val;
    }
    __syncthreads();
    core_ntt(lo, hi, tlo, thi, m, tid);
    Fr scale;
    const bool do_scale = a.scale != nullptr;
    if (do_scale) scale = a.scale[0];
    for (int el = tid; el < elems; el += nthreads) {
        int b = el & (B - 1);
        int k = el >> log_b;
        Fr v = sm_load(lo, hi, b * mpad + skew(k));
        if (do_scale) v = v * scale;
        uint64_t out_base = (k1_0 + (uint64_t)b) + (rest << a.log_n1);
        a.out[out_base + ((uint64_t)k << log_rows)] = v;
    }
}

// N <= 4: direct evaluation by one thread
__global__ void k_ntt_tiny(Fr* data, int log_n, bool inverse) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = 1 << log_n;
    Fr w = root_of_unity(log_n, inverse);
    Fr in[4], out[4];
    for (int i = 0; i < n; i++) in[i] = data[i];
    Fr wi = Fr::one();
    for (int i = 0; i < n; i++) {
        Fr acc = Fr::zero(), p = Fr::one();
        for (int j = 0; j < n; j++) { acc = acc + in[j] * p; p = p * wi; }
        out[i] = acc;
        wi = wi * w;
    }
    Fr scale = Fr::one();
    if (inverse) {
        Fr two = Fr::one() + Fr::one();
        scale = two.inv().pow_u64((uint64_t)log_n);
    }
    for (int i = 0; i < n; i++) data[i] = out[i] * scale;
}

// data[i] *= g^i.  A thread owns kPowPerThread elements strided by the CTA: one pow for its first exponent and the
// CTA stride, then a running product (2 products per element instead of a 64-bit pow per element).
constexpr int kPowPerThread = 16;
__global__ void __launch_bounds__(256) k_scale_by_powers(Fr* data, size_t n, const Fr* g) {
    const size_t base = (size_t)blockIdx.x * (256 * kPowPerThread) + threadIdx.x;
    if (base >= n) return;
    const Fr gen = g[0];
    Fr gi = gen.pow_u64((uint64_t)base);
    const Fr step = gen.pow_u64(256);
    for (int k = 0; k < kPowPerThread; k++) {
        const size_t i = base + (size_t)k * 256;
        if (i >= n) break;
        data[i] = data[i] * gi;
        gi = gi * step;
    }
}

// ---- sharded transform over G = 2^log_g ranks ("four-step, one all-to-all") -------------------
// Rank g owns the interleaved subsequence x[j*G + g] (the same split the MSM shards use).  With n = n2*G + g and
// k = k1*(N/G) + k2:   X[k] = sum_g w_G^(g*k1) * w_N^(g*k2) * Y_g[k2],   Y_g = (N/G)-point transform of rank g's data.
// k_dist_twiddle_pack: multiplies Y_g[k2] by w_N^(g*k2) and lays it out destination-major (rank h receives the
// k2 = h + G*b), ready for one all-to-all; k_dist_combine: the G-point transform across the received blocks.
// Rank h ends up with X[k] for k = h (mod G) at local index (k - h)/G — interleaved again.
__global__ void __launch_bounds__(256) k_dist_twiddle_pack(const Fr* __restrict__ y, Fr* __restrict__ send, PassArgs a, int log_g,
                                                           uint32_t rank) {
    const size_t n2 = (size_t)1 << (a.log_n - log_g);
    const size_t per = n2 >> log_g;                       // elements per destination
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n2) return;
    const size_t h = o / per, b = o - h * per;
    const uint64_t k2 = h + ((uint64_t)b << log_g);
    Fr v = y[k2];
    const uint64_t e = (uint64_t)rank * k2;               // < N
    if (e != 0) v = v * interpass_twiddle(a, e);
    send[o] = v;
}

template <int LOG_G>
__global__ void __launch_bounds__(256) k_dist_combine(const Fr* __restrict__ recv, Fr* __restrict__ out, size_t per,
                                                      const Fr* __restrict__ wg, const Fr* __restrict__ scale) {
    constexpr int G = 1 << LOG_G;
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= per) return;
    Fr x[G], w[G];
#pragma unroll
    for (int g = 0; g < G; g++) { x[g] = recv[(size_t)g * per + b]; w[g] = wg[g]; }
    Fr sc;
    if (scale) sc = scale[0];
#pragma unroll
    for (int k1 = 0; k1 < G; k1++) {
        Fr acc = x[0];
#pragma unroll
        for (int g = 1; g < G; g++) {
            const int e = (g * k1) & (G - 1);
            acc = acc + (e == 0 ? x[g] : x[g] * w[e]);
        }
        if (scale) acc = acc * sc;
        out[(size_t)k1 * per + b] = acc;
    }
}

}  // namespace

NttEngine::~NttEngine() {
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
}

const NttEngine::Tables& NttEngine::tables(int log_n, cudaStream_t stream) {
    auto it = tables_.find(log_n);
    if (it != tables_.end()) return *it->second;
    auto t = std::make_unique<Tables>();
    const int core_n = 1 << (kCoreBits - 1);
    k_build_table<<<ceil_div(core_n, 128), 128, 0, stream>>>(t->core_fwd.as<Fr>(core_n), core_n, kCoreBits, 0, false);
    k_build_table<<<ceil_div(core_n, 128), 128, 0, stream>>>(t->core_inv.as<Fr>(core_n), core_n, kCoreBits, 0, true);
    PM_LAUNCH_CHECK();
    const int tb = tw_bits_for(log_n);
    const int tw_n = 1 << tb;
    for (int level = 0; level < 3; level++) {
        if (level * tb >= log_n && level > 0) break;
        k_build_table<<<ceil_div(tw_n, 128), 128, 0, stream>>>(t->tw_fwd[level].as<Fr>(tw_n), tw_n, log_n, level * tb, false);
        k_build_table<<<ceil_div(tw_n, 128), 128, 0, stream>>>(t->tw_inv[level].as<Fr>(tw_n), tw_n, log_n, level * tb, true);
        PM_LAUNCH_CHECK();
    }
    k_build_ninv<<<1, 32, 0, stream>>>(t->n_inv.as<Fr>(1), log_n);
    PM_LAUNCH_CHECK();
    PM_CUDA(cudaStreamSynchronize(stream));   // built once; later runs may use another stream
    auto& ref = *t;
    tables_[log_n] = std::move(t);
    return ref;
}

void NttEngine::run(Fr* data, int log_n, bool inverse, cudaStream_t stream) {
    if (log_n < 0 || log_n > 32) throw CudaError("ntt: log_n out of range");
    if (log_n == 0) return;
    if (log_n <= 2) {
        k_ntt_tiny<<<1, 32, 0, stream>>>(data, log_n, inverse);
        PM_LAUNCH_CHECK();
        launches++;
        return;
    }
    static bool attr_set = false;
    if (!attr_set) {
        PM_CUDA(cudaFuncSetAttribute(k_ntt_columns<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(kStdTileBits)));
        PM_CUDA(cudaFuncSetAttribute(k_ntt_rows<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(kStdTileBits)));
        PM_CUDA(cudaFuncSetAttribute(k_ntt_columns<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(kMaxTileBits)));
        PM_CUDA(cudaFuncSetAttribute(k_ntt_rows<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(kMaxTileBits)));
        attr_set = true;
    }
    const Tables& t = tables(log_n, stream);
    const size_t n = (size_t)1 << log_n;

    // digit split: 1 pass up to 2^11, 2 passes up to 2^22 (digits <= 11, 2048-element tiles), else 3; most significant first.
    // PM_NTT_BIG_TILE=1: 2^23 / 2^24 as TWO passes of 12-bit digits on 4096-element tiles (220 KB, one CTA of 512 threads
    // per SM).  Measured slower than three passes — 2^23: 3.70 vs 3.79, 2^24: 3.55 vs 3.68 Gelem/s — because a lone CTA
    // cannot overlap its global loads / stores with another CTA's butterflies; kept as a checked alternative.
    static int big_tile = -1;
    if (big_tile < 0) {
        const char* v = getenv("PM_NTT_BIG_TILE");          // tuning hook (see above)
        big_tile = v ? atoi(v) : 0;
    }
    const bool two_big = big_tile && log_n > 2 * kStdTileBits && log_n <= 2 * kMaxTileBits;
    int npass = log_n <= kStdTileBits ? 1 : (log_n <= 2 * kStdTileBits || two_big ? 2 : 3);
    int digits[3] = {0, 0, 0};
    {
        int rem = log_n;
        for (int p = 0; p < npass; p++) {
            digits[p] = (rem + (npass - p) - 1) / (npass - p);
            rem -= digits[p];
        }
    }
    // Tile: the smallest one that holds the largest digit (>= 512 elements): 2^16 launches 128 CTAs instead of 32 on the
    // 148 SMs, and up to 2^20 (digits of 10 bits) four 1024-element CTAs share an SM instead of two of 2048 — one CTA's
    // global loads / stores overlap the others' butterflies (2^20: 3.85 -> 4.24 Gelem/s).  2^21 / 2^22 need 2048.
    // From 2^20 on every pass takes max(its digit, 10) bits: the 10-bit row pass of 2^21 and the 8-bit passes of 2^24 run on
    // 1024-element tiles as well (2^21: 4.23 -> 4.29, 2^24: 3.69 -> 3.80 Gelem/s; 512-element tiles lose the coalescing
    // of the column passes: 2.9 Gelem/s at 2^24).
    int log_tile = two_big ? kMaxTileBits : 10;
    if (log_n < 20) {
        log_tile = digits[0] > kMinTileBits ? digits[0] : kMinTileBits;
        if (log_tile > kStdTileBits) log_tile = kStdTileBits;
    }
    if (const char* v = getenv("PM_NTT_TILE_BITS")) {       // tuning hook
        const int f = atoi(v);
        if (f >= digits[0] && f >= kMinTileBits && f <= kStdTileBits && !two_big) log_tile = f;
    }
    // Per-pass tiles (PM_NTT_COL_TILE / PM_NTT_ROW_TILE: tile bits of the column passes / the row pass, each at least its
    // digit): a smaller tile puts more independent CTAs on an SM, a larger one makes the column pass read longer
    // contiguous chunks (B = tile / M adjacent columns of 32 bytes).
    auto pass_tile = [&](int digit, const char* env, int dflt) {
        int lt = dflt;
        if (const char* v = getenv(env)) {
            const int f = atoi(v);
            if (f >= kMinTileBits && f <= kStdTileBits && !two_big) lt = f;
        }
        if (lt < digit) lt = digit;
        return lt;
    };
    PassArgs a{};
    a.log_n = log_n;
    a.core = (inverse ? t.core_inv : t.core_fwd).get<Fr>();
    a.tw0 = (inverse ? t.tw_inv[0] : t.tw_fwd[0]).get<Fr>();
    a.tw1 = (inverse ? t.tw_inv[1] : t.tw_fwd[1]).get<Fr>();
    a.tw2 = (inverse ? t.tw_inv[2] : t.tw_fwd[2]).get<Fr>();
    if (time_passes) {
        if (!ev_begin) { PM_CUDA(cudaEventCreate(&ev_begin)); PM_CUDA(cudaEventCreate(&ev_end)); }
        PM_CUDA(cudaEventRecord(ev_begin, stream));
    }
    // ping-pong: caller's buffer -> scratch (first column pass), in place in scratch (second), scratch -> caller's
    // buffer (row pass, which transposes and therefore cannot run in place).  A single-pass transform is one CTA that
    // holds the whole array in shared memory between its loads and stores: in place.
    Fr* scratch = npass > 1 ? scratch_.as<Fr>(n) : data;
    int consumed = 0;
    for (int p = 0; p < npass - 1; p++) {
        a.in = p == 0 ? data : scratch;
        a.out = scratch;
        a.m = digits[p];
        a.log_s = log_n - consumed - digits[p];
        a.scale = nullptr;
        int lt = pass_tile(digits[p], "PM_NTT_COL_TILE", log_tile);
        if (lt - a.m > a.log_s) lt = a.m + a.log_s;          // no more adjacent columns than exist
        a.log_tile = lt;
        if (lt > kStdTileBits) k_ntt_columns<512, 1><<<(unsigned)(n >> lt), (1 << lt) / 8, smem_bytes(lt), stream>>>(a);
        else k_ntt_columns<256, 2><<<(unsigned)(n >> lt), (1 << lt) / 8, smem_bytes(lt), stream>>>(a);
        PM_LAUNCH_CHECK();
        launches++;
        consumed += digits[p];
    }
    a.in = scratch;
    a.out = data;
    a.m = digits[npass - 1];
    a.log_s = 0;
    a.log_n1 = npass == 1 ? 0 : digits[0];
    a.scale = inverse ? t.n_inv.get<Fr>() : nullptr;
    {
        const int lt = pass_tile(a.m, "PM_NTT_ROW_TILE", log_tile);
        a.log_tile = lt;
        const int log_rows = log_n - a.m;
        int log_b = lt - a.m;
        if (log_b < 0) throw CudaError("ntt: tile smaller than the row digit");
        if (log_b > log_rows) log_b = log_rows;
        const unsigned ctas = (unsigned)((size_t)1 << (log_rows - log_b));
        int threads = ((1 << log_b) << a.m) / 8;
        if (threads < 1) threads = 1;
        if (lt > kStdTileBits) k_ntt_rows<512, 1><<<ctas, threads, smem_bytes(lt), stream>>>(a);
        else k_ntt_rows<256, 2><<<ctas, threads, smem_bytes(lt), stream>>>(a);
        PM_LAUNCH_CHECK();
        launches++;
    }
    if (time_passes) PM_CUDA(cudaEventRecord(ev_end, stream));
}

void NttEngine::dist_local(Fr* data, Fr* send, int log_n, int log_g, uint32_t rank, bool inverse, cudaStream_t stream) {
    if (log_g < 0 || log_g > 3 || log_n < 2 * log_g || log_n > 32) throw CudaError("sharded ntt: bad geometry");
    if (rank >= (1u << log_g)) throw CudaError("sharded ntt: bad rank");
    run(data, log_n - log_g, inverse, stream);
    const Tables& t = tables(log_n, stream);
    PassArgs a{};
    a.log_n = log_n;
    a.tw0 = (inverse ? t.tw_inv[0] : t.tw_fwd[0]).get<Fr>();
    a.tw1 = (inverse ? t.tw_inv[1] : t.tw_fwd[1]).get<Fr>();
    a.tw2 = (inverse ? t.tw_inv[2] : t.tw_fwd[2]).get<Fr>();
    const size_t n2 = (size_t)1 << (log_n - log_g);
    k_dist_twiddle_pack<<<ceil_div(n2, 256), 256, 0, stream>>>(data, send, a, log_g, rank);
    PM_LAUNCH_CHECK();
    launches++;
}

void NttEngine::dist_combine(const Fr* recv, Fr* out, int log_n, int log_g, bool inverse, cudaStream_t stream) {
    if (log_g < 0 || log_g > 3 || log_n < 2 * log_g || log_n > 32) throw CudaError("sharded ntt: bad geometry");
    const size_t per = (size_t)1 << (log_n - 2 * log_g);
    if (log_g == 0) {
        PM_CUDA(cudaMemcpyAsync(out, recv, per * sizeof(Fr), cudaMemcpyDeviceToDevice, stream));
        return;
    }
    const Tables& t = tables(log_g, stream);      // tw[0][k] = w_G^k, n_inv = G^-1
    const Fr* wg = (inverse ? t.tw_inv[0] : t.tw_fwd[0]).get<Fr>();
    const Fr* scale = inverse ? t.n_inv.get<Fr>() : nullptr;
    const unsigned grid = ceil_div(per, 256);
    if (log_g == 1) k_dist_combine<1><<<grid, 256, 0, stream>>>(recv, out, per, wg, scale);
    else if (log_g == 2) k_dist_combine<2><<<grid, 256, 0, stream>>>(recv, out, per, wg, scale);
    else k_dist_combine<3><<<grid, 256, 0, stream>>>(recv, out, per, wg, scale);
    PM_LAUNCH_CHECK();
    launches++;
}

void launch_scale_by_powers(Fr* data, size_t n, const Fr* g_dev, cudaStream_t stream) {
    k_scale_by_powers<<<ceil_div(n, 256 * kPowPerThread), 256, 0, stream>>>(data, n, g_dev);
    PM_LAUNCH_CHECK();
}

}  // namespace pm
