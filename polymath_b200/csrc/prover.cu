// Prover context and the three proving phases on the device.
//
// Follows `create_proof_with_assignment` (/root/reference/src/prover.rs:66-237).  Control
// crosses the host boundary only where the reference needs a Fiat-Shamir challenge:
//   phase 1  prover.rs:73-123   -> ([a]_1, [c]_1)
//   host     x1 = H(public inputs, a, c)            (common.rs:21-30)
//   phase 2  prover.rs:128-132  -> a(x1)
//   host     pi(x1), c(x1), x2 = H(x1, a(x1), c(x1)) (common.rs:32-75)
//   phase 3  prover.rs:142-229  -> [d]_1
#include "prover.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <cuda_profiler_api.h>

#include "host/g1_host.hpp"
#include "g1_codec.cuh"
#include "msm.cuh"
#include "nccl_dyn.cuh"
#include "ntt.cuh"

namespace pm {

namespace {

enum SmallSlot { S_RA = 0, S_X1 = 5, S_Y1A = 6, S_U_AT_X1 = 7, S_A_AT_X1 = 8, S_X2 = 9, S_EVAL = 10, S_C_AT_X1 = 11, S_COUNT = 16 };

__global__ void k_phase3_consts(Fr* small) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // consts[1] = a(x1) + x2 * c(x1)    (prover.rs:191-209)
    small[S_EVAL] = small[S_A_AT_X1] + small[S_X2] * small[S_C_AT_X1];
}

constexpr size_t kStageWin = (size_t)kMaxMsmSums * sizeof(G1XYZZ);   // one MSM's partial sums
constexpr size_t kStageBytes = 3 * kStageWin + 256;
constexpr size_t kStageStatus = 3 * kStageWin;

// The all-gathered coefficient slices back in natural order.  gathered = [rank][u_loc (per) | wu_loc (per) | u2_loc (2 per)]
// (per = n / G): u[k] = block (k % G), element k / G, and likewise wu, u2.
__global__ void k_interleave_slices(const Fr* __restrict__ gathered, size_t n, uint32_t g, Fr* __restrict__ u, Fr* __restrict__ wu,
                                    Fr* __restrict__ u2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per = n / g;
    if (i >= 4 * n) return;
    if (i < n) u[i] = gathered[(i % g) * 4 * per + i / g];
    else if (i < 2 * n) { const size_t k = i - n; wu[k] = gathered[(k % g) * 4 * per + per + k / g]; }
    else { const size_t k = i - 2 * n; u2[k] = gathered[(k % g) * 4 * per + 2 * per + k / g]; }
}

// Self-describing record of one rank's MSM result inside an in-phase all-gather: the Shape its raw sums are to be decoded
// with (ranks need not agree on it) and the rank's status words.  Followed by kMaxMsmSums XYZZ records.
struct SumsHeader {
    int32_t c, nwin, nsum, count;
    uint8_t shift[32];
    uint32_t status0, status1;       // filled on the device (ST_* bits; "h has a non-zero coefficient")
    uint32_t pad[2];
};
static_assert(sizeof(SumsHeader) == 64, "header layout");
constexpr size_t kBlockBytes = sizeof(SumsHeader) + (size_t)kMaxMsmSums * sizeof(G1XYZZ);

// PM_CUDA_PROFILER=phase1|phase3: cudaProfilerStart/Stop around the device work of that phase of the FIRST proof, so
// that `ncu --profile-from-start off` captures exactly the kernels of one phase (profiles/: the --set full captures).
int profiler_phase() {
    static int which = -1;
    if (which < 0) {
        const char* v = getenv("PM_CUDA_PROFILER");
        which = !v ? 0 : (strcmp(v, "phase1") == 0 ? 1 : strcmp(v, "phase3") == 0 ? 3 : 0);
    }
    return which;
}
bool g_profiler_running = false;
void profiler_begin(int phase_id) {
    static bool done[4] = {false, false, false, false};
    if (profiler_phase() != phase_id || done[phase_id]) return;
    done[phase_id] = true;
    cudaDeviceSynchronize();
    cudaProfilerStart();
    g_profiler_running = true;
}
void profiler_end() {
    if (!g_profiler_running) return;
    cudaDeviceSynchronize();
    cudaProfilerStop();
    g_profiler_running = false;
}

int log2_exact(uint64_t v) {
    int l = 0;
    while (((uint64_t)1 << l) < v) l++;
    if (((uint64_t)1 << l) != v) throw StatusError(PM_ERR_ARG, "domain size is not a power of two");
    return l;
}

}  // namespace

ProverCtx::~ProverCtx() {
    if (nccl_comm && nccl_api().CommDestroy) nccl_api().CommDestroy(nccl_comm);
    if (host_gather) cudaFreeHost(host_gather);
    if (host_stage) cudaFreeHost(host_stage);
    if (host_hdr) cudaFreeHost(host_hdr);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
}

void ProverCtx::upload_matrix(DevMatrix& dst, const uint64_t* row_ptr, const uint32_t* col, const uint8_t* val, bool want_csc) {
    Runtime& rt = runtime();
    const uint64_t ncols = m0 + mw;
    std::vector<uint32_t> rp(nr + 1, 0), cc;
    std::vector<uint8_t> vv;
    if (nr && (!row_ptr || (row_ptr[nr] && (!col || !val)))) throw StatusError(PM_ERR_ARG, "null R1CS matrix arrays");
    const uint64_t nnz_in = nr ? row_ptr[nr] : 0;
    cc.reserve(nnz_in);
    vv.reserve(nnz_in * PM_FR_BYTES);
    std::vector<uint32_t> seen;
    for (uint64_t r = 0; r < nr; r++) {
        rp[r] = (uint32_t)cc.size();
        seen.clear();
        for (uint64_t k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
            uint32_t j = col[k];
            if (j >= ncols) throw StatusError(PM_ERR_ARG, "R1CS column index out of range");
            // m_at (common.rs:100-105) returns the first entry of a column: later duplicates are dead
            if (std::find(seen.begin(), seen.end(), j) != seen.end()) continue;
            seen.push_back(j);
            cc.push_back(j);
            vv.insert(vv.end(), val + k * PM_FR_BYTES, val + (k + 1) * PM_FR_BYTES);
        }
    }
    rp[nr] = (uint32_t)cc.size();
    if (cc.size() >= ((uint64_t)1 << 32)) throw StatusError(PM_ERR_ARG, "R1CS matrix too large");
    dst.nnz = cc.size();
    PM_CUDA(cudaMemcpyAsync(dst.row_ptr.as<uint32_t>(nr + 1), rp.data(), (nr + 1) * 4, cudaMemcpyHostToDevice, rt.stream));
    PM_CUDA(cudaMemcpyAsync(dst.col.as<uint32_t>(dst.nnz + 1), cc.data(), dst.nnz * 4, cudaMemcpyHostToDevice, rt.stream));
    PM_CUDA(cudaMemcpyAsync(dst.val.as<Fr>(dst.nnz + 1), vv.data(), dst.nnz * PM_FR_BYTES, cudaMemcpyHostToDevice, rt.stream));
    if (want_csc) {
        std::vector<uint32_t> cp(ncols + 1, 0), rows(dst.nnz);
        std::vector<uint8_t> cv(dst.nnz * PM_FR_BYTES);
        for (uint32_t j : cc) cp[j + 1]++;
        for (uint64_t j = 0; j < ncols; j++) cp[j + 1] += cp[j];
        std::vector<uint32_t> cur(cp.begin(), cp.end() - 1);
        for (uint64_t r = 0; r < nr; r++)
            for (uint32_t k = rp[r]; k < rp[r + 1]; k++) {
                uint32_t pos = cur[cc[k]]++;
                rows[pos] = (uint32_t)r;
                memcpy(cv.data() + (size_t)pos * PM_FR_BYTES, vv.data() + (size_t)k * PM_FR_BYTES, PM_FR_BYTES);
            }
        PM_CUDA(cudaMemcpyAsync(dst.col_ptr.as<uint32_t>(ncols + 1), cp.data(), (ncols + 1) * 4, cudaMemcpyHostToDevice, rt.stream));
        PM_CUDA(cudaMemcpyAsync(dst.row.as<uint32_t>(dst.nnz + 1), rows.data(), dst.nnz * 4, cudaMemcpyHostToDevice, rt.stream));
        PM_CUDA(cudaMemcpyAsync(dst.cval.as<Fr>(dst.nnz + 1), cv.data(), dst.nnz * PM_FR_BYTES, cudaMemcpyHostToDevice, rt.stream));
    }
    PM_CUDA(cudaStreamSynchronize(rt.stream));  // host vectors die here
}

// Fixed-base tables of one base array: window bits and the largest level count the 31-bit point index allows.
// The level count is then cut down jointly for both arrays by plan_tables().
static ProverCtx::MsmPlan plan_for(uint64_t count, const char* env_name, bool serves_a_side) {
    ProverCtx::MsmPlan p;
    p.stride = count ? count : 1;
    const char* env = getenv(env_name);
    if (!env) env = getenv("PM_MSM_PRECOMP");
    const bool enabled = !(env && atoi(env) == 0);
    const char* envmin = getenv("PM_MSM_PRECOMP_MIN");             // test hook: minimum size that gets tables
    // From 1024 points on: below 2^14 points the table's value is not the additions it saves but its SINGLE bucket set —
    // a table-less 2k-point MSM leaves 26 bucket sets x 10 partial sums to the host (1.07 ms of Horner per phase 1 at
    // n = 2^11, against 1.2 ms of device time): MiMC-322-sized proofs 3.6 -> 2.6 ms (profiles/r2_summary.md).
    const uint64_t min_count = envmin ? (uint64_t)atoll(envmin) : ((uint64_t)1 << 10);
    if (!enabled || count < min_count) return p;                   // small MSMs: plain windows (c chosen per call)
    // All windows of a table share ONE bucket set: the host combines a single group of partial sums instead of
    // 16-22 (7 ms -> 0.1 ms of host time per proof at n = 2^16, profiles/r1_h_summary.md), and the window can be
    // wide.  It must leave >= 2^15 buckets (walk / reduction parallelism) with tens of entries each:
    // measured 2^16: c = 14 / 16 / 18 -> 29.7 / 34.3 / 11.5 ms per proof.
    // Table sweep (profiles/sweep_r1_j_tables.jsonl, all levels): c = 20 wins from 0.4 M to 10.5 M points (10.5 M:
    // 44.7 ms vs 47.8 ms at c = 22 — the wider window saves 5 % of the additions but doubles sort and reduction);
    // the additions only outweigh that above ~32 M points.
    // The c-side table also serves the a-side MSM (a third of its points, running beside it): there the narrower window
    // wins up to ~1 M points (phase 1 at 0.39 M / 0.79 M points: c = 18 -> 5.4 / 8.3 ms, c = 20 -> 6.2 / 8.5 ms).
    const uint64_t c20_from = serves_a_side ? ((uint64_t)1 << 20) : ((uint64_t)1 << 18);
    p.c = count >= ((uint64_t)1 << 25) ? 22 : count >= c20_from ? 20 : count >= ((uint64_t)1 << 16) ? 18 : 16;
    if (env && atoi(env) >= 8) p.c = atoi(env);                // tuning hook: PM_MSM_PRECOMP=<window bits>
    const int nwin = (256 + p.c - 1) / p.c;
    p.levels = 1;
    for (int div = 1; div <= nwin; div++) {
        int lv = (nwin + div - 1) / div;
        if ((uint64_t)lv * p.stride < ((uint64_t)1 << 31)) { p.levels = lv; break; }
    }
    if (p.levels <= 1) { p.levels = 1; p.c = 0; }
    return p;
}

// Next smaller level count that still divides the windows into fewer-level bucket sets (12 -> 6 -> 4 -> 3 -> 2 -> 1).
static int fewer_levels(int c, int levels) {
    const int nwin = (256 + c - 1) / c;
    for (int div = 1; div <= nwin; div++) {
        int lv = (nwin + div - 1) / div;
        if (lv < levels) return lv;
    }
    return 1;
}

// Tables and the batched-affine pair rounds compete for HBM: a table costs levels x count x 96 B, the pair-round
// workspace of ONE bucket set (the MSM engine runs the sets in passes) ~100 B per sorted entry = levels x count x 100 B.
// Without the rounds the accumulation falls back to the XYZZ walk (10 instead of 6 Fq products per addition), which
// costs far more than a few extra bucket sets: shrink the tables until tables + workspaces fit the free memory.
// The c-side and [d]_1 MSMs share one engine (max of their workspaces), the a-side MSM runs beside the c-side on its own.
void ProverCtx::plan_tables() {
    // Every rank of a sharded proof must arrive at the SAME plan: the per-window sums are all-gathered raw and decoded
    // with the local Shape (phase{1,3}_collective).  So the plan is derived from rank-independent inputs only — the
    // largest shard ceil(total / world) (the table stride: ranks with one point less leave the last slot unused), and
    // the free memory quantised down to 4 GiB steps; attach_nccl() then compares the plans of all ranks and refuses
    // to continue on a mismatch (e.g. another process occupying one of the GPUs) instead of hanging in the collective.
    auto shard = [&](uint64_t total) { return (total + (uint64_t)world - 1) / (uint64_t)world; };
    const uint64_t cnt_c = shard(len_c()), cnt_d = world == 1 ? len_d() : d_stride(), cnt_a = shard(len_a());
    plan_c = plan_for(cnt_c, "PM_MSM_PRECOMP_C", true);     // tuning hooks: window bits per array
    plan_d = plan_for(cnt_d, "PM_MSM_PRECOMP_D", false);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
    if (world > 1) free_b &= ~(((size_t)1 << 32) - 1);
    const double budget = 0.85 * (double)free_b;
    auto table = [](uint64_t count, int levels) { return (double)count * levels * sizeof(G1Affine); };
    auto work = [](uint64_t count, int levels) { return (double)count * levels * 100.0 + 1.5e9; };
    auto footprint = [&]() {
        const double wc = work(cnt_c, plan_c.levels), wd = work(cnt_d, plan_d.levels);
        return table(cnt_c, plan_c.levels) + table(cnt_d, plan_d.levels) + (wc > wd ? wc : wd) + work(cnt_a, plan_c.levels);
    };
    while (footprint() > budget && (plan_c.levels > 1 || plan_d.levels > 1)) {
        // shrink the array whose levels currently cost the most
        const double cost_c = plan_c.levels > 1 ? table(cnt_c, plan_c.levels) + work(cnt_c, plan_c.levels) + work(cnt_a, plan_c.levels) : -1;
        const double cost_d = plan_d.levels > 1 ? table(cnt_d, plan_d.levels) + work(cnt_d, plan_d.levels) : -1;
        MsmPlan& p = cost_d >= cost_c ? plan_d : plan_c;
        p.levels = fewer_levels(p.c, p.levels);
        // one level = no table: plain windows, width chosen per call (measured: 16 bits up to 2^26 points,
        // profiles/sweep_r1_h_windows.jsonl — wider windows pay more in sort and reduction than they save in additions)
        if (p.levels <= 1) { p.levels = 1; p.c = 0; }
    }
    if (getenv("PM_MSM_DEBUG"))
        fprintf(stderr, "[plan] free=%.1f GB c-side: %llu pts c=%d levels=%d | d: %llu pts c=%d levels=%d | footprint %.1f GB\n",
                free_b / 1e9, (unsigned long long)cnt_c, plan_c.c, plan_c.levels, (unsigned long long)cnt_d, plan_d.c, plan_d.levels,
                footprint() / 1e9);
}

void ProverCtx::build_tables() {
    Runtime& rt = runtime();
    launch_build_levels(bases_c.get<G1Affine>(), local_count(len_c()), plan_c.levels, plan_c.stride, plan_c.c, rt.stream);
    launch_build_levels(bases_d.get<G1Affine>(), world == 1 ? len_d() : d_count(), plan_d.levels, plan_d.stride, plan_d.c, rt.stream);
    // (the stride is the largest shard; local_count() <= stride on every rank)
}

void ProverCtx::allocate_work() {
    log_n = log2_exact(n);
    if (sigma != n + 3) throw StatusError(PM_ERR_ARG, "sigma != n + 3 (generator.rs:70)");
    if (2 * (m0 + nr) > n) throw StatusError(PM_ERR_ARG, "SAP rows exceed the domain size");
    if (m0 < 1) throw StatusError(PM_ERR_ARG, "instance must contain the leading 1");
    ztail.as<Fr>(cols - m0);
    u.as<Fr>(n); w.as<Fr>(n); wu.as<Fr>(n);
    u2.as<Fr>(2 * n);
    scal_a.as<Fr>(len_a());
    scal_c.as<Fr>(len_c());
    q.as<Fr>(len_d());
    const uint64_t nchunks = (len_d() + kChunk - 1) / kChunk;
    chunk_vals.as<Fr>(2 * (nchunks + 1) + 2 * (nchunks / kChunk + 2) + 8);
    small.as<Fr>(S_COUNT);
    status.as<uint32_t>(4);
    acc.as<G1XYZZ>(3 * kMaxMsmSums);
    PM_CUDA(cudaMallocHost(&host_stage, kStageBytes));
    PM_CUDA(cudaMallocHost(&host_hdr, 4 * sizeof(SumsHeader)));
    PM_CUDA(cudaEventCreate(&ev0));
    PM_CUDA(cudaEventCreate(&ev1));
}

void ProverCtx::set_assignment(const uint8_t* x, const uint8_t* wv) {
    Runtime& rt = runtime();
    if (!x || (mw && !wv)) throw StatusError(PM_ERR_ARG, "null assignment");
    Fr* zt = ztail.get<Fr>();
    PM_CUDA(cudaMemcpyAsync(zt, x, m0 * sizeof(Fr), cudaMemcpyHostToDevice, rt.stream));
    if (mw) PM_CUDA(cudaMemcpyAsync(zt + m0, wv, mw * sizeof(Fr), cudaMemcpyHostToDevice, rt.stream));
    assignment_set = true;
    phase = 0;
}

static void check_status(uint32_t st) {
    if (st & ST_REMAINDER_NONZERO)
        throw StatusError(PM_ERR_UNSATISFIED, "(u^2 - w) is not divisible by Z_H: witness does not satisfy the SAP (prover.rs:108)");
    if (st & (ST_H_ZERO | ST_H_DEGREE)) throw StatusError(PM_ERR_DEGENERATE, "h is zero or deg h > n-2 (prover.rs:107)");
    if (st & ST_OPENING_REMAINDER) throw StatusError(PM_ERR_REMAINDER, "opening numerator does not vanish at x1 (prover.rs:221)");
}

// ---- phase 1 -------------------------------------------------------------------------------------------------------
// compute_a_g1 (prover.rs:330-338) and c_g1 (prover.rs:116-123) over this rank's share of the bases: scalars
// scal[i * stride + offset].  The two MSMs are independent: the a-side runs on the side stream with its own workspace so
// that its sort / reduction tail hides behind the c-side bucket accumulation.
ProverCtx::Phase1Shapes ProverCtx::phase1_msms(const Fr* sa_ptr, const Fr* sc_ptr, size_t stride, size_t offset, G1XYZZ* sums_a,
                                               G1XYZZ* sums_c) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    const G1Affine* bc = bases_c.get<G1Affine>();
    PM_CUDA(cudaEventRecord(rt.ev_fork, s));
    PM_CUDA(cudaStreamWaitEvent(rt.stream2, rt.ev_fork, 0));
    // (round 1 gave the phase-1 MSMs one more pair round than the cost model asked for, because running beside each other
    // hid their inversion latency; since every MSM now overlaps the two halves of its own buckets the model's choice is
    // the best one: 21.6 ms against 22.1 ms with the extra round, profiles/r2_summary.md)
    static int p1_bias = -100;
    if (p1_bias == -100) {
        const char* v = getenv("PM_P1_ROUNDS_BIAS");
        p1_bias = v ? atoi(v) : 0;
    }
    // the a-side bases are a prefix of bases_c: its fixed-base table serves both MSMs
    static int a_tables = -1;
    if (a_tables < 0) {
        const char* v = getenv("PM_A_TABLES");
        a_tables = v ? atoi(v) : 1;
    }
    MsmConfig cfg_a = (a_tables && plan_c.levels > 1) ? cfg_c() : MsmConfig();   // no table: its own (narrower) windows
    cfg_a.rounds_bias = p1_bias;
    MsmConfig cfg_cs = cfg_c();
    cfg_cs.rounds_bias = p1_bias;
    if (world > 1 && cfg_a.c == 0) cfg_a.c = MsmEngine::choose_window(len_a() / world);
    if (world > 1 && cfg_cs.c == 0) cfg_cs.c = MsmEngine::choose_window(len_c() / world);
    Phase1Shapes sh;
    sh.sa = rt.msm2.run(bc, sa_ptr, local_count(len_a()), sums_a, rt.stream2, cfg_a, stride, offset);
    PM_CUDA(cudaEventRecord(rt.ev_join, rt.stream2));
    sh.sc = rt.msm.run(bc, sc_ptr, local_count(len_c()), sums_c, s, cfg_cs, stride, offset);
    PM_CUDA(cudaStreamWaitEvent(s, rt.ev_join, 0));
    return sh;
}

// Replicated polynomial work (one GPU, the callback transport, small domains): prover.rs:73-123.
ProverCtx::Phase1Shapes ProverCtx::phase1_enqueue(const uint8_t* ra, G1XYZZ* sums_a, G1XYZZ* sums_c) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (!assignment_set) throw StatusError(PM_ERR_STATE, "phase 1 needs an assignment");
    if (!ra) throw StatusError(PM_ERR_ARG, "null phase-1 argument");
    phase = 0;
    last_phase1_resident = false;
    Fr* sm = small.get<Fr>();
    uint32_t* st = status.get<uint32_t>();
    profiler_begin(1);
    PM_CUDA(cudaEventRecord(ev0, s));
    PM_CUDA(cudaMemcpyAsync(sm + S_RA, ra, 2 * sizeof(Fr), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemsetAsync(st, 0, 4 * sizeof(uint32_t), s));
    launch_ra_square(sm + S_RA, s);
    SapDims d{(uint32_t)m0, (uint32_t)mw, (uint32_t)nr, n};
    Fr *pu = u.get<Fr>(), *pw = w.get<Fr>(), *pwu = wu.get<Fr>(), *pu2 = u2.get<Fr>();
    launch_sap_evals(d, A.csr(), B.csr(), C.csr(), ztail.get<Fr>(), pu, pw, pwu, s);
    rt.ntt.run(pu, log_n, true, s);    // poly_coeffs, prover.rs:94
    rt.ntt.run(pw, log_n, true, s);    // prover.rs:96
    rt.ntt.run(pwu, log_n, true, s);   // prover.rs:161 (witness_w == w: prover.rs:165 is not recomputed)
    // square_polynomial, prover.rs:315-328
    PM_CUDA(cudaMemcpyAsync(pu2, pu, n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
    PM_CUDA(cudaMemsetAsync(pu2 + n, 0, n * sizeof(Fr), s));
    rt.ntt.run(pu2, log_n + 1, false, s);
    launch_square(pu2, 2 * n, s);
    rt.ntt.run(pu2, log_n + 1, true, s);
    launch_quotient_checks(pu2, pw, n, st, s);   // prover.rs:104-108
    launch_assemble_phase1_scalars(pu, pu2, ztail.get<Fr>(), lay(), sm + S_RA, scal_a.get<Fr>(), scal_c.get<Fr>(), s);
    rt.extra_launches += 9;
    return phase1_msms(scal_a.get<Fr>(), scal_c.get<Fr>(), world, rank, sums_a, sums_c);
}

bool ProverCtx::resident() const {
    int log_g = 0;
    while ((1 << log_g) < world) log_g++;
    return nccl_comm && world > 1 && (1 << log_g) == world && log_g <= 3 && log_n >= resident_min_log && log_n >= 2 * log_g + 3;
}

// In-place transform of this rank's slice (2^lg / G elements, the residue class `rank`) of a 2^lg-point polynomial:
// local (N/G)-point NTT + twiddle / pack, ONE all-to-all over NCCL / NVLink, G-point combine.  Slice in, slice out.
void ProverCtx::sharded_ntt(Fr* loc, int lg, bool inverse, cudaStream_t s) {
    Runtime& rt = runtime();
    int log_g = 0;
    while ((1 << log_g) < world) log_g++;
    const size_t per = ((size_t)1 << lg) >> log_g, blk = per >> log_g;
    Fr* send = ntt_send.as<Fr>(per);
    Fr* recv = ntt_recv.as<Fr>(per);
    rt.ntt.dist_local(loc, send, lg, log_g, (uint32_t)rank, inverse, s);
    NcclApi& api = nccl_api();
    api.check(api.GroupStart(), "ncclGroupStart");
    for (int h = 0; h < world; h++) {
        api.check(api.Send(send + (size_t)h * blk, blk * sizeof(Fr), NcclApi::kUint8, h, nccl_comm, s), "ncclSend");
        api.check(api.Recv(recv + (size_t)h * blk, blk * sizeof(Fr), NcclApi::kUint8, h, nccl_comm, s), "ncclRecv");
    }
    api.check(api.GroupEnd(), "ncclGroupEnd");
    rt.ntt.dist_combine(recv, loc, lg, log_g, inverse, s);      // loc[i] = X[rank + G * i]
}

// Sharded-resident polynomial work: SAP rows of this rank's residue class ("row-range SpMV" over the
// interleaved split), sharded transforms, local checks and scalar assembly — the slices feed the MSM shards in place.
ProverCtx::Phase1Shapes ProverCtx::phase1_enqueue_resident(const uint8_t* ra, G1XYZZ* sums_a, G1XYZZ* sums_c) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (!assignment_set) throw StatusError(PM_ERR_STATE, "phase 1 needs an assignment");
    if (!ra) throw StatusError(PM_ERR_ARG, "null phase-1 argument");
    phase = 0;
    last_phase1_resident = true;
    NcclApi& api = nccl_api();
    Fr* sm = small.get<Fr>();
    uint32_t* st = status.get<uint32_t>();
    profiler_begin(1);
    PM_CUDA(cudaEventRecord(ev0, s));
    PM_CUDA(cudaMemcpyAsync(sm + S_RA, ra, 2 * sizeof(Fr), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemsetAsync(st, 0, 4 * sizeof(uint32_t), s));
    launch_ra_square(sm + S_RA, s);
    const uint32_t G = (uint32_t)world, r = (uint32_t)rank;
    const size_t per = n / G;
    // one buffer: [u_loc (per) | wu_loc (per) | u2_loc (2 per) | w_loc (per)] — the first 4 per elements are all-gathered
    Fr* loc = res_loc.as<Fr>(5 * per);
    Fr *u_loc = loc, *wu_loc = loc + per, *u2_loc = loc + 2 * per, *w_loc = loc + 4 * per;
    SapDims d{(uint32_t)m0, (uint32_t)mw, (uint32_t)nr, n};
    launch_sap_evals_strided(d, A.csr(), B.csr(), C.csr(), ztail.get<Fr>(), r, G, u_loc, w_loc, wu_loc, s);
    const CLayout L = lay();
    const uint64_t cnt_a = local_count(L.len_a), cnt_c = local_count(L.len_c);
    const uint64_t cnt_zt = L.tail > r ? (L.tail - r + G - 1) / G : 0;
    Fr* zt_loc = res_zt.as<Fr>(cnt_zt + 1);
    launch_ztail_strided(d, A.csr(), B.csr(), ztail.get<Fr>(), r, G, cnt_zt, zt_loc, s);
    sharded_ntt(u_loc, log_n, true, s);     // poly_coeffs, prover.rs:94
    sharded_ntt(w_loc, log_n, true, s);     // prover.rs:96
    sharded_ntt(wu_loc, log_n, true, s);    // prover.rs:161
    // the coefficient u_{k-1} next to u_k lives one rank down: ring exchange of the u slices (rank -> rank + 1)
    Fr* u_prev = res_prev.as<Fr>(per);
    api.check(api.GroupStart(), "ncclGroupStart");
    api.check(api.Send(u_loc, per * sizeof(Fr), NcclApi::kUint8, (rank + 1) % world, nccl_comm, s), "ncclSend");
    api.check(api.Recv(u_prev, per * sizeof(Fr), NcclApi::kUint8, (rank + world - 1) % world, nccl_comm, s), "ncclRecv");
    api.check(api.GroupEnd(), "ncclGroupEnd");
    // square_polynomial, prover.rs:315-328: the slice of the zero-padded 2n-array is [u_loc | 0]
    PM_CUDA(cudaMemcpyAsync(u2_loc, u_loc, per * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
    PM_CUDA(cudaMemsetAsync(u2_loc + per, 0, per * sizeof(Fr), s));
    sharded_ntt(u2_loc, log_n + 1, false, s);
    launch_square(u2_loc, 2 * per, s);
    sharded_ntt(u2_loc, log_n + 1, true, s);
    launch_quotient_checks_strided(u2_loc, w_loc, n, r, G, st, s);   // prover.rs:104-108 on the slices
    Fr* sa_loc = res_scal_a.as<Fr>(cnt_a + 1);
    Fr* sc_loc = res_scal_c.as<Fr>(cnt_c + 1);
    launch_assemble_phase1_strided(u_loc, u_prev, u2_loc, zt_loc, L, sm + S_RA, r, G, cnt_a, cnt_c, sa_loc, sc_loc, s);
    // The opening phase works on contiguous coefficient ranges: ONE all-gather brings the coefficient slices of u, wu,
    // u^2 to every rank (natural order restored), issued before the MSMs so that it is long done when they end.
    Fr* gath = ntt_gather.as<Fr>(4 * (size_t)n);
    api.check(api.AllGather(loc, gath, 4 * per * sizeof(Fr), NcclApi::kUint8, nccl_comm, s), "ncclAllGather");
    k_interleave_slices<<<ceil_div(4 * (size_t)n, 256), 256, 0, s>>>(gath, n, G, u.get<Fr>(), wu.get<Fr>(), u2.get<Fr>());
    PM_LAUNCH_CHECK();
    rt.extra_launches += 10;
    return phase1_msms(sa_loc, sc_loc, 1, 0, sums_a, sums_c);
}

void ProverCtx::phase1_partial(const uint8_t* ra, uint8_t* partials_out) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (!partials_out) throw StatusError(PM_ERR_ARG, "null phase-1 argument");
    G1XYZZ* ac = acc.get<G1XYZZ>();
    Phase1Shapes sh = phase1_enqueue(ra, ac, ac + kMaxMsmSums);
    const MsmEngine::Shape &sa = sh.sa, &sc = sh.sc;
    uint32_t* st = status.get<uint32_t>();
    uint8_t* hs = static_cast<uint8_t*>(host_stage);
    PM_CUDA(cudaMemcpyAsync(hs, ac, sa.count() * sizeof(G1XYZZ), cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaMemcpyAsync(hs + kStageWin, ac + kMaxMsmSums, sc.count() * sizeof(G1XYZZ), cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaMemcpyAsync(hs + kStageStatus, st, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaEventRecord(ev1, s));
    PM_CUDA(cudaStreamSynchronize(s));
    profiler_end();
    float ms = 0;
    PM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    phase_ms[0] = ms;
    uint32_t stv;
    memcpy(&stv, hs + kStageStatus, 4);
    check_status(stv);
    host::combine_shifted(hs, sa.nwin, sa.c, sa.nsum, sa.shift).to_wire(partials_out);
    host::combine_shifted(hs + kStageWin, sc.nwin, sc.c, sc.nsum, sc.shift).to_wire(partials_out + sizeof(G1XYZZ));
    phase = 10;   // partial done, waiting for finish
}

// ---- sharded flow with the collective inside the phase (NCCL all-gather on the device, stream-ordered) ----
uint8_t* ProverCtx::gather_stage(size_t bytes) {
    if (bytes > host_gather_bytes) {
        if (host_gather) PM_CUDA(cudaFreeHost(host_gather));
        host_gather = nullptr;
        host_gather_bytes = 0;
        PM_CUDA(cudaMallocHost(&host_gather, bytes));
        host_gather_bytes = bytes;
    }
    return static_cast<uint8_t*>(host_gather);
}

void ProverCtx::attach_nccl(const char* libnccl_path, const uint8_t id[128]) {
    if (nccl_comm) throw StatusError(PM_ERR_STATE, "context already has a communicator");
    if (!id) throw StatusError(PM_ERR_ARG, "null NCCL id");
    NcclApi& api = nccl_api();
    api.load(libnccl_path);
    NcclUniqueId uid;
    memcpy(uid.internal, id, sizeof uid.internal);
    void* comm = nullptr;
    api.check(api.CommInitRank(&comm, world, uid, rank), "ncclCommInitRank");
    nccl_comm = comm;
    if (const char* v = getenv("PM_RESIDENT_MIN_LOG")) resident_min_log = atoi(v);   // tuning / test hook
    // Settings that select the code path of a collective phase (sharded-resident or replicated polynomial work, the
    // geometry of the exchanges) must agree on every rank: compare them now, where a mismatch is an error message and not
    // a hang inside a phase.  (The MSM results themselves travel self-described: see SumsHeader.)
    {
        Runtime& rt = runtime();
        const uint64_t mine[8] = {n, (uint64_t)world, cols, (uint64_t)resident_min_log, m0, mw, nr, 0};
        uint64_t* dev = reinterpret_cast<uint64_t*>(gathered.as<uint8_t>((size_t)(world + 1) * sizeof mine + 256));
        std::vector<uint64_t> all((size_t)world * 8);
        PM_CUDA(cudaMemcpyAsync(dev + (size_t)world * 8, mine, sizeof mine, cudaMemcpyHostToDevice, rt.stream));
        api.check(api.AllGather(dev + (size_t)world * 8, dev, sizeof mine, NcclApi::kUint8, nccl_comm, rt.stream), "ncclAllGather");
        PM_CUDA(cudaMemcpyAsync(all.data(), dev, all.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        for (int r = 0; r < world; r++)
            if (memcmp(all.data() + (size_t)r * 8, mine, sizeof mine) != 0)
                throw StatusError(PM_ERR_STATE, "rank " + std::to_string(rank) + " and rank " + std::to_string(r) +
                                                    " disagree on the circuit dimensions or the sharding settings of the context");
    }
}

namespace {

// host side of an in-phase all-gather: fill this rank's headers, then (after the exchange) decode every rank's block
void fill_header(SumsHeader& h, const MsmEngine::Shape& sh) {
    memset(&h, 0, sizeof h);
    h.c = sh.c; h.nwin = sh.nwin; h.nsum = sh.nsum; h.count = sh.count();
    memcpy(h.shift, sh.shift, sizeof h.shift);
}
host::XyzzH decode_block(const uint8_t* block) {
    SumsHeader h;
    memcpy(&h, block, sizeof h);
    if (h.count < 0 || h.count > kMaxMsmSums || h.nwin * h.nsum != h.count) throw StatusError(PM_ERR_STATE, "corrupt partial-sum block");
    return host::combine_shifted(block + sizeof(SumsHeader), h.nwin, h.c, h.nsum, h.shift);
}
// Sum of one MSM over all ranks' blocks (spaced `pitch` bytes).  Ranks that planned the same shape — the normal case —
// are added element-wise first and decoded ONCE: (G - 1) * count additions instead of G Horner passes with their
// ~20 doublings each (0.25 ms of host time per proof at eight ranks).
host::XyzzH decode_blocks(const uint8_t* first, size_t pitch, int world) {
    SumsHeader h0;
    memcpy(&h0, first, sizeof h0);
    bool uniform = h0.count >= 0 && h0.count <= kMaxMsmSums && h0.nwin * h0.nsum == h0.count;
    for (int r = 1; r < world && uniform; r++) {
        SumsHeader h;
        memcpy(&h, first + (size_t)r * pitch, sizeof h);
        uniform = h.c == h0.c && h.nwin == h0.nwin && h.nsum == h0.nsum && h.count == h0.count && memcmp(h.shift, h0.shift, sizeof h.shift) == 0;
    }
    if (!uniform) {
        host::XyzzH sum = host::XyzzH::inf();
        for (int r = 0; r < world; r++) host::xyzz_add(sum, decode_block(first + (size_t)r * pitch));
        return sum;
    }
    std::vector<uint8_t> merged((size_t)h0.count * sizeof(G1XYZZ));
    for (int i = 0; i < h0.count; i++) {
        host::XyzzH acc = host::XyzzH::inf();
        for (int r = 0; r < world; r++)
            host::xyzz_add(acc, host::XyzzH::from_wire(first + (size_t)r * pitch + sizeof(SumsHeader) + (size_t)i * sizeof(G1XYZZ)));
        acc.to_wire(merged.data() + (size_t)i * sizeof(G1XYZZ));
    }
    return host::combine_shifted(merged.data(), h0.nwin, h0.c, h0.nsum, h0.shift);
}
void block_status(const uint8_t* block, uint32_t& st0, uint32_t& st1) {
    SumsHeader h;
    memcpy(&h, block, sizeof h);
    st0 |= h.status0;
    st1 |= h.status1;
}

}  // namespace

void ProverCtx::phase1_collective(const uint8_t* ra, uint8_t* a_out, uint8_t* c_out) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (!nccl_comm) throw StatusError(PM_ERR_STATE, "no communicator attached (pm_ctx_attach_nccl)");
    if (!a_out || !c_out) throw StatusError(PM_ERR_ARG, "null phase-1 argument");
    // this rank's two blocks: [header | a sums] [header | c sums]
    uint8_t* mine = xchg.as<uint8_t>(2 * kBlockBytes);
    G1XYZZ* sums_a = reinterpret_cast<G1XYZZ*>(mine + sizeof(SumsHeader));
    G1XYZZ* sums_c = reinterpret_cast<G1XYZZ*>(mine + kBlockBytes + sizeof(SumsHeader));
    const bool res = resident();
    Phase1Shapes sh = res ? phase1_enqueue_resident(ra, sums_a, sums_c) : phase1_enqueue(ra, sums_a, sums_c);
    uint32_t* st = status.get<uint32_t>();
    SumsHeader* hh = static_cast<SumsHeader*>(host_hdr);
    fill_header(hh[0], sh.sa);
    fill_header(hh[1], sh.sc);
    // headers first (their status words are then overwritten from the device), sums are already in place
    PM_CUDA(cudaMemcpyAsync(mine, &hh[0], sizeof(SumsHeader), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemcpyAsync(mine + kBlockBytes, &hh[1], sizeof(SumsHeader), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemcpyAsync(mine + offsetof(SumsHeader, status0), st, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    uint8_t* g = gathered.as<uint8_t>((size_t)world * 2 * kBlockBytes + 256);
    NcclApi& api = nccl_api();
    api.check(api.AllGather(mine, g, 2 * kBlockBytes, NcclApi::kUint8, nccl_comm, s), "ncclAllGather");
    uint8_t* hg = gather_stage((size_t)world * 2 * kBlockBytes + 256);
    PM_CUDA(cudaMemcpyAsync(hg, g, (size_t)world * 2 * kBlockBytes, cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaEventRecord(ev1, s));
    PM_CUDA(cudaStreamSynchronize(s));
    profiler_end();
    float ms = 0;
    PM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    phase_ms[0] = ms;
    // status: replicated work -> every rank reports the same bits; resident work -> the union over the slices, and
    // "h == 0" only if no rank saw a non-zero coefficient
    uint32_t st0 = 0, st1 = 0;
    for (int r = 0; r < world; r++) block_status(hg + (size_t)r * 2 * kBlockBytes, st0, st1);
    if (res && st1 == 0) st0 |= ST_H_ZERO;
    check_status(st0);
    host::XyzzH sum_a = decode_blocks(hg, 2 * kBlockBytes, world), sum_c = decode_blocks(hg + kBlockBytes, 2 * kBlockBytes, world);
    host::xyzz_to_affine_wire(sum_a, a_out);
    host::xyzz_to_affine_wire(sum_c, c_out);
    phase = 1;
}

void ProverCtx::phase1_finish(const uint8_t* gathered, int count, uint8_t* a_out, uint8_t* c_out) {
    if (phase != 10) throw StatusError(PM_ERR_STATE, "phase-1 finish must follow phase-1 partial");
    if (!gathered || count < 1 || !a_out || !c_out) throw StatusError(PM_ERR_ARG, "bad phase-1 finish argument");
    // gathered = count records of [a-partial (192 B) | c-partial (192 B)]
    host::xyzz_to_affine_wire(host::sum_partials(gathered, count, 2 * sizeof(G1XYZZ)), a_out);
    host::xyzz_to_affine_wire(host::sum_partials(gathered + sizeof(G1XYZZ), count, 2 * sizeof(G1XYZZ)), c_out);
    phase = 1;
}

void ProverCtx::phase2(const uint8_t* x1, const uint8_t* y1_alpha, uint8_t* a_at_x1_out) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (phase != 1) throw StatusError(PM_ERR_STATE, "phase 2 must follow phase 1");
    if (!x1 || !y1_alpha || !a_at_x1_out) throw StatusError(PM_ERR_ARG, "null phase-2 argument");
    Fr* sm = small.get<Fr>();
    PM_CUDA(cudaEventRecord(ev0, s));
    PM_CUDA(cudaMemcpyAsync(sm + S_X1, x1, sizeof(Fr), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemcpyAsync(sm + S_Y1A, y1_alpha, sizeof(Fr), cudaMemcpyHostToDevice, s));
    const uint64_t nchunks = (n + kChunk - 1) / kChunk;
    launch_chunk_eval_plain(u.get<Fr>(), n, sm + S_X1, chunk_vals.get<Fr>(), s);      // u_poly.evaluate(&x1), prover.rs:132
    launch_combine_chunks(chunk_vals.get<Fr>(), nchunks, sm + S_X1, sm + S_U_AT_X1, s);
    launch_a_at_x1(sm + S_U_AT_X1, sm + S_RA, sm + S_X1, sm + S_A_AT_X1, s);
    rt.extra_launches += 3;
    uint8_t* hs = static_cast<uint8_t*>(host_stage);
    PM_CUDA(cudaMemcpyAsync(hs, sm + S_A_AT_X1, sizeof(Fr), cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaEventRecord(ev1, s));
    PM_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    PM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    phase_ms[1] = ms;
    memcpy(a_at_x1_out, hs, PM_FR_BYTES);
    phase = 2;
}

NumeratorSrc ProverCtx::numerator_src() const {
    NumeratorSrc src;
    const Fr* sm = small.get<Fr>();
    src.u = u.get<Fr>();
    src.wu = wu.get<Fr>();
    src.u2 = u2.get<Fr>();
    src.ra_ext = sm + S_RA;
    src.consts = sm + S_X2;
    src.n = n;
    src.sigma = sigma;
    src.len = len_d();
    return src;
}

// ---- self-test of the sharded-resident kernels with VIRTUAL ranks on one GPU ------------------------------------------
// For a world == 1 context that has just finished a proof: replays the polynomial work of phase 1 and the division of
// phase 3 the way G ranks would — strided SAP rows, sharded transforms (the all-to-all emulated by copies between the
// virtual ranks' buffers), ring exchange, local checks and scalar assembly, chunk-range division with the carry exchange —
// and counts the elements that differ from what the replicated path left in the context (which the tests compare with
// the oracle).  Covers everything of the multi-GPU flow except the NCCL calls themselves.
namespace {
__global__ void k_count_mismatch_strided(const Fr* __restrict__ full, const Fr* __restrict__ loc, uint64_t count, uint32_t g, uint32_t r,
                                         unsigned long long* __restrict__ bad) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    if (full[(uint64_t)r + t * g] != loc[t]) atomicAdd(bad, 1ull);
}
}  // namespace

uint64_t ProverCtx::selftest_resident(int G) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    int log_g = 0;
    while ((1 << log_g) < G) log_g++;
    if (world != 1 || (1 << log_g) != G || G < 2 || G > 8 || log_n < 2 * log_g + 3)
        throw StatusError(PM_ERR_ARG, "self-test needs an unsharded context, 2 <= G <= 8 a power of two, n >= 8 G^2");
    if (phase != 0 || !assignment_set) throw StatusError(PM_ERR_STATE, "self-test must follow a complete proof");
    const size_t per = n / G;
    const CLayout L = lay();
    Fr* sm = small.get<Fr>();
    DevBuf b_loc, b_zt, b_send, b_recv, b_prev, b_sa, b_sc, b_bad, b_st, b_q2, b_rv, b_rlo, b_work;
    Fr* loc = b_loc.as<Fr>(5 * per * G);
    const uint64_t zt_cap = (L.tail + G - 1) / G + 1, sc_cap = (L.len_c + G - 1) / G + 1;
    Fr* zt = b_zt.as<Fr>(zt_cap * G);
    Fr* send = b_send.as<Fr>(2 * per * G);
    Fr* recv = b_recv.as<Fr>(2 * per * G);
    Fr* prev = b_prev.as<Fr>(per * G);
    Fr* sa = b_sa.as<Fr>(sc_cap * G);
    Fr* sc = b_sc.as<Fr>(sc_cap * G);
    unsigned long long* bad = b_bad.as<unsigned long long>(1);
    uint32_t* st = b_st.as<uint32_t>(4 * G);
    PM_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), s));
    PM_CUDA(cudaMemsetAsync(st, 0, 4 * G * sizeof(uint32_t), s));
    SapDims d{(uint32_t)m0, (uint32_t)mw, (uint32_t)nr, n};
    auto rloc = [&](int r) { return loc + (size_t)r * 5 * per; };
    for (int r = 0; r < G; r++) {
        Fr* l = rloc(r);
        launch_sap_evals_strided(d, A.csr(), B.csr(), C.csr(), ztail.get<Fr>(), r, G, l, l + 4 * per, l + per, s);
        const uint64_t cnt_zt = L.tail > (uint64_t)r ? (L.tail - r + G - 1) / G : 0;
        launch_ztail_strided(d, A.csr(), B.csr(), ztail.get<Fr>(), r, G, cnt_zt, zt + (size_t)r * zt_cap, s);
    }
    // sharded transform of the slices at offset `off` of every virtual rank's buffer, 2^lg points in total
    auto vntt = [&](size_t off, int lg, bool inverse) {
        const size_t p2 = ((size_t)1 << lg) >> log_g, blk = p2 >> log_g;
        for (int r = 0; r < G; r++) rt.ntt.dist_local(rloc(r) + off, send + (size_t)r * 2 * per, lg, log_g, (uint32_t)r, inverse, s);
        for (int h = 0; h < G; h++)
            for (int r = 0; r < G; r++)
                PM_CUDA(cudaMemcpyAsync(recv + (size_t)h * 2 * per + (size_t)r * blk, send + (size_t)r * 2 * per + (size_t)h * blk,
                                        blk * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        for (int h = 0; h < G; h++) rt.ntt.dist_combine(recv + (size_t)h * 2 * per, rloc(h) + off, lg, log_g, inverse, s);
    };
    vntt(0, log_n, true);
    vntt(4 * per, log_n, true);
    vntt(per, log_n, true);
    for (int r = 0; r < G; r++) {
        PM_CUDA(cudaMemcpyAsync(prev + (size_t)r * per, rloc((r + G - 1) % G), per * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        PM_CUDA(cudaMemcpyAsync(rloc(r) + 2 * per, rloc(r), per * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        PM_CUDA(cudaMemsetAsync(rloc(r) + 3 * per, 0, per * sizeof(Fr), s));
    }
    vntt(2 * per, log_n + 1, false);
    for (int r = 0; r < G; r++) launch_square(rloc(r) + 2 * per, 2 * per, s);
    vntt(2 * per, log_n + 1, true);
    for (int r = 0; r < G; r++) {
        Fr* l = rloc(r);
        launch_quotient_checks_strided(l + 2 * per, l + 4 * per, n, r, G, st + 4 * r, s);
        const uint64_t cnt_a = L.len_a > (uint64_t)r ? (L.len_a - r + G - 1) / G : 0, cnt_c = L.len_c > (uint64_t)r ? (L.len_c - r + G - 1) / G : 0;
        launch_assemble_phase1_strided(l, prev + (size_t)r * per, l + 2 * per, zt + (size_t)r * zt_cap, L, sm + S_RA, r, G, cnt_a, cnt_c,
                                       sa + (size_t)r * sc_cap, sc + (size_t)r * sc_cap, s);
        auto cmp = [&](const Fr* full, const Fr* part, uint64_t count) {
            if (count) k_count_mismatch_strided<<<ceil_div(count, 256), 256, 0, s>>>(full, part, count, G, r, bad);
        };
        cmp(scal_a.get<Fr>(), sa + (size_t)r * sc_cap, cnt_a);
        cmp(scal_c.get<Fr>(), sc + (size_t)r * sc_cap, cnt_c);
        cmp(u.get<Fr>(), l, per);
        cmp(wu.get<Fr>(), l + per, per);
        cmp(u2.get<Fr>(), l + 2 * per, 2 * per);
        cmp(w.get<Fr>(), l + 4 * per, per);
        PM_LAUNCH_CHECK();
    }
    // phase 3: the division by chunk ranges with the carry exchange, against the q of the replicated division
    NumeratorSrc src = numerator_src();
    const uint64_t c1 = d_chunks(), per_c = (c1 + G - 1) / G;
    auto clo = [&](int r) { uint64_t v = (uint64_t)r * per_c; return v < c1 ? v : c1; };
    Fr* q2 = b_q2.as<Fr>(len_d());
    PM_CUDA(cudaMemsetAsync(q2, 0xff, (len_d() - 1) * sizeof(Fr), s));
    Fr* rv = b_rv.as<Fr>((size_t)G + 4);
    uint64_t* rlo = b_rlo.as<uint64_t>((size_t)G + 1);
    std::vector<uint64_t> lo((size_t)G + 1);
    for (int r = 0; r <= G; r++) lo[r] = clo(r);
    PM_CUDA(cudaMemcpyAsync(rlo, lo.data(), lo.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    const uint64_t nchunks = (len_d() + kChunk - 1) / kChunk;
    Fr* work = b_work.as<Fr>(2 * (nchunks + 1) + 2 * (nchunks / kChunk + 2) + 8);
    for (int r = 0; r < G; r++) launch_numerator_range_eval(src, sm + S_X1, clo(r), clo(r + 1) - clo(r), work, rv + r, s);
    for (int r = 0; r < G; r++) {
        // the virtual ranks share one `work` buffer: restore rank r's chunk values (the second-level values of a long
        // range are indexed from 0 and were overwritten by the other ranks' evaluations)
        launch_numerator_range_eval(src, sm + S_X1, clo(r), clo(r + 1) - clo(r), work, rv + G + 1, s);
        launch_numerator_range_carry(rv, rlo, (uint32_t)r, (uint32_t)G, sm + S_X1, rv + G, st + 4 * r, s);
        launch_numerator_range_divide(src, sm + S_X1, clo(r), clo(r + 1) - clo(r), rv + G, q2, work, s);
    }
    k_count_mismatch_strided<<<ceil_div(len_d() - 1, 256), 256, 0, s>>>(q.get<Fr>(), q2, len_d() - 1, 1, 0, bad);
    PM_LAUNCH_CHECK();
    unsigned long long hbad = 0;
    std::vector<uint32_t> hst(4 * (size_t)G);
    PM_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof hbad, cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaMemcpyAsync(hst.data(), st, hst.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaStreamSynchronize(s));
    uint32_t any_h = 0;
    for (int r = 0; r < G; r++) {
        if (hst[4 * r] != 0) hbad += 1000000;        // a satisfied witness must raise no status bit on any virtual rank
        any_h |= hst[4 * r + 1];
    }
    if (!any_h) hbad += 1000000;
    return hbad;
}

// ---- phase 3 -------------------------------------------------------------------------------------------------------
// range_division (collective flow): every rank divides only its own chunk range of the numerator — chunk values of the
// range, ONE all-gather of the G range values ("carry exchange"), the range's quotient coefficients — and feeds them
// to its contiguous shard of x_powers_y_gamma_z.  Otherwise the whole division runs here (one GPU / callback transport).
MsmEngine::Shape ProverCtx::phase3_enqueue(const uint8_t* x2, const uint8_t* c_at_x1, G1XYZZ* sums_d, bool range_division) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (phase != 2) throw StatusError(PM_ERR_STATE, "phase 3 must follow phase 2");
    if (!x2 || !c_at_x1) throw StatusError(PM_ERR_ARG, "null phase-3 argument");
    Fr* sm = small.get<Fr>();
    uint32_t* st = status.get<uint32_t>();
    profiler_begin(3);
    PM_CUDA(cudaEventRecord(ev0, s));
    PM_CUDA(cudaMemcpyAsync(sm + S_X2, x2, sizeof(Fr), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemcpyAsync(sm + S_C_AT_X1, c_at_x1, sizeof(Fr), cudaMemcpyHostToDevice, s));
    k_phase3_consts<<<1, 32, 0, s>>>(sm);
    PM_LAUNCH_CHECK();
    NumeratorSrc src = numerator_src();
    Fr* qp = q.get<Fr>();
    if (range_division && world > 1) {
        NcclApi& api = nccl_api();
        const uint64_t c_lo = chunk_lo(rank), cnt = chunk_lo(rank + 1) - c_lo;
        Fr* rv = range_vals.as<Fr>((size_t)world + 4);          // [0, world): gathered; [world]: mine; [world + 1]: carry in
        uint64_t* rlo = range_lo_dev.as<uint64_t>((size_t)world + 1);
        if (!range_lo_uploaded) {
            std::vector<uint64_t> lo((size_t)world + 1);
            for (int r = 0; r <= world; r++) lo[r] = chunk_lo(r);
            PM_CUDA(cudaMemcpyAsync(rlo, lo.data(), lo.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
            PM_CUDA(cudaStreamSynchronize(s));
            range_lo_uploaded = true;
        }
        rt.extra_launches += 1 + launch_numerator_range_eval(src, sm + S_X1, c_lo, cnt, chunk_vals.get<Fr>(), rv + world, s);
        api.check(api.AllGather(rv + world, rv, sizeof(Fr), NcclApi::kUint8, nccl_comm, s), "ncclAllGather");
        launch_numerator_range_carry(rv, rlo, (uint32_t)rank, (uint32_t)world, sm + S_X1, rv + world + 1, st, s);
        rt.extra_launches += 1 + launch_numerator_range_divide(src, sm + S_X1, c_lo, cnt, rv + world + 1, qp, chunk_vals.get<Fr>(), s);
    } else {
        rt.extra_launches += 1 + launch_divide_numerator(src, sm + S_X1, qp, chunk_vals.get<Fr>(), st, s);   // prover.rs:211-225
    }
    MsmConfig cfg = cfg_d();
    const uint64_t lo = world == 1 ? 0 : d_lo(rank), cnt_d = world == 1 ? src.len - 1 : d_count();
    if (world > 1 && cfg.c == 0) cfg.c = MsmEngine::choose_window((src.len - 1) / world);
    return rt.msm.run(bases_d.get<G1Affine>(), qp + lo, cnt_d, sums_d, s, cfg, 1, 0);  // prover.rs:229
}

void ProverCtx::phase3_partial(const uint8_t* x2, const uint8_t* c_at_x1, uint8_t* partial_out) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (!partial_out) throw StatusError(PM_ERR_ARG, "null phase-3 argument");
    G1XYZZ* ac = acc.get<G1XYZZ>() + 2 * kMaxMsmSums;
    MsmEngine::Shape sd = phase3_enqueue(x2, c_at_x1, ac, false);
    uint32_t* st = status.get<uint32_t>();
    uint8_t* hs = static_cast<uint8_t*>(host_stage);
    PM_CUDA(cudaMemcpyAsync(hs, ac, sd.count() * sizeof(G1XYZZ), cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaMemcpyAsync(hs + kStageStatus, st, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaEventRecord(ev1, s));
    PM_CUDA(cudaStreamSynchronize(s));
    profiler_end();
    float ms = 0;
    PM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    phase_ms[2] = ms;
    uint32_t stv;
    memcpy(&stv, hs + kStageStatus, 4);
    check_status(stv);
    host::combine_shifted(hs, sd.nwin, sd.c, sd.nsum, sd.shift).to_wire(partial_out);
    phase = 30;
}

void ProverCtx::phase3_collective(const uint8_t* x2, const uint8_t* c_at_x1, uint8_t* d_out) {
    Runtime& rt = runtime();
    cudaStream_t s = rt.stream;
    if (!nccl_comm) throw StatusError(PM_ERR_STATE, "no communicator attached (pm_ctx_attach_nccl)");
    if (!d_out) throw StatusError(PM_ERR_ARG, "null phase-3 argument");
    uint8_t* mine = xchg.as<uint8_t>(2 * kBlockBytes);
    G1XYZZ* sums_d = reinterpret_cast<G1XYZZ*>(mine + sizeof(SumsHeader));
    MsmEngine::Shape sd = phase3_enqueue(x2, c_at_x1, sums_d, true);
    uint32_t* st = status.get<uint32_t>();
    SumsHeader* hh = static_cast<SumsHeader*>(host_hdr);
    fill_header(hh[2], sd);
    PM_CUDA(cudaMemcpyAsync(mine, &hh[2], sizeof(SumsHeader), cudaMemcpyHostToDevice, s));
    PM_CUDA(cudaMemcpyAsync(mine + offsetof(SumsHeader, status0), st, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    uint8_t* g = gathered.as<uint8_t>((size_t)world * 2 * kBlockBytes + 256);
    NcclApi& api = nccl_api();
    api.check(api.AllGather(mine, g, kBlockBytes, NcclApi::kUint8, nccl_comm, s), "ncclAllGather");
    uint8_t* hg = gather_stage((size_t)world * 2 * kBlockBytes + 256);
    PM_CUDA(cudaMemcpyAsync(hg, g, (size_t)world * kBlockBytes, cudaMemcpyDeviceToHost, s));
    PM_CUDA(cudaEventRecord(ev1, s));
    PM_CUDA(cudaStreamSynchronize(s));
    profiler_end();
    float ms = 0;
    PM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    phase_ms[2] = ms;
    uint32_t st0 = 0, st1 = 0;
    for (int r = 0; r < world; r++) block_status(hg + (size_t)r * kBlockBytes, st0, st1);
    check_status(st0);
    host::XyzzH sum_d = decode_blocks(hg, kBlockBytes, world);
    host::xyzz_to_affine_wire(sum_d, d_out);
    phase = 0;
}

void ProverCtx::phase3_finish(const uint8_t* gathered, int count, uint8_t* d_out) {
    if (phase != 30) throw StatusError(PM_ERR_STATE, "phase-3 finish must follow phase-3 partial");
    if (!gathered || count < 1 || !d_out) throw StatusError(PM_ERR_ARG, "bad phase-3 finish argument");
    host::xyzz_to_affine_wire(host::sum_partials(gathered, count), d_out);
    phase = 0;
}

}  // namespace pm

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
using namespace pm;

struct pm_ctx {
    ProverCtx impl;
};

namespace {

void init_dims(ProverCtx& c, const pm_r1cs_view& r) {
    c.m0 = r.num_instance_variables;
    c.mw = r.num_r1cs_witness_variables;
    c.nr = r.num_r1cs_constraints;
    if (c.m0 + c.mw >= ((uint64_t)1 << 32) || c.nr >= ((uint64_t)1 << 31)) throw StatusError(PM_ERR_ARG, "R1CS too large");
    c.cols = 3 * c.m0 + c.mw + c.nr;   // SAPMatrices::size, common.rs:131-135
}

// Upload the points of one key vector that belong to this rank.  `global_off` is the vector's offset
// inside the concatenated base array; the rank keeps global indices g with g % world == rank, stored
// compactly at local index g / world.
void upload_points(const ProverCtx& c, G1Affine* dst_base, uint64_t global_off, const uint8_t* src, size_t stride,
                   uint64_t len, uint64_t expect, const char* name, bool contiguous = false) {
    if (len != expect) throw StatusError(PM_ERR_ARG, std::string("unexpected length of ") + name);
    if (len && !src) throw StatusError(PM_ERR_ARG, std::string("null ") + name);
    Runtime& rt = runtime();
    if (contiguous && c.world > 1) {
        // the d-side vector: this rank's contiguous range [d_lo, d_hi) as an unsharded upload of that sub-vector
        ProverCtx one;
        one.world = 1;
        one.rank = 0;
        one.validate_key = c.validate_key;
        const uint64_t lo = c.d_lo(c.rank), cnt = c.d_count();
        if (cnt) upload_points(one, dst_base, 0, src + lo * stride, stride, cnt, cnt, name, false);
        return;
    }
    if (c.world == 1 && stride == PM_G1_BYTES) {
        PM_CUDA(cudaMemcpyAsync(dst_base + global_off, src, len * sizeof(G1Affine), cudaMemcpyHostToDevice, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        return;
    }
    // first global index >= global_off owned by this rank
    uint64_t g0 = global_off + ((uint64_t)c.rank + c.world - global_off % c.world) % c.world;
    if (stride == PM_G1_COMPRESSED_BYTES) {
        // compressed key vector (`serialize_compressed` bytes without the length prefix): decode on the device
        std::vector<uint8_t> packed;
        const uint8_t* first = src + (g0 - global_off) * 48;
        uint64_t cnt = g0 < global_off + len ? (global_off + len - g0 + c.world - 1) / c.world : 0;
        if (c.world > 1) {
            packed.reserve(cnt * 48);
            for (uint64_t g = g0; g < global_off + len; g += c.world) packed.insert(packed.end(), src + (g - global_off) * 48, src + (g - global_off) * 48 + 48);
            first = packed.data();
        }
        if (cnt) {
            DevBuf din, dbad;
            uint8_t* pi = din.as<uint8_t>(cnt * 48);
            unsigned long long* bad = dbad.as<unsigned long long>(1);
            PM_CUDA(cudaMemcpyAsync(pi, first, cnt * 48, cudaMemcpyHostToDevice, rt.stream));
            PM_CUDA(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), rt.stream));
            launch_g1_decompress(pi, cnt, c.validate_key, dst_base + g0 / c.world, bad, rt.stream);
            rt.extra_launches++;
            unsigned long long hbad = 0;
            PM_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof hbad, cudaMemcpyDeviceToHost, rt.stream));
            PM_CUDA(cudaStreamSynchronize(rt.stream));
            if (hbad != ~0ull)
                throw StatusError(PM_ERR_ARG, std::string("invalid compressed point in ") + name + ": " +
                                                  g1_decode_status_name((unsigned)(hbad & 7)));
        }
        return;
    }
    std::vector<uint8_t> packed;
    for (uint64_t g = g0; g < global_off + len; g += c.world) {
        const uint8_t* p = src + (g - global_off) * stride;
        if (stride >= 104 && p[96] != 0) packed.insert(packed.end(), PM_G1_BYTES, 0);
        else packed.insert(packed.end(), p, p + PM_G1_BYTES);
    }
    if (!packed.empty()) {
        PM_CUDA(cudaMemcpyAsync(dst_base + g0 / c.world, packed.data(), packed.size(), cudaMemcpyHostToDevice, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
    }
}

struct KeySlice { uint64_t off, len; bool in_d; };
KeySlice key_slice(const ProverCtx& c, int which) {
    const uint64_t n = c.n;
    const CLayout L = c.lay();
    switch (which) {
        case 0: return {0, n + 1, false};
        case 1: return {L.off_ya, 3, false};
        case 2: return {L.off_zh, n - 1, false};
        case 3: return {L.off_yg, 2, false};
        case 4: return {0, c.len_d(), true};
        case 5: return {L.off_lcs, c.cols - c.m0, false};
        default: throw StatusError(PM_ERR_ARG, "bad key vector index");
    }
}

}  // namespace

extern "C" {

int pm_ctx_create(const pm_pk_view* pk, pm_ctx** out) { return pm_ctx_create_sharded(pk, 0, 1, out); }

static int ctx_create_impl(const pm_pk_view* pk, int rank, int world, bool validate_key, pm_ctx** out);
int pm_ctx_create_sharded(const pm_pk_view* pk, int rank, int world, pm_ctx** out) { return ctx_create_impl(pk, rank, world, true, out); }
int pm_ctx_create_unchecked(const pm_pk_view* pk, int rank, int world, pm_ctx** out) { return ctx_create_impl(pk, rank, world, false, out); }

static int ctx_create_impl(const pm_pk_view* pk, int rank, int world, bool validate_key, pm_ctx** out) {
    return guarded([&] {
        if (!pk || !out) throw StatusError(PM_ERR_ARG, "null argument");
        if (pk->point_stride < PM_G1_BYTES && pk->point_stride != PM_G1_COMPRESSED_BYTES)
            throw StatusError(PM_ERR_ARG, "point stride must be >= 96, or 48 for compressed points");
        if (world < 1 || rank < 0 || rank >= world) throw StatusError(PM_ERR_ARG, "bad rank/world");
        runtime();
        std::unique_ptr<pm_ctx> h(new pm_ctx());
        ProverCtx& c = h->impl;
        c.rank = rank;
        c.world = world;
        c.validate_key = validate_key;
        init_dims(c, pk->r1cs);
        c.n = pk->n;
        c.sigma = pk->sigma;
        c.allocate_work();
        c.upload_matrix(c.A, pk->r1cs.a_row_ptr, pk->r1cs.a_col, pk->r1cs.a_val, false);
        c.upload_matrix(c.B, pk->r1cs.b_row_ptr, pk->r1cs.b_col, pk->r1cs.b_val, false);
        c.upload_matrix(c.C, pk->r1cs.c_row_ptr, pk->r1cs.c_col, pk->r1cs.c_val, false);
        c.plan_tables();
        G1Affine* bc = c.bases_c.as<G1Affine>(c.plan_c.levels * c.plan_c.stride);
        G1Affine* bd = c.bases_d.as<G1Affine>(c.plan_d.levels * c.plan_d.stride);
        // the gaps of the c-side layout (CLayout) are points at infinity
        PM_CUDA(cudaMemsetAsync(bc, 0, c.plan_c.stride * sizeof(G1Affine), runtime().stream));
        const size_t st = pk->point_stride;
        const uint64_t n = c.n;
        const CLayout L = c.lay();
        upload_points(c, bc, 0, pk->x_powers_g1, st, pk->x_powers_g1_len, n + 1, "x_powers_g1");
        upload_points(c, bc, L.off_ya, pk->x_powers_y_alpha_g1, st, pk->x_powers_y_alpha_g1_len, 3, "x_powers_y_alpha_g1");
        upload_points(c, bc, L.off_yg, pk->x_powers_y_gamma_g1, st, pk->x_powers_y_gamma_g1_len, 2, "x_powers_y_gamma_g1");
        upload_points(c, bc, L.off_zh, pk->x_powers_zh_by_y_alpha_g1, st, pk->x_powers_zh_by_y_alpha_g1_len, n - 1, "x_powers_zh_by_y_alpha_g1");
        upload_points(c, bc, L.off_lcs, pk->uj_wj_lcs_by_y_alpha_g1, st, pk->uj_wj_lcs_by_y_alpha_g1_len, c.cols - c.m0, "uj_wj_lcs_by_y_alpha_g1");
        upload_points(c, bd, 0, pk->x_powers_y_gamma_z_g1, st, pk->x_powers_y_gamma_z_g1_len, c.len_d(), "x_powers_y_gamma_z_g1", true);
        c.build_tables();
        *out = h.release();
    });
}

void pm_ctx_destroy(pm_ctx* ctx) {
    if (!ctx) return;
    cudaDeviceSynchronize();
    delete ctx;
}

int pm_setup(const pm_r1cs_view* r1cs, const uint8_t x[PM_FR_BYTES], const uint8_t z[PM_FR_BYTES], pm_ctx** out,
             uint8_t x_g2[192], uint8_t z_g2[192]) {
    return pm_setup_sharded(r1cs, x, z, 0, 1, out, x_g2, z_g2);
}

int pm_host_sum_partials(const uint8_t* parts, int count, size_t stride, uint8_t out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!parts || count < 0 || stride < sizeof(G1XYZZ) || !out) throw StatusError(PM_ERR_ARG, "bad argument");
        host::xyzz_to_affine_wire(host::sum_partials(parts, count, stride), out);
    });
}

int pm_setup_sharded(const pm_r1cs_view* r1cs, const uint8_t x[PM_FR_BYTES], const uint8_t z[PM_FR_BYTES], int rank, int world,
                     pm_ctx** out, uint8_t x_g2[192], uint8_t z_g2[192]) {
    return guarded([&] {
        if (!r1cs || !x || !z || !out || !x_g2 || !z_g2) throw StatusError(PM_ERR_ARG, "null argument");
        if (world < 1 || rank < 0 || rank >= world) throw StatusError(PM_ERR_ARG, "bad rank/world");
        runtime();
        std::unique_ptr<pm_ctx> h(new pm_ctx());
        ProverCtx& c = h->impl;
        c.rank = rank;
        c.world = world;
        init_dims(c, *r1cs);
        uint64_t rows = 2 * (c.m0 + c.nr);           // Radix2EvaluationDomain::new(rows), generator.rs:60-66
        uint64_t n = 1;
        while (n < rows) n <<= 1;
        c.n = n;
        c.sigma = n + 3;                              // generator.rs:70
        c.allocate_work();
        c.upload_matrix(c.A, r1cs->a_row_ptr, r1cs->a_col, r1cs->a_val, true);
        c.upload_matrix(c.B, r1cs->b_row_ptr, r1cs->b_col, r1cs->b_val, true);
        c.upload_matrix(c.C, r1cs->c_row_ptr, r1cs->c_col, r1cs->c_val, true);
        c.plan_tables();
        c.bases_c.as<G1Affine>(c.plan_c.levels * c.plan_c.stride);
        c.bases_d.as<G1Affine>(c.plan_d.levels * c.plan_d.stride);
        run_setup(c, x, z, x_g2, z_g2);
        c.build_tables();
        *out = h.release();
    });
}

int pm_ctx_dims(const pm_ctx* ctx, uint64_t* n, uint64_t* sigma, uint64_t* num_columns) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        if (n) *n = ctx->impl.n;
        if (sigma) *sigma = ctx->impl.sigma;
        if (num_columns) *num_columns = ctx->impl.cols;
    });
}

int pm_ctx_key_len(const pm_ctx* ctx, int which, uint64_t* len) {
    return guarded([&] {
        if (!ctx || !len) throw StatusError(PM_ERR_ARG, "null argument");
        *len = key_slice(ctx->impl, which).len;
    });
}

int pm_ctx_export_key(const pm_ctx* ctx, int which, uint8_t* out, size_t stride) {
    return guarded([&] {
        if (!ctx || !out || (stride < PM_G1_BYTES && stride != PM_G1_COMPRESSED_BYTES)) throw StatusError(PM_ERR_ARG, "bad export arguments");
        const ProverCtx& c = ctx->impl;
        if (c.world != 1) throw StatusError(PM_ERR_STATE, "key export needs an unsharded context");
        KeySlice ks = key_slice(c, which);
        const G1Affine* src = (ks.in_d ? c.bases_d.get<G1Affine>() : c.bases_c.get<G1Affine>()) + ks.off;
        Runtime& rt = runtime();
        if (stride == PM_G1_COMPRESSED_BYTES) {
            DevBuf dout;
            uint8_t* po = dout.as<uint8_t>(ks.len * 48 + 16);
            launch_g1_compress(src, ks.len, po, rt.stream);
            PM_CUDA(cudaMemcpyAsync(out, po, ks.len * 48, cudaMemcpyDeviceToHost, rt.stream));
            PM_CUDA(cudaStreamSynchronize(rt.stream));
            return;
        }
        if (stride == PM_G1_BYTES) {
            PM_CUDA(cudaMemcpyAsync(out, src, ks.len * sizeof(G1Affine), cudaMemcpyDeviceToHost, rt.stream));
            PM_CUDA(cudaStreamSynchronize(rt.stream));
            return;
        }
        std::vector<uint8_t> tmp(ks.len * PM_G1_BYTES);
        PM_CUDA(cudaMemcpyAsync(tmp.data(), src, tmp.size(), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        static const uint8_t zero96[PM_G1_BYTES] = {0};
        for (uint64_t i = 0; i < ks.len; i++) {
            uint8_t* q = out + i * stride;
            memset(q, 0, stride);
            memcpy(q, tmp.data() + i * PM_G1_BYTES, PM_G1_BYTES);
            if (stride >= 104 && memcmp(q, zero96, PM_G1_BYTES) == 0) q[96] = 1;
        }
    });
}

int pm_ctx_set_assignment(pm_ctx* ctx, const uint8_t* x, const uint8_t* w) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.set_assignment(x, w);
        PM_CUDA(cudaStreamSynchronize(runtime().stream));
    });
}

static void require_unsharded(const ProverCtx& c) {
    if (c.world != 1) throw StatusError(PM_ERR_STATE, "sharded context: use the *_partial / *_finish entry points");
}

int pm_prove_phase1_resident(pm_ctx* ctx, const uint8_t r_a[2 * PM_FR_BYTES], uint8_t a_out[PM_G1_BYTES], uint8_t c_out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        require_unsharded(ctx->impl);
        uint8_t part[2 * sizeof(G1XYZZ)];
        ctx->impl.phase1_partial(r_a, part);
        ctx->impl.phase1_finish(part, 1, a_out, c_out);
    });
}

int pm_prove_phase1_partial(pm_ctx* ctx, const uint8_t r_a[2 * PM_FR_BYTES], uint8_t partials_out[2 * PM_XYZZ_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.phase1_partial(r_a, partials_out);
    });
}
int pm_prove_phase1_finish(pm_ctx* ctx, const uint8_t* gathered, int count, uint8_t a_out[PM_G1_BYTES], uint8_t c_out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.phase1_finish(gathered, count, a_out, c_out);
    });
}
int pm_prove_phase3_partial(pm_ctx* ctx, const uint8_t x2[PM_FR_BYTES], const uint8_t c_at_x1[PM_FR_BYTES], uint8_t partial_out[PM_XYZZ_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.phase3_partial(x2, c_at_x1, partial_out);
    });
}
int pm_prove_phase3_finish(pm_ctx* ctx, const uint8_t* gathered, int count, uint8_t d_out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.phase3_finish(gathered, count, d_out);
    });
}
int pm_nccl_unique_id(const char* libnccl_path, uint8_t id_out[128]) {
    return guarded([&] {
        if (!id_out) throw StatusError(PM_ERR_ARG, "null argument");
        NcclApi& api = nccl_api();
        api.load(libnccl_path);
        NcclUniqueId uid;
        api.check(api.GetUniqueId(&uid), "ncclGetUniqueId");
        memcpy(id_out, uid.internal, sizeof uid.internal);
    });
}
int pm_ctx_attach_nccl(pm_ctx* ctx, const char* libnccl_path, const uint8_t id[128]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.attach_nccl(libnccl_path, id);
    });
}
int pm_ctx_has_collective(const pm_ctx* ctx) { return ctx && ctx->impl.nccl_comm ? 1 : 0; }
int pm_prove_phase1_collective(pm_ctx* ctx, const uint8_t r_a[2 * PM_FR_BYTES], uint8_t a_out[PM_G1_BYTES], uint8_t c_out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.phase1_collective(r_a, a_out, c_out);
    });
}
int pm_prove_phase3_collective(pm_ctx* ctx, const uint8_t x2[PM_FR_BYTES], const uint8_t c_at_x1[PM_FR_BYTES], uint8_t d_out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.phase3_collective(x2, c_at_x1, d_out);
    });
}
int pm_ctx_shard(const pm_ctx* ctx, int* rank, int* world) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        if (rank) *rank = ctx->impl.rank;
        if (world) *world = ctx->impl.world;
    });
}

int pm_prove_phase1(pm_ctx* ctx, const uint8_t* x, const uint8_t* w, const uint8_t r_a[2 * PM_FR_BYTES],
                    uint8_t a_out[PM_G1_BYTES], uint8_t c_out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        require_unsharded(ctx->impl);
        ctx->impl.set_assignment(x, w);
        uint8_t part[2 * sizeof(G1XYZZ)];
        ctx->impl.phase1_partial(r_a, part);
        ctx->impl.phase1_finish(part, 1, a_out, c_out);
    });
}

int pm_prove_phase2(pm_ctx* ctx, const uint8_t x1[PM_FR_BYTES], const uint8_t y1_alpha[PM_FR_BYTES], uint8_t a_at_x1_out[PM_FR_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        ctx->impl.phase2(x1, y1_alpha, a_at_x1_out);
    });
}

int pm_prove_phase3(pm_ctx* ctx, const uint8_t x2[PM_FR_BYTES], const uint8_t c_at_x1[PM_FR_BYTES], uint8_t d_out[PM_G1_BYTES]) {
    return guarded([&] {
        if (!ctx) throw StatusError(PM_ERR_ARG, "null context");
        require_unsharded(ctx->impl);
        uint8_t part[sizeof(G1XYZZ)];
        ctx->impl.phase3_partial(x2, c_at_x1, part);
        ctx->impl.phase3_finish(part, 1, d_out);
    });
}

int pm_ctx_debug_read(pm_ctx* ctx, int which, uint8_t* out, uint64_t capacity_elems, uint64_t* len) {
    return guarded([&] {
        if (!ctx || !len) throw StatusError(PM_ERR_ARG, "null argument");
        ProverCtx& c = ctx->impl;
        const Fr* src = nullptr;
        uint64_t cnt = 0;
        switch (which) {
            case 0: src = c.u.get<Fr>(); cnt = c.n; break;
            case 1: src = c.w.get<Fr>(); cnt = c.n; break;
            case 2: src = c.wu.get<Fr>(); cnt = c.n; break;
            case 3: src = c.u2.get<Fr>(); cnt = 2 * c.n; break;
            case 4: src = c.ztail.get<Fr>(); cnt = c.cols - c.m0; break;
            case 5: src = c.scal_c.get<Fr>(); cnt = c.len_c(); break;
            case 6: src = c.q.get<Fr>(); cnt = c.len_d() - 1; break;
            default: throw StatusError(PM_ERR_ARG, "bad debug selector");
        }
        *len = cnt;
        if (!out) return;
        if (capacity_elems < cnt) throw StatusError(PM_ERR_ARG, "debug buffer too small");
        Runtime& rt = runtime();
        PM_CUDA(cudaMemcpyAsync(out, src, cnt * sizeof(Fr), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
    });
}

int pm_ctx_selftest_resident(pm_ctx* ctx, int virtual_world, uint64_t* mismatches) {
    return guarded([&] {
        if (!ctx || !mismatches) throw StatusError(PM_ERR_ARG, "null argument");
        *mismatches = ctx->impl.selftest_resident(virtual_world);
    });
}

int pm_ctx_phase_ms(const pm_ctx* ctx, double ms[3]) {
    return guarded([&] {
        if (!ctx || !ms) throw StatusError(PM_ERR_ARG, "null argument");
        for (int i = 0; i < 3; i++) ms[i] = ctx->impl.phase_ms[i];
    });
}

}  // extern "C"
