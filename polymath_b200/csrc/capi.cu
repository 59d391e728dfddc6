// C ABI of libpolymath_b200.so — standalone kernel entry points (include/polymath_b200.h).
// Host buffers in, host buffers out; every call stages through device memory owned here.
#include <mutex>
#include <cstdlib>
#include <vector>

#include "../../include/polymath_b200.h"
#include "common.cuh"
#include "field_kernels.cuh"
#include "fixed_base.cuh"
#include "g1_codec.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "runtime.cuh"
#include "synth.cuh"
#include "host/g1_host.hpp"

namespace pm {

const std::string& last_error();

Runtime& runtime() {
    static Runtime rt;
    return rt;
}

std::recursive_mutex& api_mutex() {
    static std::recursive_mutex m;
    return m;
}

Runtime::Runtime() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        throw CudaError(std::string("no CUDA device available (the product path has no CPU fallback): ") +
                        cudaGetErrorString(e));
    // The bucket accumulation gathers 48- / 96-byte records at random from multi-GB tables: with the default L2 fetch
    // granularity every miss pulls a whole 128-byte line (measured 160 / 223 bytes of DRAM traffic per gathered x / point,
    // profiles/r2_summary.md).  PM_L2_FETCH = 32 | 64 | 128 sets the hint (tuning hook).
    if (const char* v = getenv("PM_L2_FETCH")) {
        const int g = atoi(v);
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
    }
    PM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    PM_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    PM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    PM_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
}

uint64_t Runtime::total_launches() const {
    return extra_launches + ntt.launches + msm.launches + msm2.launches + fixed_base.launches;
}

// Pack caller-strided affine points into the device layout (96 B, (0,0) = infinity).
void pack_points_host(const uint8_t* src, size_t stride, size_t n, std::vector<uint8_t>& dst) {
    dst.resize(n * PM_G1_BYTES);
    for (size_t i = 0; i < n; i++) {
        const uint8_t* p = src + i * stride;
        uint8_t* q = dst.data() + i * PM_G1_BYTES;
        if (stride >= 104 && p[96] != 0) memset(q, 0, PM_G1_BYTES);
        else memcpy(q, p, PM_G1_BYTES);
    }
}

template <class F>
static int field_batch(FieldOp op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, bool fq) {
    return guarded([&] {
        if (n == 0) return;
        if (!a || !b || !out) throw StatusError(PM_ERR_ARG, "null buffer");
        Runtime& rt = runtime();
        DevBuf da, db, dc;
        F* pa = da.as<F>(n);
        F* pb = db.as<F>(n);
        F* pc = dc.as<F>(n);
        PM_CUDA(cudaMemcpyAsync(pa, a, n * sizeof(F), cudaMemcpyHostToDevice, rt.stream));
        PM_CUDA(cudaMemcpyAsync(pb, b, n * sizeof(F), cudaMemcpyHostToDevice, rt.stream));
        if constexpr (sizeof(F) == sizeof(Fr)) launch_fr_batch(op, (const Fr*)pa, (const Fr*)pb, (Fr*)pc, n, rt.stream);
        else launch_fq_batch(op, (const Fq*)pa, (const Fq*)pb, (Fq*)pc, n, rt.stream);
        rt.extra_launches++;
        PM_CUDA(cudaMemcpyAsync(out, pc, n * sizeof(F), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        (void)fq;
    });
}

}  // namespace pm

using namespace pm;

extern "C" {

const char* pm_last_error(void) { return last_error().c_str(); }
int pm_abi_version(void) { return 1; }
// Keep the context's local-memory pool at its high-water mark (cudaDeviceLmemResizeToMax): several kernels here run
// with stack frames of up to 384 bytes per thread, and by default the driver may shrink the pool after such a kernel
// and grow it again at a later launch — a device-wide synchronisation plus reallocation in the middle of a phase.
// Must run before the device's primary context is created (first CUDA call of the process on that device); later
// it can only report whether the flag is in place.  Returns 1 if the flag is set, 0 if not, < 0 on error.
int pm_runtime_configure(int device) {
    unsigned flags = 0;
    int dev = device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    // runtime API first; when the primary context is already active the driver-level call below still applies
    cudaError_t e = cudaSetDevice(dev) == cudaSuccess ? cudaSetDeviceFlags(cudaDeviceLmemResizeToMax) : cudaErrorInvalidDevice;
    (void)e;
    cudaGetLastError();
    if (cudaGetDeviceFlags(&flags) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (flags & cudaDeviceLmemResizeToMax) ? 1 : 0;
}

int pm_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}
uint64_t pm_kernel_launches(void) {
    uint64_t v = 0;
    guarded([&] { v = runtime().total_launches(); });
    return v;
}

int pm_fr_mul_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) { return field_batch<Fr>(FieldOp::Mul, a, b, out, n, false); }
int pm_fr_add_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) { return field_batch<Fr>(FieldOp::Add, a, b, out, n, false); }
int pm_fr_sub_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) { return field_batch<Fr>(FieldOp::Sub, a, b, out, n, false); }
int pm_fq_mul_batch(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) { return field_batch<Fq>(FieldOp::Mul, a, b, out, n, true); }
int pm_fr_inv_batch(const uint8_t* a, uint8_t* out, size_t n) { return field_batch<Fr>(FieldOp::Inv, a, a, out, n, false); }
int pm_fq_inv_batch(const uint8_t* a, uint8_t* out, size_t n) { return field_batch<Fq>(FieldOp::Inv, a, a, out, n, true); }

int pm_ntt_fr(uint8_t* data, unsigned log_n, int inverse, const uint8_t* coset_gen) {
    return guarded([&] {
        if (!data || log_n > 32) throw StatusError(PM_ERR_ARG, "bad ntt arguments");
        Runtime& rt = runtime();
        const size_t n = (size_t)1 << log_n;
        DevBuf d, g;
        Fr* p = d.as<Fr>(n);
        PM_CUDA(cudaMemcpyAsync(p, data, n * sizeof(Fr), cudaMemcpyHostToDevice, rt.stream));
        Fr* gd = nullptr;
        if (coset_gen) {
            // forward: scale by g^i then transform; inverse: transform then scale by g^-i (caller passes g; we invert on device)
            gd = g.as<Fr>(2);
            PM_CUDA(cudaMemcpyAsync(gd, coset_gen, sizeof(Fr), cudaMemcpyHostToDevice, rt.stream));
        }
        if (coset_gen && !inverse) { launch_scale_by_powers(p, n, gd, rt.stream); rt.extra_launches++; }
        rt.ntt.run(p, (int)log_n, inverse != 0, rt.stream);
        if (coset_gen && inverse) {
            launch_fr_inverse(gd, gd + 1, rt.stream);
            launch_scale_by_powers(p, n, gd + 1, rt.stream);
            rt.extra_launches += 2;
        }
        PM_CUDA(cudaMemcpyAsync(data, p, n * sizeof(Fr), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
    });
}

int pm_ntt_dist_local(void* data_dev, void* send_dev, unsigned log_n, unsigned log_g, unsigned rank, int inverse, void* cuda_stream) {
    return guarded([&] {
        if (!data_dev || !send_dev) throw StatusError(PM_ERR_ARG, "null device buffer");
        Runtime& rt = runtime();
        cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);   // NULL = the default stream, as in the CUDA runtime
        rt.ntt.dist_local(static_cast<Fr*>(data_dev), static_cast<Fr*>(send_dev), (int)log_n, (int)log_g, rank, inverse != 0, s);
    });
}
int pm_ntt_dist_combine(const void* recv_dev, void* out_dev, unsigned log_n, unsigned log_g, int inverse, void* cuda_stream) {
    return guarded([&] {
        if (!recv_dev || !out_dev) throw StatusError(PM_ERR_ARG, "null device buffer");
        Runtime& rt = runtime();
        cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
        rt.ntt.dist_combine(static_cast<const Fr*>(recv_dev), static_cast<Fr*>(out_dev), (int)log_n, (int)log_g, inverse != 0, s);
    });
}

static int msm_host_call(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, int window_bits,
                         int heavy_threshold, int levels, uint8_t out[PM_G1_BYTES]);

int pm_msm_g1_window(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, int window_bits,
                     int heavy_threshold, uint8_t out[PM_G1_BYTES]) {
    return msm_host_call(bases, base_stride, scalars, n, window_bits, heavy_threshold, 1, out);
}
int pm_msm_g1_levels(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, int window_bits,
                     int levels, uint8_t out[PM_G1_BYTES]) {
    return msm_host_call(bases, base_stride, scalars, n, window_bits, 0, levels, out);
}

static int msm_host_call(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, int window_bits,
                         int heavy_threshold, int levels, uint8_t out[PM_G1_BYTES]) {
    return guarded([&] {
        if (levels < 1) levels = 1;
        if (levels > 1 && window_bits <= 0) throw StatusError(PM_ERR_ARG, "levels need an explicit window");
        if (!out || (n && (!bases || !scalars)) || base_stride < PM_G1_BYTES) throw StatusError(PM_ERR_ARG, "bad msm arguments");
        Runtime& rt = runtime();
        std::vector<uint8_t> packed;
        const uint8_t* src = bases;
        if (base_stride != PM_G1_BYTES) { pack_points_host(bases, base_stride, n, packed); src = packed.data(); }
        DevBuf db, ds, dres;
        G1Affine* pb = db.as<G1Affine>((n ? n : 1) * (size_t)levels);
        Fr* ps = ds.as<Fr>(n ? n : 1);
        G1XYZZ* wins = dres.as<G1XYZZ>(kMaxMsmSums);
        PM_CUDA(cudaMemcpyAsync(pb, src, n * sizeof(G1Affine), cudaMemcpyHostToDevice, rt.stream));
        PM_CUDA(cudaMemcpyAsync(ps, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, rt.stream));
        launch_build_levels(pb, n, levels, n, window_bits, rt.stream);
        MsmConfig cfg;
        cfg.c = window_bits;
        cfg.heavy = heavy_threshold;
        cfg.levels = levels;
        cfg.level_stride = n;
        MsmEngine::Shape sh = rt.msm.run(pb, ps, n, wins, rt.stream, cfg);
        std::vector<uint8_t> hw((size_t)sh.count() * sizeof(G1XYZZ));
        PM_CUDA(cudaMemcpyAsync(hw.data(), wins, hw.size(), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        host::xyzz_to_affine_wire(host::combine_shifted(hw.data(), sh.nwin, sh.c, sh.nsum, sh.shift), out);
    });
}

int pm_msm_set_tuning(int rounds) {
    return guarded([&] { MsmEngine::set_tuning(rounds); });
}

int pm_msm_g1(const uint8_t* bases, size_t base_stride, const uint8_t* scalars, size_t n, uint8_t out[PM_G1_BYTES]) {
    return pm_msm_g1_window(bases, base_stride, scalars, n, 0, 0, out);
}

int pm_g1_decompress_batch(const uint8_t* in, size_t n, int validate, uint8_t* out) {
    return guarded([&] {
        if (n == 0) return;
        if (!in || !out) throw StatusError(PM_ERR_ARG, "null buffer");
        Runtime& rt = runtime();
        DevBuf din, dout, dbad;
        uint8_t* pi = din.as<uint8_t>(n * 48);
        G1Affine* po = dout.as<G1Affine>(n);
        unsigned long long* bad = dbad.as<unsigned long long>(1);
        PM_CUDA(cudaMemcpyAsync(pi, in, n * 48, cudaMemcpyHostToDevice, rt.stream));
        PM_CUDA(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), rt.stream));
        launch_g1_decompress(pi, n, validate != 0, po, bad, rt.stream);
        rt.extra_launches++;
        unsigned long long hbad = 0;
        PM_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof hbad, cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaMemcpyAsync(out, po, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        if (hbad != ~0ull)
            throw StatusError(PM_ERR_ARG, "invalid compressed G1 point at index " + std::to_string(hbad >> 3) + ": " +
                                              g1_decode_status_name((unsigned)(hbad & 7)));
    });
}

int pm_g1_compress_batch(const uint8_t* in, size_t n, uint8_t* out) {
    return guarded([&] {
        if (n == 0) return;
        if (!in || !out) throw StatusError(PM_ERR_ARG, "null buffer");
        Runtime& rt = runtime();
        DevBuf din, dout;
        G1Affine* pi = din.as<G1Affine>(n);
        uint8_t* po = dout.as<uint8_t>(n * 48);
        PM_CUDA(cudaMemcpyAsync(pi, in, n * sizeof(G1Affine), cudaMemcpyHostToDevice, rt.stream));
        launch_g1_compress(pi, n, po, rt.stream);
        rt.extra_launches++;
        PM_CUDA(cudaMemcpyAsync(out, po, n * 48, cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
    });
}

int pm_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out) {
    return guarded([&] {
        if (n == 0) return;
        if (!scalars || !out) throw StatusError(PM_ERR_ARG, "null buffer");
        Runtime& rt = runtime();
        DevBuf ds, dout;
        Fr* ps = ds.as<Fr>(n);
        G1Affine* po = dout.as<G1Affine>(n);
        PM_CUDA(cudaMemcpyAsync(ps, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, rt.stream));
        rt.fixed_base.run(ps, n, po, rt.stream);
        PM_CUDA(cudaMemcpyAsync(out, po, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, rt.stream));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
    });
}

static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
int pm_timer_start(void) {
    return guarded([&] {
        Runtime& rt = runtime();
        if (!g_t0) { PM_CUDA(cudaEventCreate(&g_t0)); PM_CUDA(cudaEventCreate(&g_t1)); }
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        PM_CUDA(cudaEventRecord(g_t0, rt.stream));
    });
}
int pm_timer_stop(double* ms) {
    return guarded([&] {
        Runtime& rt = runtime();
        if (!g_t0 || !ms) throw StatusError(PM_ERR_STATE, "timer not started");
        PM_CUDA(cudaEventRecord(g_t1, rt.stream));
        PM_CUDA(cudaEventSynchronize(g_t1));
        float f = 0;
        PM_CUDA(cudaEventElapsedTime(&f, g_t0, g_t1));
        *ms = f;
    });
}
int pm_bench_set_kernel_timing(int enable) {
    return guarded([&] {
        Runtime& rt = runtime();
        rt.msm.time_accumulate = enable != 0;
        rt.ntt.time_passes = enable != 0;
    });
}
int pm_bench_last_kernel_ms(double ms[2]) {
    return guarded([&] {
        Runtime& rt = runtime();
        ms[0] = ms[1] = 0;
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        float f = 0;
        if (rt.msm.ev_acc_begin && cudaEventElapsedTime(&f, rt.msm.ev_acc_begin, rt.msm.ev_acc_end) == cudaSuccess) ms[0] = f;
        if (rt.ntt.ev_begin && cudaEventElapsedTime(&f, rt.ntt.ev_begin, rt.ntt.ev_end) == cudaSuccess) ms[1] = f;
        cudaGetLastError();
    });
}

int pm_bench_last_msm(double out[4]) {
    return guarded([&] {
        Runtime& rt = runtime();
        out[0] = out[1] = out[2] = out[3] = 0;
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        float f = 0;
        if (rt.msm.ev_acc_begin && cudaEventElapsedTime(&f, rt.msm.ev_acc_begin, rt.msm.ev_acc_end) == cudaSuccess) out[0] = f;
        if (rt.msm.last_rounds > 0 && rt.msm.ev_bwd_begin &&
            cudaEventElapsedTime(&f, rt.msm.ev_bwd_begin, rt.msm.ev_bwd_end) == cudaSuccess) out[1] = f;
        out[2] = rt.msm.last_rounds;
        out[3] = (double)rt.msm.last_entries;
        cudaGetLastError();
    });
}

int pm_bench_imad_peak(double* mads_per_s) {
    return guarded([&] { runtime(); *mads_per_s = measure_imad_peak(2000); });
}

int pm_bench_field_mul(int field, double* muls_per_s) {
    return guarded([&] {
        runtime();
        switch (field) {
            case 0: *muls_per_s = measure_fr_mul_rate(4000); break;
            case 1: *muls_per_s = measure_fq_mul_rate(2000); break;
            case 2: *muls_per_s = measure_fq_variant_rate(1, 2000); break;   // Fq squaring
            case 3: *muls_per_s = measure_fq_variant_rate(2, 2000); break;   // Fq Karatsuba product
            case 4: *muls_per_s = measure_fr_variant_rate(1, 4000); break;   // Fr squaring
            case 5: *muls_per_s = measure_fr_variant_rate(2, 4000); break;   // Fr Karatsuba product
            default: throw StatusError(PM_ERR_ARG, "unknown field / variant");
        }
    });
}

int pm_bench_ntt(unsigned log_n, int inverse, int iters, double* ms_avg) {
    return guarded([&] {
        if (log_n > 30 || iters <= 0 || !ms_avg) throw StatusError(PM_ERR_ARG, "bad bench arguments");
        Runtime& rt = runtime();
        const size_t n = (size_t)1 << log_n;
        DevBuf d;
        Fr* p = d.as<Fr>(n);
        launch_fill_fr(p, n, 0x1234 + log_n, rt.stream);
        // bit 1 of `inverse`: the coset variant of pm_ntt_fr (scale by g^i before a forward / by g^-i after an inverse transform)
        const bool inv = (inverse & 1) != 0, coset = (inverse & 2) != 0;
        DevBuf g;
        Fr* gd = nullptr;
        if (coset) {
            gd = g.as<Fr>(2);
            launch_fill_fr(gd, 1, 0x7, rt.stream);
            launch_fr_inverse(gd, gd + 1, rt.stream);
        }
        auto once = [&]() {
            if (coset && !inv) launch_scale_by_powers(p, n, gd, rt.stream);
            rt.ntt.run(p, (int)log_n, inv, rt.stream);
            if (coset && inv) launch_scale_by_powers(p, n, gd + 1, rt.stream);
        };
        once();  // warm-up (builds twiddle tables)
        cudaEvent_t e0, e1;
        PM_CUDA(cudaEventCreate(&e0));
        PM_CUDA(cudaEventCreate(&e1));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        PM_CUDA(cudaEventRecord(e0, rt.stream));
        for (int i = 0; i < iters; i++) once();
        PM_CUDA(cudaEventRecord(e1, rt.stream));
        PM_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms_avg = ms / iters;
    });
}

int pm_bench_g1_codec(size_t n, double* ms_decompress, double* ms_compress) {
    return guarded([&] {
        if (n == 0 || !ms_decompress || !ms_compress) throw StatusError(PM_ERR_ARG, "bad bench arguments");
        Runtime& rt = runtime();
        DevBuf ds, dp, dc, dbad;
        Fr* ps = ds.as<Fr>(n);
        G1Affine* pp = dp.as<G1Affine>(n);
        uint8_t* pc = dc.as<uint8_t>(n * 48);
        unsigned long long* bad = dbad.as<unsigned long long>(1);
        launch_fill_fr(ps, n, 0xc0dec, rt.stream);
        rt.fixed_base.run(ps, n, pp, rt.stream);       // points = [s_i]G
        cudaEvent_t e0, e1, e2;
        PM_CUDA(cudaEventCreate(&e0)); PM_CUDA(cudaEventCreate(&e1)); PM_CUDA(cudaEventCreate(&e2));
        launch_g1_compress(pp, n, pc, rt.stream);      // warm-up
        PM_CUDA(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), rt.stream));
        PM_CUDA(cudaEventRecord(e0, rt.stream));
        launch_g1_compress(pp, n, pc, rt.stream);
        PM_CUDA(cudaEventRecord(e1, rt.stream));
        launch_g1_decompress(pc, n, false, pp, bad, rt.stream);
        PM_CUDA(cudaEventRecord(e2, rt.stream));
        PM_CUDA(cudaEventSynchronize(e2));
        float fc = 0, fd = 0;
        PM_CUDA(cudaEventElapsedTime(&fc, e0, e1));
        PM_CUDA(cudaEventElapsedTime(&fd, e1, e2));
        unsigned long long hbad = 0;
        PM_CUDA(cudaMemcpy(&hbad, bad, sizeof hbad, cudaMemcpyDeviceToHost));
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (hbad != ~0ull) throw StatusError(PM_ERR_STATE, "codec bench: round trip failed");
        *ms_compress = fc;
        *ms_decompress = fd;
    });
}

int pm_bench_fixed_base(size_t n, int iters, double* ms_avg) {
    return guarded([&] {
        if (n == 0 || iters <= 0 || !ms_avg) throw StatusError(PM_ERR_ARG, "bad bench arguments");
        Runtime& rt = runtime();
        DevBuf ds, dp;
        Fr* ps = ds.as<Fr>(n);
        G1Affine* pp = dp.as<G1Affine>(n);
        launch_fill_fr(ps, n, 0xf1bed, rt.stream);
        rt.fixed_base.run(ps, n, pp, rt.stream);       // warm-up (builds the window table)
        cudaEvent_t e0, e1;
        PM_CUDA(cudaEventCreate(&e0));
        PM_CUDA(cudaEventCreate(&e1));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        PM_CUDA(cudaEventRecord(e0, rt.stream));
        for (int i = 0; i < iters; i++) rt.fixed_base.run(ps, n, pp, rt.stream);
        PM_CUDA(cudaEventRecord(e1, rt.stream));
        PM_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms_avg = ms / iters;
    });
}

static int g_bench_msm_skew = 0;
int pm_bench_set_msm_skew(int mode) {
    g_bench_msm_skew = mode;
    return PM_OK;
}

int pm_bench_msm(size_t n, int window_bits, int iters, double* ms_avg, double* ms_accumulate) {
    return pm_bench_msm_levels(n, window_bits, 1, iters, ms_avg, ms_accumulate);
}

int pm_bench_msm_levels(size_t n, int window_bits, int levels, int iters, double* ms_avg, double* ms_accumulate) {
    return guarded([&] {
        if (levels < 1) levels = 1;
        if (levels > 1 && window_bits <= 0) throw StatusError(PM_ERR_ARG, "levels need an explicit window");
        if (n == 0 || iters <= 0 || !ms_avg) throw StatusError(PM_ERR_ARG, "bad bench arguments");
        Runtime& rt = runtime();
        DevBuf db, ds, dres;
        G1Affine* pb = db.as<G1Affine>(n * (size_t)levels);
        Fr* ps = ds.as<Fr>(n);
        G1XYZZ* acc = dres.as<G1XYZZ>(kMaxMsmSums);
        launch_fill_fr(ps, n, 0xabcdef, rt.stream);
        rt.fixed_base.run(ps, n, pb, rt.stream);       // bases = [s_i]G for pseudo-random s_i
        launch_fill_fr(ps, n, 0x5eed, rt.stream);      // uniform scalars
        if (g_bench_msm_skew) launch_skew_msm_inputs(ps, pb, n, 0x5ca1ab1e, rt.stream);   // before the levels: infinity stays infinity
        launch_build_levels(pb, n, levels, n, window_bits, rt.stream);
        MsmConfig cfg;
        cfg.c = window_bits;
        cfg.levels = levels;
        cfg.level_stride = n;
        rt.msm.time_accumulate = true;
        rt.msm.run(pb, ps, n, acc, rt.stream, cfg);    // warm-up
        cudaEvent_t e0, e1;
        PM_CUDA(cudaEventCreate(&e0));
        PM_CUDA(cudaEventCreate(&e1));
        PM_CUDA(cudaStreamSynchronize(rt.stream));
        double acc_ms = 0;
        float total = 0;
        for (int i = 0; i < iters; i++) {
            PM_CUDA(cudaEventRecord(e0, rt.stream));
            rt.msm.run(pb, ps, n, acc, rt.stream, cfg);
            PM_CUDA(cudaEventRecord(e1, rt.stream));
            PM_CUDA(cudaEventSynchronize(e1));
            float ms = 0, ams = 0;
            PM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            PM_CUDA(cudaEventElapsedTime(&ams, rt.msm.ev_acc_begin, rt.msm.ev_acc_end));
            total += ms;
            acc_ms += ams;
        }
        rt.msm.time_accumulate = false;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms_avg = total / iters;
        if (ms_accumulate) *ms_accumulate = acc_ms / iters;
    });
}

}  // extern "C"
