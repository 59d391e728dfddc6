#!/usr/bin/env python
"""bench.py — Polymath prove latency on synthetic SAP circuits (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--log-n 20]

One "step" = one `Polymath::prove` of the S-mimc(2^log_n) circuit with the proving
key resident on the device.  `value` times K proves whose witness is already resident in HBM
(CUDA events on the library's stream, host transcript round-trips included); `e2e` times the same K
proves through the public host-buffer call (`pm_polymath_prove`: H2D of instance+witness from
pinned memory, D2H of the proof pieces, inside the timed region).  One JSON line on stdout.

`--impl reference` times the CPU restatement of the reference's arkworks path (oracle/cpu_ref.cpp,
OpenMP on all host cores) on a bounded sample of the same workload, scaled to the metric's unit.
"""
import argparse
import ctypes as C
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC_FMT = "prove_ms_2p{log_n}_sap_constraints"
TRAFFIC_BWD = 37.20e9   # dram__bytes_read.sum + dram__bytes_write.sum of that launch (30.34 + 6.85 GB), ncu --set full (profiles/prof_bwd_r1_k_details.csv)


# --------------------------------------------------------------------------------------------
# CPU baseline (oracle port) — bounded sample, scaled
# --------------------------------------------------------------------------------------------
def ark_window(n):
    if n < 32:
        return 3
    return ((n - 1).bit_length()) * 69 // 100 + 2


def cpu_baseline(log_n, msm_log=None, ntt_log=None):
    """Time the C++/OpenMP port on a bounded sample and scale to one prove at n = 2^log_n.

    prove(n) = MSMs over ~14n + 29 points + 3 iNTT(n) + NTT(2n) + iNTT(2n).
    MSM cost is scaled by points x windows (arkworks window rule at each size); NTT by (N/2) log2 N.
    """
    # sample sizes: 2^21-point MSM + 2^22 NTT (about 2-3 s on 16 threads); PM_REF_SAMPLE_LOG shrinks both (CPU tests)
    shrink = os.environ.get("PM_REF_SAMPLE_LOG")
    if msm_log is None:
        msm_log = int(shrink) if shrink else 21
    if ntt_log is None:
        ntt_log = int(shrink) + 1 if shrink else 22
    from oracle import cpp
    cpp.use_all_cores()          # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses the host's cores
    n = 1 << log_n
    m = 1 << msm_log
    rnd = random.Random(7)
    bases = cpp.make_bases_wire(m)
    scalars = bytearray(os.urandom(32 * m))
    scalars[31::32] = bytes(b & 0x3F for b in scalars[31::32])     # < 2^254 < r: valid Montgomery limbs
    scalars = bytes(scalars)
    cpp.msm_wire(bases[:96 * 1024], scalars[:32 * 1024], 1024)   # warm up threads
    t0 = time.perf_counter()
    cpp.msm_wire(bases, scalars, m)
    t_msm = time.perf_counter() - t0
    buf = bytearray(os.urandom(32 << ntt_log))
    buf[31::32] = bytes(b & 0x3F for b in buf[31::32])
    t0 = time.perf_counter()
    cpp.ntt_wire(buf, ntt_log, False)
    t_ntt = time.perf_counter() - t0

    def msm_work(pts):
        c = ark_window(pts)
        return pts * ((255 + c - 1) // c)

    sizes = [n + 4, 3 * n + 5, 10 * n + 22]     # the three MSM launches of one prove (a, c, d)
    msm_s = sum(msm_work(s) for s in sizes) / msm_work(m) * t_msm

    def ntt_work(lg):
        return (1 << lg) // 2 * lg

    ntt_s = (3 * ntt_work(log_n) + 2 * ntt_work(log_n + 1)) / ntt_work(ntt_log) * t_ntt
    total_ms = (msm_s + ntt_s) * 1e3
    return {
        "value": total_ms, "unit": "ms", "cores": cpp.num_threads(), "kind": "port",
        "sample": "oracle/cpu_ref.cpp (OpenMP): one G1 MSM of 2^%d points (%.2f s) and one Fr NTT of 2^%d (%.3f s), "
                  "scaled by points*windows resp. (N/2)log2N to one prove at n=2^%d (MSMs of n+4, 3n+5, 10n+22 points; "
                  "3 iNTT(n) + NTT(2n) + iNTT(2n)); SpMV/scan terms omitted" % (msm_log, t_msm, ntt_log, t_ntt, log_n),
        "msm_mpts_per_s": m / t_msm / 1e6, "ntt_gelem_per_s": (1 << ntt_log) / t_ntt / 1e9,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    last = None
    for _ in range(args.warmup):
        cpu_baseline(args.log_n, msm_log=14, ntt_log=16)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_baseline(args.log_n)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    last["value"] = value
    out = {
        "impl": "reference", "metric": METRIC_FMT.format(log_n=args.log_n), "value": value, "unit": "ms",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "S-mimc(2^%d) prove, CPU port of the arkworks primitives on a bounded sample (see cpu_baseline.sample)" % args.log_n,
                   "log_n": args.log_n},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampling
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    In-process NVML (pynvml: one handle, two light queries per sample) — a looping `nvidia-smi` process costs
    milliseconds of driver time per query and showed up as +6.5 ms per prove at n = 2^16 (profiles/r1_h_summary.md);
    nvidia-smi is only the fallback when pynvml is missing.  The sampler starts before the warm-up; `mark()` is called
    at the start of the timed region and only later samples count.
    """
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period_s=0.1):
        self.index = index
        self.period = period_s
        self.samples = []          # (sm_mhz, sm_max_mhz, [reason names])
        self.first = 0
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.source = None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        # LOCAL_RANK indexes the visible devices; NVML indexes the physical ones
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                idx = int(ids[self.index])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, 0), (nv.nvmlClocksEventReasonHwThermalSlowdown, 1),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, 2), (nv.nvmlClocksEventReasonSwPowerCap, 3)]
        self.source = "nvml"
        while not self.stop_flag.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.samples.append((int(sm), int(mx), [self.NAMES[i] for b, i in bits if r & b]))
            self.stop_flag.wait(self.period)
        nv.nvmlShutdown()

    def _run_smi(self):
        self.source = "nvidia-smi"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits", "-lms", "500"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) >= 6 and parts[0].isdigit() and parts[1].isdigit():
                self.samples.append((int(parts[0]), int(parts[1]),
                                     [self.NAMES[i] for i in range(4) if parts[2 + i].lower().startswith("active")]))
            if self.stop_flag.is_set():
                break

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def wait_first(self, timeout):
        t0 = time.perf_counter()
        while not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are only used if none arrives later."""
        self.first = len(self.samples)

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        if self.thread:
            self.thread.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        timed = self.samples[self.first:]
        use = timed if timed else self.samples[-1:]      # a run shorter than the sampling period: last warm-up sample
        sm = sorted(s[0] for s in use)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[1] for s in use),
                "reasons": sorted({r for s in use for r in s[2]}), "samples": len(timed),
                "samples_incl_warmup": len(self.samples), "source": self.source}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    lmem_flag = None
    if not os.environ.get("PM_BENCH_NO_LMEM_FLAG"):
        from polymath_b200.lib import load as _load
        _l = _load()
        _l.pm_runtime_configure.argtypes = [C.c_int]
        lmem_flag = _l.pm_runtime_configure(local_rank)       # before torch creates the device's primary context
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl")

    from polymath_b200 import circuits
    from polymath_b200.api import Polymath, StdRng, _lib
    from polymath_b200 import codec
    from polymath_b200.lib import check
    lib = _lib()

    log_n = args.log_n
    n = 1 << log_n
    r1cs, instance, witness, rng = circuits.synthetic_mimc(n, seed=1)
    t0 = time.perf_counter()
    if world > 1:
        from polymath_b200 import sharded
        prover = sharded.ShardedProver(r1cs, rng, rank, world)
        vk_bytes = prover.vk_bytes
    else:
        pk, vk_bytes = Polymath.setup(r1cs, rng)
        prover = None
    setup_s = time.perf_counter() - t0

    inst_wire = codec.frs_to_wire(instance)
    wit_wire = codec.frs_to_wire(witness)
    # pinned host buffers for the e2e leg
    pin_inst = torch.empty(len(inst_wire), dtype=torch.uint8).pin_memory()
    pin_wit = torch.empty(len(wit_wire), dtype=torch.uint8).pin_memory()
    pin_inst.copy_(torch.frombuffer(bytearray(inst_wire), dtype=torch.uint8))
    pin_wit.copy_(torch.frombuffer(bytearray(wit_wire), dtype=torch.uint8))
    p_inst = C.c_char_p(pin_inst.data_ptr())
    p_wit = C.c_char_p(pin_wit.data_ptr())
    proof = C.create_string_buffer(176)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def prove_resident():
        if prover is not None:
            return prover.prove_resident(p_inst, rng)
        check(lib.pm_polymath_prove_resident(pk._h, p_inst, rng._h, proof))
        return proof.raw

    def prove_e2e():
        if prover is not None:
            return prover.prove(p_inst, p_wit, rng)
        check(lib.pm_polymath_prove(pk._h, p_inst, p_wit, rng._h, proof))
        return proof.raw

    # resident leg ------------------------------------------------------------------------
    if prover is not None:
        prover.set_assignment(p_inst, p_wit)
    else:
        check(lib.pm_ctx_set_assignment(pk._h, p_inst, p_wit))
    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation) contends with CUDA calls for the
    # driver, which would otherwise fall into the timed region of a short run; only samples taken after the mark count
    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("PM_BENCH_NO_CLOCKS")) else None
    if sampler:
        sampler.start()
        sampler.wait_first(5.0)
    for _ in range(args.warmup):
        prove_resident()
    check(lib.pm_bench_set_kernel_timing(1))
    launches0 = lib.pm_kernel_launches()
    barrier()
    if sampler:
        sampler.mark()
    check(lib.pm_timer_start())
    wall0 = time.perf_counter()
    acc_ms, bwd_ms, msm_geom = [], [], (0, 0)
    for _ in range(args.steps):
        last_proof = prove_resident()
        km = (C.c_double * 4)()
        check(lib.pm_bench_last_msm(km))        # the [d]_1 MSM is the last one of a prove
        acc_ms.append(km[0])
        bwd_ms.append(km[1])
        msm_geom = (int(km[2]), int(km[3]))
    ms = C.c_double()
    check(lib.pm_timer_stop(C.byref(ms)))
    barrier()
    wall_resident = (time.perf_counter() - wall0) * 1e3
    launches = lib.pm_kernel_launches() - launches0
    check(lib.pm_bench_set_kernel_timing(0))
    dev_ms = ms.value
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    value = dev_ms / args.steps
    phase_ms = pk.phase_ms() if prover is None else prover.phase_ms()

    # e2e leg -----------------------------------------------------------------------------
    prove_e2e()
    barrier()
    check(lib.pm_timer_start())
    for _ in range(args.steps):
        prove_e2e()
    check(lib.pm_timer_stop(C.byref(ms)))
    barrier()
    clocks = sampler.stop() if sampler else None          # rank 0, sampled over both timed legs (resident + e2e)
    e2e_ms = ms.value
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = e2e_ms / args.steps

    # standalone kernel sweep across the ranks (BASELINE.json configs[4]): sharded Fr NTT with one NCCL all-to-all,
    # G1 MSM split by point range (every rank a 1/world share; the 192-byte partial sums are not timed)
    sweep_dist = {}
    if world > 1:
        from polymath_b200 import sharded as _sh
        ntt_log = 24
        sn = _sh.ShardedNtt(ntt_log, rank, world)
        sn.data.random_(0, 256)
        sn.data.view(-1, 32)[:, 31] &= 0x3F                 # < 2^254 < r: valid Montgomery limbs
        sn.run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            sn.run()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sweep_dist["fr_ntt_gelem_per_s_2p%d_sharded" % ntt_log] = (1 << ntt_log) / (float(t.item()) * 1e-3) / 1e9
        dms, dacc = C.c_double(), C.c_double()
        check(lib.pm_bench_msm((1 << 24) // world, 0, 2, C.byref(dms), C.byref(dacc)))
        t = torch.tensor([dms.value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sweep_dist["g1_msm_mpts_per_s_2p24_sharded"] = (1 << 24) / (float(t.item()) * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel ---------------------------------------------------------------
    # The [d]_1 MSM (10n + 22 points) is 2/3 of a prove; its bucket accumulation runs as R batched-affine
    # pair rounds + a short XYZZ walk.  The heaviest single launch is the first round's k_pairs_backward:
    # half of all bucket additions of the MSM (entries / 2 affine additions).  Algorithmic convention of
    # one bucket addition = one XYZZ mixed addition = 10 Fq-modmul x 300 IMAD; the kernel
    # executes 5 products per addition (plus 1 in k_pairs_forward), reported as `executed_frac`.
    d = C.c_double()
    check(lib.pm_bench_imad_peak(C.byref(d)))
    imad_peak = d.value
    d_points = (10 * n + 22) // world
    c_ark = ark_window(10 * n + 22)
    w_ark = (255 + c_ark - 1) // c_ark
    med = lambda v: sorted(v)[len(v) // 2] if v else 0.0
    acc, bwd = med(acc_ms), med(bwd_ms)
    rounds, entries = msm_geom
    stage_imad = d_points * w_ark * 10 * 300          # whole stage: N*W madds x 10 Fq-modmul x 300 IMAD
    stage = {
        "kernels": "k_pairs_forward / k_invert_* / k_pairs_backward x %d rounds + k_accumulate_rounds" % rounds if rounds
                   else "k_accumulate (XYZZ walk)",
        "ms": acc, "share_of_step": acc / value if value else None,
        "achieved": stage_imad / (acc * 1e-3) / 1e12 if acc > 0 else None, "unit": "TIMAD/s",
        "frac": stage_imad / (acc * 1e-3) / imad_peak if acc > 0 else None,
        "algorithmic": "%d points x %d windows (arkworks window rule c=%d) x 10 Fq-modmul x 300 IMAD" % (d_points, w_ark, c_ark),
    }
    if rounds and bwd > 0:
        adds = entries // 2
        algo_imad = adds * 10 * 300
        roofline = {
            "kernel": "k_pairs_backward<first round> (batched-affine bucket additions of the [d]_1 MSM)", "bound": "imad",
            "achieved": algo_imad / (bwd * 1e-3) / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
            "frac": algo_imad / (bwd * 1e-3) / imad_peak,
            "executed_frac": adds * 5 * 300 / (bwd * 1e-3) / imad_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full (profiles/r1_e_summary.md);
            # valid for the 1-GPU 2^20 workload the capture was taken on, null otherwise
            "traffic": TRAFFIC_BWD if (world == 1 and log_n == 20) else None,
            "algorithmic_bytes": adds * (2 * 96 + 48 + 96),
            "kernel_ms": bwd, "kernel_share_of_step": bwd / value if value else None,
            "peak_source": "measured live: dependency-free IMAD.WIDE.U32 issue rate (pm_bench_imad_peak); "
                           "north_star names the INT32 IMAD pipe as the roofline for MSM / field multiplication",
            "algorithmic": "%d bucket additions (sorted entries / 2) x 10 Fq-modmul x 300 IMAD (SURVEY.md 8d: one XYZZ mixed "
                           "addition each); the kernel executes 5 products per addition" % adds,
            "stage": stage,
        }
        # the same launch seen from the HBM side (the contract's other roofline): algorithmic bytes and measured DRAM
        # traffic over the live kernel time, against the measured copy peak of MEASURED_PEAKS.json
        try:
            _hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
        except Exception:
            _hbm = 6650.0
        roofline["hbm_view"] = {
            "achieved": roofline["algorithmic_bytes"] / (bwd * 1e-3) / 1e9, "peak": _hbm, "unit": "GB/s",
            "frac": roofline["algorithmic_bytes"] / (bwd * 1e-3) / 1e9 / _hbm,
            "traffic_frac": (roofline["traffic"] / (bwd * 1e-3) / 1e9 / _hbm) if roofline["traffic"] else None,
        }
    else:
        roofline = {
            "kernel": "k_accumulate (bucket accumulation of the [d]_1 MSM)", "bound": "imad",
            "achieved": stage["achieved"], "peak": imad_peak / 1e12, "unit": "TIMAD/s", "frac": stage["frac"],
            "traffic": None, "kernel_ms": acc, "kernel_share_of_step": stage["share_of_step"],
            "peak_source": "measured live: dependency-free IMAD.WIDE.U32 issue rate (pm_bench_imad_peak)",
            "algorithmic": stage["algorithmic"],
        }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    check(lib.pm_bench_ntt(log_n + 1, 0, 5, C.byref(d)))
    ntt_ms = d.value
    ntt_gbs = 64.0 * (2 * n) / (ntt_ms * 1e-3) / 1e9
    roofline_hbm = {"kernel": "Fr NTT 2^%d (column + row pass)" % (log_n + 1), "bound": "hbm", "achieved": ntt_gbs,
                    "peak": hbm_peak, "unit": "GB/s", "frac": ntt_gbs / hbm_peak,
                    "traffic": 171.6e6 if log_n == 20 else None,
                    "note": "a 255-bit-field NTT is multiplier-bound: fmaheavy pipe 65-68 % active at 4 % of HBM peak (profiles/r1_c_summary.md)",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "gelem_per_s": (2 * n) / (ntt_ms * 1e-3) / 1e9}
    acc_d = C.c_double()
    check(lib.pm_bench_msm(1 << 22, 0, 2, C.byref(d), C.byref(acc_d)))
    msm_mpts = (1 << 22) / (d.value * 1e-3) / 1e6

    cpu = None
    if not args.no_cpu_baseline and world == 1:       # the CPU arm is timed on rank 0 at N = 1 only
        try:
            cpu = cpu_baseline(log_n)
        except Exception as e:  # the oracle library is optional on the product path
            cpu = {"value": None, "unit": "ms", "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}

    h2d = len(inst_wire) + len(wit_wire) + 64 + 64 + 64
    d2h = 2 * 96 + 4 + 32 + 96 + 4
    out = {
        "metric": METRIC_FMT.format(log_n=log_n), "value": value, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "S-mimc(2^%d): MiMC chain of tests/mimc.rs with %d rounds (SURVEY.md 8d), "
                               "Polymath prove, proving key resident on the device" % (log_n, n // 4 - 1),
                   "log_n": log_n, "msm_points_per_prove": 14 * n + 31,
                   "l2": "inputs larger than L2: key %.2f GB, MSM workspace > 0.7 GB per launch" % ((14 * n + 29) * 96 / 1e9),
                   "parallelism": "1 GPU" if world == 1 else "MSM split by point range over %d GPUs, all-gather of partial sums (%s)" % (
                       world, "ncclAllGather on device buffers inside the phases" if getattr(prover, "collective", "") == "nccl"
                       else "torch.distributed callback")},
        "e2e": {"value": e2e_value, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks, "lmem_resize_to_max": lmem_flag,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
        "phase_ms": {"phase1": phase_ms[0], "phase2": phase_ms[1], "phase3": phase_ms[2]},
        "wall_ms_per_step": wall_resident / args.steps, "setup_s": setup_s,
        "kernel_sweep": dict({"g1_msm_mpts_per_s_2p22": msm_mpts, "fr_ntt_gelem_per_s_2p%d" % (log_n + 1): roofline_hbm["gelem_per_s"]},
                             **sweep_dist),
        "proof_hex": last_proof.hex(),
        # acceptance of the last timed proof by the host verifier (pm_polymath_verify = verifier.rs:19-62: Merlin
        # challenges recomputed, two-pairing check against the key's [x]_2, [z]_2); outside the timed region
        "proof_verified": bool(Polymath.verify(vk_bytes, instance[1:], last_proof)),
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
