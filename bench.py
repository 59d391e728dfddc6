#!/usr/bin/env python
"""bench.py — Polymath prove latency on synthetic SAP circuits (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--log-n 20] [--workload mimc|dummy]

One "step" = one `Polymath::prove` of the S-mimc(2^log_n) circuit (`--workload dummy` = the S-dummy
circuit of benches/bench.rs:38-61: infinity bases, one hot bucket) with the proving key resident on the device.
`value` times K proves whose witness is already resident in HBM (CUDA events on the library's stream, host
transcript round-trips included); `e2e` times the same K proves through the public host-buffer call
(`pm_polymath_prove`: H2D of instance + witness from pinned memory, D2H of the proof pieces, inside the timed region).
One JSON line on stdout.

Beyond the timing the line carries its own evidence:
  * `proof_check`: a proof drawn from a FIXED blinding seed after the timed legs, compared with the committed golden
    proof of the same workload (tests/golden/bench_proofs.json, produced by the one-GPU path and accepted by the
    oracle's pairing check in tests/test_golden_gpu.py).  Every N must reproduce those bytes: a mismatch aborts the run.
  * `cpu_baseline` (N = 1): ONE complete CPU prove of the same circuit, same key (exported from the device), same
    witness, same blinding — oracle/fast.py, C++/OpenMP on all host cores — timed, and its 176 proof bytes compared
    with the device's (`proof_matches_device`).
  * `leg_2p24` (N >= 2): setup + proves of S-mimc(2^24) sharded over the N GPUs (BASELINE.json configs[3]).

`--impl reference` times that CPU prove K times (each step one complete prove at the arm's --log-n) with a key built
by `python -m polymath_b200.keydump` in a subprocess when a GPU is present (the arm's own process never loads the
CUDA library), else over arbitrary curve points of the same shape.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC_FMT = "prove_ms_2p{log_n}_sap_constraints"
CHECK_SEED = 0xBE7C4          # blinding seed of the `proof_check` proof
GOLDEN_PROOFS = os.path.join(ROOT, "tests", "golden", "bench_proofs.json")
# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels of the [d]_1 bucket-accumulation stage and of
# the Fr NTT, summed from the `ncu --set full` captures named below; valid for the 1-GPU S-mimc(2^20) workload only.
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_traffic.json")


def ark_window(n):
    if n < 32:
        return 3
    return ((n - 1).bit_length()) * 69 // 100 + 2


# --------------------------------------------------------------------------------------------
# CPU prove (oracle/fast.py) — the cpu_baseline leg and the reference arm
# --------------------------------------------------------------------------------------------
def _csr_from_r1cs(r1cs, k):
    """Matrix k (0 = A, 1 = B, 2 = C) of a polymath_b200.api.R1CS as oracle.fast.Csr (zero-copy views of its buffers)."""
    import numpy as np
    from oracle import fast
    rp = np.frombuffer(r1cs._keep[3 * k], dtype=np.uint64)
    nnz = int(rp[-1])
    col = np.frombuffer(r1cs._keep[3 * k + 1], dtype=np.uint32)[:nnz]
    val = np.frombuffer(r1cs._keep[3 * k + 2], dtype=np.uint8)[:nnz * 32].view(np.uint64)
    return fast.Csr(rp, col, val)


class CpuProver:
    """S-mimc / S-dummy prove on the host cores with oracle/fast.py."""

    def __init__(self, r1cs, instance, witness, key, key_kind):
        import numpy as np
        from oracle import cpp, fast
        from oracle.poly import Domain
        from polymath_b200 import codec               # wire conversion of the inputs only
        self.fast = fast
        self.cores = cpp.use_all_cores()              # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses the host's cores
        self.m0, self.mw, self.nr = r1cs.m0, r1cs.mw, r1cs.nr
        rows = 2 * (self.m0 + self.nr)
        self.n = 1 << (rows - 1).bit_length()
        self.omega = Domain(self.n).group_gen
        self.a, self.b, self.c = (_csr_from_r1cs(r1cs, k) for k in range(3))
        self.x = np.frombuffer(codec.frs_to_wire(instance), dtype=np.uint64).reshape(-1, 4)
        self.w = np.frombuffer(codec.frs_to_wire(witness), dtype=np.uint64).reshape(-1, 4)
        self.key, self.key_kind = key, key_kind

    def prove(self, ra, timings=None):
        return self.fast.prove(self.key, self.a, self.b, self.c, self.m0, self.mw, self.nr, self.n, self.n + 3, self.omega,
                               self.x, self.w, ra, timings=timings)


def _key_from_dir(path):
    import numpy as np
    from oracle import fast
    return {name: np.fromfile(os.path.join(path, name + ".bin"), dtype=np.uint64).reshape(-1, 12) for name in fast.KEY_NAMES}


def _key_from_device(pk):
    import numpy as np
    from oracle import fast
    from polymath_b200 import keydump
    key = {}
    for i, name in enumerate(keydump.KEY_NAMES):
        assert name == fast.KEY_NAMES[i]
        buf, ln = keydump.export_raw(pk, i)
        key[name] = np.frombuffer(buf, dtype=np.uint64)[:ln * 12].reshape(-1, 12)
    return key


def _check_ra():
    """The two blinding coefficients StdRng::seed_from_u64(CHECK_SEED) yields (prover.rs:110), as ints — from the ORACLE's
    restatement of the generator, so the CPU arm needs nothing from the product library."""
    from oracle.rng import StdRng as ORng, fr_rand
    rng = ORng.seed_from_u64(CHECK_SEED)
    return [fr_rand(rng), fr_rand(rng)]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from polymath_b200 import keydump                  # circuit generation only (no CUDA call, library not loaded)
    from oracle.rng import StdRng as ORng, fr_rand as o_fr_rand

    class _OracleRng:                                  # the oracle's StdRng behind the `fr_rand()` interface of the circuits
        def __init__(self, seed):
            self.r = ORng.seed_from_u64(seed)

        def fr_rand(self):
            return o_fr_rand(self.r)

    log_n = args.log_n
    r1cs, instance, witness, _ = keydump.build_workload(args.workload, log_n, 1, rng=_OracleRng(1))
    key, key_kind, vk_bytes = None, None, None
    tmp = tempfile.mkdtemp(prefix="pm_refkey_")
    if not os.environ.get("PM_REF_SYNTHETIC_KEY"):
        try:
            res = subprocess.run([sys.executable, "-m", "polymath_b200.keydump", "--log-n", str(log_n), "--seed", "1",
                                  "--workload", args.workload, "--out", tmp], cwd=ROOT, capture_output=True, text=True, timeout=1200)
            if res.returncode == 0:
                key = _key_from_dir(tmp)
                vk_bytes = open(os.path.join(tmp, "vk.bin"), "rb").read()
                key_kind = "real key: Polymath::setup of the same circuit and seed, built on the GPU by a subprocess (polymath_b200.keydump)"
        except Exception:
            key = None
    if key is None:
        from oracle import fast
        n = 1 << log_n
        key = fast.synthetic_key(n, r1cs.m0, 3 * r1cs.m0 + r1cs.mw + r1cs.nr)
        key_kind = "no GPU for the setup: arbitrary curve points in the shape of the key (same MSM work, proof not checkable)"
    cpu = CpuProver(r1cs, instance, witness, key, key_kind)
    ra = _check_ra()
    # warm-up: complete proves as well (thread pool, page faults of the 1.4 GB key)
    for _ in range(args.warmup):
        cpu.prove(ra)
    vals, tm = [], {}
    proof = None
    t_all = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        proof = cpu.prove(ra, timings=tm)
        vals.append((time.perf_counter() - t0) * 1e3)
    wall = time.perf_counter() - t_all
    value = sum(vals) / len(vals)
    proof_hex = proof.serialize_compressed().hex()
    golden = _golden_proof(args.workload, log_n)
    sample = ("%d complete proves of %s(2^%d) — the five MSMs, four iFFTs, fft/square/ifft(2n), Horner, sparse assembly and "
              "division by (X - x1) of prover.rs:66-237 in oracle/cpu_ref.cpp (C++/OpenMP, portable u128 field arithmetic: "
              "slower per core than arkworks' assembly-free Montgomery code by an estimated 1.5-2x), transcript in Python; %s"
              % (args.steps, "S-" + args.workload, log_n, key_kind))
    cpu_line = {"value": value, "unit": "ms", "cores": cpu.cores, "kind": "port", "sample": sample,
                "split_ms_last": {k: v * 1e3 / args.steps for k, v in tm.items()},
                "proof_hex": proof_hex,
                "proof_matches_golden": (proof_hex == golden) if (golden and vk_bytes is not None) else None}
    out = {
        "impl": "reference", "metric": METRIC_FMT.format(log_n=log_n), "value": value, "unit": "ms",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": _workload_config(args.workload, log_n),
        "cpu_baseline": cpu_line,
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def _workload_config(workload, log_n):
    n = 1 << log_n
    if workload == "dummy":
        desc = ("S-dummy(2^%d): benches/bench.rs:38-61 with %d variables / constraints (a*b = c repeated, unused witness copies "
                "-> infinity bases, every y scalar equal -> one hot bucket per window), Polymath prove" % (log_n, n // 2 - 2))
    else:
        desc = ("S-mimc(2^%d): MiMC chain of tests/mimc.rs with %d rounds (SURVEY.md 8d), Polymath prove, proving key resident "
                "on the device" % (log_n, n // 4 - 1))
    return {"workload": desc, "log_n": log_n, "msm_points_per_prove": 14 * n + 31}


def _golden_proof(workload, log_n):
    try:
        return json.load(open(GOLDEN_PROOFS)).get("%s_2p%d_seed1_check" % (workload, log_n))
    except Exception:
        return None


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
def _max_over_ranks(v, world):
    if world == 1:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _timed_proves(lib, check, fn, steps, barrier, world):
    """K calls of `fn` between two barriers, CUDA events on the library's stream, max over ranks -> ms per call."""
    ms = C.c_double()
    barrier()
    check(lib.pm_timer_start())
    last = None
    for _ in range(steps):
        last = fn()
    check(lib.pm_timer_stop(C.byref(ms)))
    barrier()
    return _max_over_ranks(ms.value, world) / steps, last


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

    from polymath_b200 import keydump
    from polymath_b200.api import Polymath, StdRng, _lib
    from polymath_b200 import codec
    from polymath_b200.lib import check
    lib = _lib()
    dp = C.POINTER(C.c_double)
    lib.pm_bench_fixed_base.argtypes = [C.c_size_t, C.c_int, dp]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Workload:
        """Circuit + resident key (+ NCCL communicator when sharded) + pinned host buffers of one size."""

        def __init__(self, log_n):
            self.log_n, self.n = log_n, 1 << log_n
            t0 = time.perf_counter()
            self.r1cs, self.instance, self.witness, self.rng = keydump.build_workload(args.workload, log_n, 1)
            self.synth_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            if world > 1:
                from polymath_b200 import sharded
                self.prover = sharded.ShardedProver(self.r1cs, self.rng, rank, world)
                self.pk, self.vk_bytes = self.prover.pk, self.prover.vk_bytes
            else:
                self.pk, self.vk_bytes = Polymath.setup(self.r1cs, self.rng)
                self.prover = None
            torch.cuda.synchronize()
            self.setup_s = time.perf_counter() - t0
            self.inst_wire = codec.frs_to_wire(self.instance)
            self.wit_wire = codec.frs_to_wire(self.witness)
            # pinned host buffers for the e2e leg
            self.pin_inst = torch.empty(len(self.inst_wire), dtype=torch.uint8).pin_memory()
            self.pin_wit = torch.empty(len(self.wit_wire), dtype=torch.uint8).pin_memory()
            self.pin_inst.copy_(torch.frombuffer(bytearray(self.inst_wire), dtype=torch.uint8))
            self.pin_wit.copy_(torch.frombuffer(bytearray(self.wit_wire), dtype=torch.uint8))
            self.p_inst = C.c_char_p(self.pin_inst.data_ptr())
            self.p_wit = C.c_char_p(self.pin_wit.data_ptr())
            self.proof = C.create_string_buffer(176)
            if self.prover is not None:
                self.prover.set_assignment(self.p_inst, self.p_wit)
            else:
                check(lib.pm_ctx_set_assignment(self.pk._h, self.p_inst, self.p_wit))

        def prove_resident(self, rng=None):
            rng = rng or self.rng
            if self.prover is not None:
                return self.prover.prove_resident(self.p_inst, rng)
            check(lib.pm_polymath_prove_resident(self.pk._h, self.p_inst, rng._h, self.proof))
            return self.proof.raw

        def prove_e2e(self):
            if self.prover is not None:
                return self.prover.prove(self.p_inst, self.p_wit, self.rng)
            check(lib.pm_polymath_prove(self.pk._h, self.p_inst, self.p_wit, self.rng._h, self.proof))
            return self.proof.raw

        def check_proof(self):
            """Proof with the blinding of StdRng::seed_from_u64(CHECK_SEED): the same bytes on every N, pinned by the golden."""
            pr = self.prove_resident(StdRng.seed_from_u64(CHECK_SEED))
            hexs = pr.hex()
            golden = _golden_proof(args.workload, self.log_n)
            ok = bool(Polymath.verify(self.vk_bytes, self.instance[1:], pr))
            if golden is not None and hexs != golden:
                raise SystemExit("proof_check of %s(2^%d) on %d GPU(s) differs from tests/golden/bench_proofs.json:\n got  %s\n want %s"
                                 % (args.workload, self.log_n, world, hexs, golden))
            if not ok:
                raise SystemExit("the host verifier rejects the proof_check proof of 2^%d" % self.log_n)
            return {"seed": CHECK_SEED, "proof_hex": hexs, "verified": ok,
                    "matches_golden": (hexs == golden) if golden is not None else None,
                    "golden": "tests/golden/bench_proofs.json (one-GPU proof, accepted by the oracle pairing in tests/test_golden_gpu.py)"}

        def close(self):
            self.pk.close()

    log_n = args.log_n
    n = 1 << log_n
    wl = Workload(log_n)

    # resident leg ------------------------------------------------------------------------
    # the clock sampler is started BEFORE the warm-up: its start-up (NVML initialisation) contends with CUDA calls for the
    # driver, which would otherwise fall into the timed region of a short run; only samples taken after the mark count
    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("PM_BENCH_NO_CLOCKS")) else None
    if sampler:
        sampler.start()
        sampler.wait_first(5.0)
    for _ in range(args.warmup):
        wl.prove_resident()
    check(lib.pm_bench_set_kernel_timing(1))
    launches0 = lib.pm_kernel_launches()
    barrier()
    if sampler:
        sampler.mark()
    check(lib.pm_timer_start())
    wall0 = time.perf_counter()
    acc_ms, msm_geom = [], (0, 0)
    last_proof = None
    for _ in range(args.steps):
        last_proof = wl.prove_resident()
        km = (C.c_double * 4)()
        check(lib.pm_bench_last_msm(km))        # the [d]_1 MSM is the last one of a prove
        acc_ms.append(km[0])
        msm_geom = (int(km[2]), int(km[3]))
    ms = C.c_double()
    check(lib.pm_timer_stop(C.byref(ms)))
    barrier()
    wall_resident = (time.perf_counter() - wall0) * 1e3
    launches = lib.pm_kernel_launches() - launches0
    check(lib.pm_bench_set_kernel_timing(0))
    value = _max_over_ranks(ms.value, world) / args.steps
    phase_ms = wl.pk.phase_ms()

    # e2e leg -----------------------------------------------------------------------------
    wl.prove_e2e()
    e2e_value, _ = _timed_proves(lib, check, wl.prove_e2e, args.steps, barrier, world)
    clocks = sampler.stop() if sampler else None          # rank 0, sampled over both timed legs (resident + e2e)
    proof_check = wl.check_proof()

    # standalone kernel sweep across the ranks (BASELINE.json configs[4]): sharded Fr NTT with one NCCL all-to-all,
    # G1 MSM split by point range (every rank a 1/world share; the 192-byte partial sums are not timed)
    sweep_dist = {}
    d = C.c_double()
    if world > 1:
        from polymath_b200 import sharded as _sh
        for ntt_log in (20, 24, 26):
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
            t_ntt = _max_over_ranks(e0.elapsed_time(e1) / 5, world)
            sweep_dist["fr_ntt_gelem_per_s_2p%d_sharded" % ntt_log] = (1 << ntt_log) / (t_ntt * 1e-3) / 1e9
            del sn
            torch.cuda.empty_cache()
        dacc = C.c_double()
        for msm_log in (20, 24, 26):
            check(lib.pm_bench_msm((1 << msm_log) // world, 0, 2, C.byref(d), C.byref(dacc)))
            sweep_dist["g1_msm_mpts_per_s_2p%d_sharded" % msm_log] = (1 << msm_log) / (_max_over_ranks(d.value, world) * 1e-3) / 1e6
        check(lib.pm_bench_set_msm_skew(1))
        check(lib.pm_bench_msm((1 << 24) // world, 0, 2, C.byref(d), C.byref(dacc)))
        check(lib.pm_bench_set_msm_skew(0))
        sweep_dist["g1_msm_mpts_per_s_2p24_sharded_skewed"] = (1 << 24) / (_max_over_ranks(d.value, world) * 1e-3) / 1e6

    # the 2^24 leg (BASELINE.json configs[3]): setup + proves sharded over the N GPUs ------------------------------------
    leg24 = None
    if world > 1 and log_n == 20 and args.workload == "mimc" and not args.no_2p24:
        wl.close()
        torch.cuda.empty_cache()
        big = Workload(24)
        big.prove_resident()
        t24, _ = _timed_proves(lib, check, big.prove_resident, 3, barrier, world)
        ph24 = big.pk.phase_ms()
        big.prove_e2e()
        e24, _ = _timed_proves(lib, check, big.prove_e2e, 2, barrier, world)
        chk24 = big.check_proof()
        leg24 = {"metric": METRIC_FMT.format(log_n=24), "value": t24, "unit": "ms", "steps": 3, "warmup": 1,
                 "e2e_ms": e24, "phase_ms": {"phase1": ph24[0], "phase2": ph24[1], "phase3": ph24[2]},
                 "setup_s": big.setup_s, "circuit_synthesis_s": big.synth_s, "proof_check": chk24,
                 "parallelism": "MSM bases split by point range over %d GPUs, sharded NTTs (one NCCL all-to-all each), "
                                "all-gather of partial sums inside the phases" % world}
        big.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant stage ----------------------------------------------------------------------------------
    # The [d]_1 MSM (10n + 22 points) is 2/3 of a prove and its bucket accumulation — R batched-affine pair rounds (the
    # lower and the upper half of the buckets on two streams) + a short XYZZ walk — is the dominant stage.  Stage-level
    # figure of SURVEY.md 8(d): N points x W windows (arkworks' window rule) x one XYZZ mixed addition = 10 Fq-modmul x
    # 300 IMAD, over the stage's time measured live with CUDA events, against the IMAD.WIDE issue peak measured live.
    check(lib.pm_bench_imad_peak(C.byref(d)))
    imad_peak = d.value
    d_points = (10 * n + 22 + world - 1) // world
    c_ark = ark_window(10 * n + 22)
    w_ark = (255 + c_ark - 1) // c_ark
    med = lambda v: sorted(v)[len(v) // 2] if v else 0.0
    acc = med(acc_ms)
    rounds, entries = msm_geom
    stage_imad = d_points * w_ark * 10 * 300
    # products the stage actually executes: 6 per slot pair of every round (1 forward + 5 backward), 10 per point the
    # walk still adds; 276 wide IMADs per product as compiled (cuobjdump: IMAD.WIDE.U32 of Fq operator*)
    executed = sum((entries >> (r + 1)) * 6 for r in range(rounds)) + (entries >> rounds) * 10
    traffic = None
    traffic_src = None
    try:
        tj = json.load(open(TRAFFIC_FILE))
        if world == 1 and log_n == 20 and args.workload == "mimc":
            traffic, traffic_src = tj["d_msm_accumulation_stage_bytes"], tj["source"]
    except Exception:
        tj = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    algorithmic_bytes = d_points * 128          # 96-byte base + 32-byte scalar, read once
    roofline = {
        "kernel": ("bucket accumulation of the [d]_1 MSM: k_pairs_forward / k_invert_* / k_pairs_backward x %d rounds (two bucket "
                   "halves on two streams) + k_accumulate_rounds" % rounds) if rounds else "k_accumulate (XYZZ walk of the [d]_1 MSM)",
        "bound": "imad", "achieved": stage_imad / (acc * 1e-3) / 1e12 if acc > 0 else None, "peak": imad_peak / 1e12,
        "unit": "TIMAD/s", "frac": stage_imad / (acc * 1e-3) / imad_peak if acc > 0 else None,
        "executed_frac": executed * 276 / (acc * 1e-3) / imad_peak if acc > 0 else None,
        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": algorithmic_bytes,
        "kernel_ms": acc, "kernel_share_of_step": acc / value if value else None,
        "peak_source": "measured live: dependency-free IMAD.WIDE.U32 issue rate (pm_bench_imad_peak); north_star names the "
                       "INT32 IMAD pipe as the roofline for MSM / field multiplication",
        "algorithmic": "%d points x %d windows (arkworks window rule c=%d) x 10 Fq-modmul x 300 IMAD (SURVEY.md 8d), over the whole "
                       "stage; executed_frac counts the products the kernels run (6 per batched-affine addition, 10 per walk "
                       "addition, 276 wide IMADs each)" % (d_points, w_ark, c_ark),
        "hbm_view": {"achieved": algorithmic_bytes / (acc * 1e-3) / 1e9 if acc > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                     "frac": algorithmic_bytes / (acc * 1e-3) / 1e9 / hbm_peak if acc > 0 else None,
                     "traffic_frac": (traffic / (acc * 1e-3) / 1e9 / hbm_peak) if (traffic and acc > 0) else None},
    }
    check(lib.pm_bench_ntt(log_n + 1, 0, 5, C.byref(d)))
    ntt_ms = d.value
    ntt_gbs = 64.0 * (2 * n) / (ntt_ms * 1e-3) / 1e9
    roofline_hbm = {"kernel": "Fr NTT 2^%d (column + row pass, ping-pong)" % (log_n + 1), "bound": "hbm", "achieved": ntt_gbs,
                    "peak": hbm_peak, "unit": "GB/s", "frac": ntt_gbs / hbm_peak,
                    "traffic": tj.get("ntt_2p21_bytes") if log_n == 20 else None, "traffic_source": tj.get("source") if log_n == 20 else None,
                    "note": "a 255-bit-field NTT is multiplier-bound: 4 % of HBM peak with the fmaheavy pipe ~60 % active "
                            "(profiles/r2_summary.md); ceiling from the measured Fr product rate reported beside it",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "gelem_per_s": (2 * n) / (ntt_ms * 1e-3) / 1e9}
    check(lib.pm_bench_field_mul(0, C.byref(d)))
    fr_rate = d.value
    ntt_muls = (2 * n) // 2 * (log_n + 1)
    roofline_hbm["fr_mul_view"] = {"achieved": ntt_muls / (ntt_ms * 1e-3) / 1e9, "peak": fr_rate / 1e9, "unit": "G Fr-mul/s",
                                   "frac": ntt_muls / (ntt_ms * 1e-3) / fr_rate,
                                   "algorithmic": "(N/2) log2 N butterfly products (SURVEY.md 8d) over the measured register-resident Fr product rate"}
    acc_d = C.c_double()
    check(lib.pm_bench_msm(1 << 22, 0, 2, C.byref(d), C.byref(acc_d)))
    msm_mpts = (1 << 22) / (d.value * 1e-3) / 1e6
    # the skewed inputs of SURVEY.md 8d (89 % one repeated scalar, 10 % zero, 1 % infinity bases): hot buckets
    check(lib.pm_bench_set_msm_skew(1))
    check(lib.pm_bench_msm(1 << 22, 0, 2, C.byref(d), C.byref(acc_d)))
    check(lib.pm_bench_set_msm_skew(0))
    sweep_dist["g1_msm_mpts_per_s_2p22_skewed"] = (1 << 22) / (d.value * 1e-3) / 1e6
    # setup (BASELINE.json configs[2]: "prove + setup (fixed-base batch mul)"): the generator's fixed-base batch alone
    fb_n = 1 << 22
    check(lib.pm_bench_fixed_base(fb_n, 2, C.byref(d)))
    fb_ms = d.value
    key_points = 14 * n + 29
    setup = {"setup_s": wl.setup_s, "key_points": key_points,
             "includes": "trapdoor powers, Lagrange / lcs scalars, six G1 vectors, [x]_2 [z]_2, and the fixed-base MSM tables of the "
                         "prover (%.1f GB)" % ((14 * n) * 13 * 96 / 1e9),
             "fixed_base_mpts_per_s": fb_n / (fb_ms * 1e-3) / 1e6,
             "fixed_base_roofline": {"bound": "imad", "unit": "TIMAD/s", "peak": imad_peak / 1e12,
                                     "achieved": fb_n * (22 * 10 + 40) * 300 / (fb_ms * 1e-3) / 1e12,
                                     "frac": fb_n * (22 * 10 + 40) * 300 / (fb_ms * 1e-3) / imad_peak,
                                     "algorithmic": "2^22 scalars x (22 window additions x 10 Fq-modmul + ~40 for the batched "
                                                    "normalisation) x 300 IMAD (k_fixed_base + k_batch_to_affine)"},
             "circuit_synthesis_s": wl.synth_s}

    cpu = None
    if not args.no_cpu_baseline and world == 1:       # the CPU arm is timed on rank 0 at N = 1 only
        try:
            cp = CpuProver(wl.r1cs, wl.instance, wl.witness, _key_from_device(wl.pk), "key exported from the device context")
            tm = {}
            t0 = time.perf_counter()
            cproof = cp.prove(_check_ra(), timings=tm)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            chex = cproof.serialize_compressed().hex()
            cpu = {"value": cpu_ms, "unit": "ms", "cores": cp.cores, "kind": "port",
                   "sample": "ONE complete prove of the same circuit, key (exported from the device), witness and blinding as "
                             "`proof_check`: oracle/fast.py — the five MSMs, four iFFTs, fft/square/ifft(2n), Horner, assembly and "
                             "division of prover.rs:66-237 in oracle/cpu_ref.cpp (C++/OpenMP on all host cores, portable u128 "
                             "arithmetic; arkworks' own field code is an estimated 1.5-2x faster per core)",
                   "split_ms": {k: v * 1e3 for k, v in tm.items()}, "proof_hex": chex,
                   "proof_matches_device": chex == proof_check["proof_hex"]}
            if not cpu["proof_matches_device"]:
                raise SystemExit("CPU oracle proof differs from the device proof:\n cpu    %s\n device %s" % (chex, proof_check["proof_hex"]))
        except SystemExit:
            raise
        except Exception as e:  # the oracle library is optional on the product path
            cpu = {"value": None, "unit": "ms", "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}

    h2d = len(wl.inst_wire) + len(wl.wit_wire) + 64 + 64 + 64
    d2h = 2 * 96 + 4 + 32 + 96 + 4
    cfg = _workload_config(args.workload, log_n)
    cfg["l2"] = "inputs larger than L2: key %.2f GB, MSM workspace > 0.7 GB per launch" % ((14 * n + 29) * 96 / 1e9)
    cfg["parallelism"] = "1 GPU" if world == 1 else "MSM split by point range over %d GPUs, all-gather of partial sums (%s)" % (
        world, "ncclAllGather on device buffers inside the phases" if getattr(wl.prover, "collective", "") == "nccl"
        else "torch.distributed callback")
    out = {
        "metric": METRIC_FMT.format(log_n=log_n), "value": value, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e_value, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks, "lmem_resize_to_max": lmem_flag,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
        "phase_ms": {"phase1": phase_ms[0], "phase2": phase_ms[1], "phase3": phase_ms[2]},
        "wall_ms_per_step": wall_resident / args.steps, "setup_s": wl.setup_s, "setup": setup,
        "kernel_sweep": dict({"g1_msm_mpts_per_s_2p22": msm_mpts, "fr_ntt_gelem_per_s_2p%d" % (log_n + 1): roofline_hbm["gelem_per_s"]},
                             **sweep_dist),
        "proof_hex": last_proof.hex(),
        # acceptance of the last timed proof by the host verifier (pm_polymath_verify = verifier.rs:19-62: Merlin
        # challenges recomputed, two-pairing check against the key's [x]_2, [z]_2); outside the timed region
        "proof_verified": bool(Polymath.verify(wl.vk_bytes, wl.instance[1:], last_proof)),
        "proof_check": proof_check,
        "leg_2p24": leg24,
    }
    print(json.dumps(out), flush=True)
    if args.write_golden:
        os.makedirs(os.path.dirname(args.write_golden) or ".", exist_ok=True)
        try:
            g = json.load(open(args.write_golden))
        except Exception:
            g = {}
        g["%s_2p%d_seed1_check" % (args.workload, log_n)] = proof_check["proof_hex"]
        json.dump(g, open(args.write_golden, "w"), indent=1, sort_keys=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--workload", default="mimc", choices=["mimc", "dummy"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-2p24", action="store_true", help="skip the S-mimc(2^24) leg of a multi-GPU run")
    ap.add_argument("--write-golden", default=None, help="merge this run's proof_check into the given JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
