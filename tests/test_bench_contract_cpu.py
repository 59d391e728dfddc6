"""bench.py contract (task brief, section 4): the JSON line of our arm as recorded on the B200 (profiles/) carries every
key the driver reads, and the reference arm runs on the CPU and prints its line.  No GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def test_recorded_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r2_ak.json")))      # the round's final capture (tools/gpu_ak.sh)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"].startswith("prove_ms_2p20") and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert _num(e["value"]) and e["unit"] == "ms" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                      # measured separately, through host buffers
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1
    assert r["traffic"] >= r["algorithmic_bytes"] > 0    # measured DRAM traffic is above the algorithmic bytes
    c = d["cpu_baseline"]
    assert _num(c["value"]) and c["unit"] == "ms" and c["cores"] >= 1 and c["kind"] in ("port", "reference") and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["proof_verified"] is True and len(d["proof_hex"]) == 352
    # round 2: stage-level roofline with the executed fraction beside it, measured traffic with its source, the golden
    # proof check, a complete CPU prove whose bytes equal the device's, the NTT view, the sweep incl. skewed inputs
    assert 0 < r["executed_frac"] < r["frac"] and r["traffic_source"] and r["kernel_share_of_step"] > 0.5
    assert d["proof_check"]["matches_golden"] is True and d["proof_check"]["verified"] is True
    assert c["proof_matches_device"] is True and set(c["split_ms"]) >= {"sap", "ntt", "msm_phase1", "opening", "msm_d"}
    h = d["roofline_hbm"]
    assert h["bound"] == "hbm" and 0 < h["frac"] < 1 and h["traffic"] >= 64 * (1 << 21) and 0 < h["fr_mul_view"]["frac"] < 1
    assert {"g1_msm_mpts_per_s_2p22", "fr_ntt_gelem_per_s_2p21", "g1_msm_mpts_per_s_2p22_skewed"} <= set(d["kernel_sweep"])
    assert d["setup"]["fixed_base_roofline"]["bound"] == "imad"


def test_reference_arm_runs_on_the_cpu_and_prints_its_line():
    """The reference arm is ONE complete CPU prove per step (oracle/fast.py); here at 2^10 over a key-shaped set of
    arbitrary points (no GPU in this container to run the setup), on the box at --log-n 20 with the real key."""
    env = dict(os.environ, PM_REF_SYNTHETIC_KEY="1", OMP_NUM_THREADS="1")      # as under torchrun
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--log-n", "10"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "prove_ms_2p10_sap_constraints" and d["unit"] == "ms"
    assert "complete prove" in d["cpu_baseline"]["sample"] and set(d["cpu_baseline"]["split_ms_last"]) >= {"ntt", "msm_d"}
    assert len(d["cpu_baseline"]["proof_hex"]) == 352 and d["config"]["log_n"] == 10
    assert d["higher_is_better"] is False and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    # the port asks for every core of the process although OMP_NUM_THREADS=1 was exported
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    assert d["cpu_baseline"]["cores"] == avail
    # other ranks of a torchrun launch print nothing and exit 0
    out2 = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=600)
    assert out2.returncode == 0 and out2.stdout.strip() == ""
