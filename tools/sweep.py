"""Standalone kernel sweep (BASELINE.json configs[4]): IMAD peak, field-mul rates, Fr NTT and G1 MSM
throughput on resident synthetic data.  Prints one JSON object per line."""
import argparse
import ctypes as C
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polymath_b200.lib import require_device, check  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntt", default="16,18,20,21,22,24")
    ap.add_argument("--msm", default="16,18,20,22")
    ap.add_argument("--windows", default="0")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--levels", default="1", help="comma list; 0 = as many levels as windows")
    ap.add_argument("--rounds", default="-1", help="comma list of pair-round settings (-1 = automatic, 0 = XYZZ walk only)")
    ap.add_argument("--skip-basics", action="store_true")
    ap.add_argument("--coset", action="store_true", help="also time the coset variants of the NTT (scaling pass included)")
    ap.add_argument("--skew", action="store_true", help="MSM inputs of SURVEY.md 8d: 89 %% one repeated scalar, 10 %% zero, 1 %% infinity bases")
    ap.add_argument("--codec", default="", help="comma list of log2 sizes for the G1 (de)compression kernels")
    a = ap.parse_args()
    lib = require_device()
    d = C.c_double()
    imad = 1.0
    if not a.skip_basics:
        check(lib.pm_bench_imad_peak(C.byref(d)))
        imad = d.value
        print(json.dumps({"kernel": "imad_wide_peak", "mads_per_s": imad}), flush=True)
    for f, name, cost in (() if a.skip_basics else ((0, "fr_mul", 136), (1, "fq_mul", 300))):
        check(lib.pm_bench_field_mul(f, C.byref(d)))
        print(json.dumps({"kernel": name, "muls_per_s": d.value, "imad_equiv_per_s": d.value * cost,
                          "frac_of_imad_peak": d.value * cost / imad}), flush=True)
    for lg in [int(x) for x in a.ntt.split(",") if x]:
        for inv in ((0, 1, 2, 3) if a.coset else (0, 1)):
            check(lib.pm_bench_ntt(lg, inv, a.iters, C.byref(d)))
            n = 1 << lg
            print(json.dumps({"kernel": "ntt_fr", "log_n": lg, "inverse": inv & 1, "coset": bool(inv & 2), "ms": d.value,
                              "gelem_per_s": n / d.value / 1e6, "algo_gb_per_s": 64 * n / d.value / 1e6,
                              "butterfly_muls_per_s": (n / 2) * lg / (d.value * 1e-3)}), flush=True)
    for lg in [int(x) for x in a.codec.split(",") if x]:
        check(lib.pm_bench_g1_codec(1 << lg, C.byref(d), C.byref(acc := C.c_double())))
        print(json.dumps({"kernel": "g1_decompress", "log_n": lg, "ms": d.value, "mpts_per_s": (1 << lg) / d.value / 1e3,
                          "compress_ms": acc.value}), flush=True)
    acc = C.c_double()
    check(lib.pm_bench_set_msm_skew(1 if a.skew else 0))
    for lg in [float(x) for x in a.msm.split(",") if x]:
        for w in [int(x) for x in a.windows.split(",")]:
            n = int(round(2 ** lg))
            for lv in [int(x) for x in a.levels.split(",")]:
                if lv == 0:
                    lv = (256 + w - 1) // w
                if lv > 1 and w == 0:
                    continue
                for rd in [int(x) for x in a.rounds.split(",")]:
                    check(lib.pm_msm_set_tuning(rd))
                    check(lib.pm_bench_msm_levels(n, w, lv, a.iters, C.byref(d), C.byref(acc)))
                    print(json.dumps({"kernel": "msm_g1", "log_n": lg, "window": w, "levels": lv, "rounds": rd, "skewed": bool(a.skew), "ms": d.value,
                                      "ms_accumulate": acc.value, "mpts_per_s": n / d.value / 1e3}), flush=True)
                check(lib.pm_msm_set_tuning(-1))
    check(lib.pm_bench_set_msm_skew(0))


if __name__ == "__main__":
    main()
