#!/usr/bin/env python3
"""How much do the reference's own outputs move from run to run (float atomics), and how far is the engine from each run?
Runs oracle/_ref/ref_harness parity K times on this GPU and evaluates every quantity tests/test_reference_gpu.py asserts on:
max over the runs of |engine - reference| next to the test bound, and the reference's own spread over the runs.
    python tools/ref_spread.py [K=12] > gpurun_out/ref_spread.json"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hw1f_b200 as hw  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
SEED, N = 20251018, 1 << 20
K = int(sys.argv[1]) if len(sys.argv) > 1 else 12

runs = []
with tempfile.TemporaryDirectory() as d:
    for k in range(K):
        out = os.path.join(d, f"p{k}.json")
        subprocess.run([HARNESS, "parity", str(SEED), out], check=True, stdout=subprocess.DEVNULL, cwd=d, timeout=600)
        runs.append(json.load(open(out)))

report = {"runs": K, "modes": {}}
for mode in ("decomposed", "reference_order"):
    eng = hw.Engine(device=0)
    eng.set_mode(hw._ffi.MODE_DECOMPOSED if mode == "decomposed" else hw._ffi.MODE_REFERENCE_ORDER)
    mine = eng.bond_curve(hw.Rng(SEED, N))
    n_steps = eng.steps_to(5.0)
    rows = {}

    def put(name, val):
        rows.setdefault(name, []).append(float(val))

    for r in runs:
        P, f = np.array(r["P"], np.float32), np.array(r["f"], np.float32)
        put("P_all_rel", np.abs(mine["P"] / P - 1).max())
        put("P_T>=2_rel", np.abs(mine["P"][20:] / P[20:] - 1).max())
        put("f_all_abs", np.abs(mine["f"] - f).max())
        put("f_T>=2_abs", np.abs(mine["f"][20:] - f[20:]).max())
        th = eng.theta_calibrate(mine["f"])["theta_rec"]
        th_ref = np.array(r["theta_rec"], np.float32)
        put("theta_e2e_T>=2_abs", np.abs(th[20:] - th_ref[20:]).max())
        put("theta_e2e_all_abs", np.abs(th - th_ref).max())
        z = eng.zbc_cv(hw.Rng(SEED + 54321, N), P, f, n_steps_S1=n_steps)
        put("zbc_moments_rel", np.abs(np.array(z["mom"]) / np.array(r["zbc_moments"]) - 1).max())
        put("zbc_mean_X_rel", abs(z["mean_X"] / r["zbc_mean_X"] - 1))
        put("zbc_price_cv_rel", abs(z["price_cv"] / r["zbc_price_cv"] - 1))
        put("zbc_beta_rel", abs(z["beta"] / r["zbc_beta"] - 1))
        put("zbc_corr_rel", abs(z["corr"] / r["zbc_corr"] - 1))
        put("zbc_beta_in_se", abs(z["beta"] - r["zbc_beta"]) / z["beta_se"])
        put("zbc_corr_in_se", abs(z["corr"] - r["zbc_corr"]) / z["corr_se"])
        put("zbc_price_in_se", abs(z["price_cv"] - r["zbc_price_cv"]) / z["se_cv"])
        v = eng.vega(hw.Rng(SEED, N), P, f, n_steps_S1=n_steps)
        put("vega_pathwise_rel", abs(v["vega_pathwise"] / r["vega_pathwise"] - 1))
        put("vega_pathwise_in_se", abs(v["vega_pathwise"] - r["vega_pathwise"]) / v["vega_pathwise_se"])
        put("vega_fd_abs", abs(v["vega_fd"] - r["vega_fd"]))
        put("vega_fd_recal_abs", abs(v["vega_fd_recal"] - r["vega_fd_recal"]))
    report["modes"][mode] = {k: {"max": max(v), "median": float(np.median(v))} for k, v in rows.items()}
    eng.close()

# the reference against itself: spread of each scalar over the runs
spread = {}
for key in ("zbc_mean_X", "zbc_price_cv", "zbc_beta", "zbc_corr", "vega_pathwise", "vega_fd", "vega_fd_recal"):
    vals = np.array([r[key] for r in runs], np.float64)
    spread[key] = {"max_minus_min": float(vals.max() - vals.min()), "rel": float((vals.max() - vals.min()) / abs(vals.mean()))}
Ps = np.array([r["P"] for r in runs], np.float64)
fs = np.array([r["f"] for r in runs], np.float64)
spread["P_rel"] = float(((Ps.max(0) - Ps.min(0)) / Ps.mean(0)).max())
spread["f_abs"] = float((fs.max(0) - fs.min(0)).max())
report["reference_run_to_run"] = spread
print(json.dumps(report, indent=1))
