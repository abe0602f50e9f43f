#!/usr/bin/env python3
"""Measured parity margins (not a test: the tests assert the tolerances, this prints how far inside they are).

Engine (both arithmetic modes, through the C ABI) against
  * the CPU oracle (oracle/) on the same seeds at 2^14 and 2^17 subsequences, and
  * the unmodified reference kernels run on a B200 (tests/golden/ref_b200_seed20251018.json, 2^20 subsequences).
Run on the GPU box:  python tests/parity_report.py > gpurun_out/parity_report.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import hw1f_b200 as hw  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

SEED = 4242


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a / b - 1.0)))


def vs_oracle(eng, o, n):
    out = {}
    c = eng.bond_curve(hw.Rng(SEED, n))
    P, f = o.bond_curve(SEED, n)
    out["P_max_rel"] = rel(c["P"][1:], P[1:])
    out["f_max_abs"] = float(np.abs(c["f"] - f).max())
    th = eng.theta_calibrate(c["f"])
    out["theta_max_abs"] = float(np.abs(th["theta_rec"] - o.theta(c["f"])[0]).max())
    z = eng.zbc_cv(hw.Rng(SEED + 1, n), c["P"], c["f"], n_steps_S1=500)
    mom = o.zbc_moments(SEED + 1, n, c["P"], c["f"], n_steps_S1=500)
    out["zbc_moments_max_rel"] = rel(z["mom"], mom)
    alg = o.zbc_algebra(mom, 2 * n, float(c["P"][100]))
    out["zbc_price_cv_rel"] = abs(z["price_cv"] / alg["price_cv"] - 1.0)
    out["zbc_beta_rel"] = abs(z["beta"] / alg["beta"] - 1.0)
    v = eng.vega_pathwise(hw.Rng(SEED + 2, n), c["P"], c["f"], n_steps_S1=500)
    s, _ = o.vega_pathwise_sums(SEED + 2, n, c["P"], c["f"], n_steps_S1=500)
    out["vega_pathwise_rel"] = abs(v["vega_pathwise_f64"] / (s / n) - 1.0)
    return out


def vs_reference_fixture(eng):
    """tests/golden/ref_b200_seed20251018.json: outputs of the UNMODIFIED reference kernels on a B200 (oracle/_ref/
    ref_harness parity <seed>); seeds as in tests/test_reference_gpu.py"""
    path = os.path.join(ROOT, "tests", "golden", "ref_b200_seed20251018.json")
    if not os.path.exists(path):
        return None
    ref = json.load(open(path))
    n, seed = int(ref["n_paths"]), int(ref["seed"])
    P, f = np.array(ref["P"], np.float32), np.array(ref["f"], np.float32)
    c = eng.bond_curve(hw.Rng(seed, n))
    out = {"fixture": os.path.relpath(path, ROOT), "n_paths": n,
           "P_max_rel_T_ge_2": rel(c["P"][20:], P[20:]), "P_max_rel_all_T": rel(c["P"][1:], P[1:]),
           "f_max_abs_T_ge_2": float(np.abs(c["f"][20:] - f[20:]).max())}
    th = eng.theta_calibrate(f)
    out["theta_max_abs"] = float(np.abs(th["theta_rec"] - np.array(ref["theta_rec"], np.float32)).max())
    z = eng.zbc_cv(hw.Rng(seed + 54321, n), P, f, n_steps_S1=500)
    out["zbc_moments_max_rel"] = rel(z["mom"], ref["zbc_moments"])
    out["zbc_price_cv_rel"] = abs(z["price_cv"] / ref["zbc_price_cv"] - 1.0)
    out["zbc_beta_rel"] = abs(z["beta"] / ref["zbc_beta"] - 1.0)
    v = eng.vega_pathwise(hw.Rng(seed, n), P, f, n_steps_S1=500)
    out["vega_pathwise_rel"] = abs(v["vega_pathwise"] / ref["vega_pathwise"] - 1.0)
    return out


def live_reference(eng):
    """the unmodified reference kernels run TWICE on this GPU with the same seed (oracle/_ref/ref_harness parity): the
    spread between the two runs is the reference's own float-atomic reordering noise; then the engine against run 1 for
    every estimator of the north star, theta end to end (engine f -> recover_theta vs the reference's f -> its
    recover_theta) and both FD vegas included"""
    import subprocess
    import tempfile
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return None
    seed, n = 20251018, 1 << 20
    runs = []
    with tempfile.TemporaryDirectory() as td:
        for k in range(2):
            o = os.path.join(td, f"p{k}.json")
            subprocess.run([harness, "parity", str(seed), o], check=True, stdout=subprocess.DEVNULL, cwd=td, timeout=600)
            runs.append(json.load(open(o)))
    a, b = runs
    arr = lambda r, k: np.array(r[k], np.float64)
    out = {"seed": seed, "n_paths": n, "reference_vs_reference": {
        "P_max_rel": rel(arr(a, "P")[1:], arr(b, "P")[1:]), "f_max_abs": float(np.abs(arr(a, "f") - arr(b, "f")).max()),
        "theta_max_abs": float(np.abs(arr(a, "theta_rec") - arr(b, "theta_rec")).max()),
        "zbc_moments_max_rel": rel(a["zbc_moments"], b["zbc_moments"]),
        "zbc_price_cv_rel": abs(a["zbc_price_cv"] / b["zbc_price_cv"] - 1), "zbc_beta_rel": abs(a["zbc_beta"] / b["zbc_beta"] - 1),
        "zbc_corr_rel": abs(a["zbc_corr"] / b["zbc_corr"] - 1),
        "vega_pathwise_rel": abs(a["vega_pathwise"] / b["vega_pathwise"] - 1),
        "vega_fd_abs": abs(a["vega_fd"] - b["vega_fd"]), "vega_fd_recal_abs": abs(a["vega_fd_recal"] - b["vega_fd_recal"])}}
    P, f = np.array(a["P"], np.float32), np.array(a["f"], np.float32)
    c = eng.bond_curve(hw.Rng(seed, n))
    ci = eng.bond_curve_ci()
    o = Oracle()
    P_orc, f_orc = o.bond_curve(seed, n)
    e = {"P_max_rel_T_ge_2": rel(c["P"][20:], P[20:]), "P_max_rel_all_T": rel(c["P"][1:], P[1:]),
         "f_max_abs_T_ge_2": float(np.abs(c["f"][20:] - f[20:]).max()), "f_max_abs_all_T": float(np.abs(c["f"] - f).max()),
         "f_max_in_units_of_f_se": float((np.abs(c["f"] - f)[1:] / ci["f_se"][1:]).max()),
         "reference_P_vs_oracle_max_rel_all_T": rel(P[1:], P_orc[1:]), "reference_P_vs_oracle_max_rel_T_ge_2": rel(P[20:], P_orc[20:]),
         "engine_P_vs_oracle_max_rel_all_T": rel(c["P"][1:], P_orc[1:]),
         "reference_f_vs_oracle_max_abs": float(np.abs(f - f_orc).max()), "engine_f_vs_oracle_max_abs": float(np.abs(c["f"] - f_orc).max())}
    th_mine = eng.theta_calibrate(c["f"])                      # end to end: the engine's own f
    th_ref = np.array(a["theta_rec"], np.float32)
    d_th = np.abs(th_mine["theta_rec"] - th_ref)
    e["theta_end_to_end_max_abs_T_ge_2"] = float(d_th[20:].max())
    e["theta_end_to_end_max_abs_all_T"] = float(d_th.max())
    e["theta_end_to_end_max_in_units_of_theta_se"] = float((d_th[1:] / ci["theta_se"][1:]).max())
    e["theta_same_f_max_abs"] = float(np.abs(eng.theta_calibrate(f)["theta_rec"] - th_ref).max())
    z = eng.zbc_cv(hw.Rng(seed + 54321, n), P, f, n_steps_S1=500)
    e["zbc_moments_max_rel"] = rel(z["mom"], a["zbc_moments"])
    e["zbc_price_cv_rel"] = abs(z["price_cv"] / a["zbc_price_cv"] - 1.0)
    e["zbc_beta_rel"] = abs(z["beta"] / a["zbc_beta"] - 1.0)
    e["zbc_corr_rel"] = abs(z["corr"] / a["zbc_corr"] - 1.0)
    e["zbc_beta_se_rel"] = z["beta_se"] / z["beta_f64"]
    v = eng.vega(hw.Rng(seed, n), P, f, n_steps_S1=500)
    e["vega_pathwise_rel"] = abs(v["vega_pathwise"] / a["vega_pathwise"] - 1.0)
    e["vega_fd_abs"] = abs(v["vega_fd"] - a["vega_fd"])
    e["vega_fd_rel"] = abs(v["vega_fd"] / a["vega_fd"] - 1.0)
    e["vega_fd_recal_abs"] = abs(v["vega_fd_recal"] - a["vega_fd_recal"])
    e["vega_fd_recal_rel"] = abs(v["vega_fd_recal"] / a["vega_fd_recal"] - 1.0)
    e["vega_pathwise_se_rel"] = v["vega_pathwise_se"] / v["vega_pathwise_f64"]
    out["engine_vs_reference_run1"] = e
    return out


def main():
    o = Oracle()
    rep = {"seed": SEED, "tolerances_in_tests": {"P_max_rel": 1e-6, "f_max_abs": 2e-6, "theta_max_abs": 1e-6,
                                                 "zbc_moments_max_rel": 3e-6, "vega_pathwise_rel": 5e-6},
           "north_star_tolerance": 1e-5, "modes": {}}
    for name, mode in (("decomposed", hw._ffi.MODE_DECOMPOSED), ("reference_order", hw._ffi.MODE_REFERENCE_ORDER)):
        eng = hw.Engine(device=0)
        eng.set_mode(mode)
        rep["modes"][name] = {"vs_oracle_2^14": vs_oracle(eng, o, 1 << 14), "vs_oracle_2^17": vs_oracle(eng, o, 1 << 17),
                              "vs_reference_kernels_on_b200": vs_reference_fixture(eng),
                              "vs_reference_kernels_live": live_reference(eng)}
        eng.close()
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
