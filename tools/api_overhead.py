#!/usr/bin/env python3
"""Host overhead of the public calls: wall-clock per call next to the CUDA-event time of its simulation
part (sim_ms / ms_* fields).  Run on the GPU box: python tools/api_overhead.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hw1f_b200 as hw  # noqa: E402

N = 1 << 20


def timed(fn, steps=20, warmup=3):
    for i in range(warmup):
        fn(i)
    ev = []
    t0 = time.perf_counter()
    for i in range(steps):
        ev.append(fn(100 + i))
    wall = (time.perf_counter() - t0) * 1e3 / steps
    return wall, sum(ev) / len(ev)


def main():
    eng = hw.Engine(device=0)
    c = eng.bond_curve(hw.Rng(1234, N))
    P, f = c["P"], c["f"]
    out = {}
    out["bond_curve"] = timed(lambda i: eng.bond_curve(hw.Rng(i, N))["sim_ms"])
    out["zbc_cv(n=-1)"] = timed(lambda i: eng.zbc_cv(hw.Rng(i, N), P, f)["sim_ms"])
    out["zbc_cv(n=500)"] = timed(lambda i: eng.zbc_cv(hw.Rng(i, N), P, f, n_steps_S1=500)["sim_ms"])
    out["vega_pathwise(n=500)"] = timed(lambda i: eng.vega_pathwise(hw.Rng(i, N), P, f, n_steps_S1=500)["ms_pathwise"])
    out["vega_fd(n=500)"] = timed(lambda i: eng.vega_fd(hw.Rng(i, N), P, f, n_steps_S1=500)["ms_fd"])
    out["vega_fd_recalibrated(n=500)"] = timed(lambda i: eng.vega_fd_recalibrated(hw.Rng(i, N), n_steps_S1=500)["ms_fd_recal"])
    out["vega sequence(n=500)"] = timed(lambda i: (lambda v: v["ms_pathwise"] + v["ms_fd"] + v["ms_fd_recal"])(
        eng.vega(hw.Rng(i, N), P, f, n_steps_S1=500)))
    out["fused(n=500)"] = timed(lambda i: eng.fused(hw.Rng(i, N), P, f, n_steps_S1=500)["sim_ms"])
    res = {k: {"wall_ms": round(w, 4), "event_ms": round(e, 4), "host_overhead_ms": round(w - e, 4)} for k, (w, e) in out.items()}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
