"""Per-workload wall-clock of the engine's public API next to the reference's own host functions
(oracle/_ref/ref_harness workload ...).  Run on the GPU box:
    python tools/workloads.py > gpurun_out/workloads.json
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import hw1f_b200 as hw  # noqa: E402

N = 1 << 20
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def timed(fn, steps=10, warmup=2):
    for i in range(warmup):
        fn(i)
    t0 = time.perf_counter()
    for i in range(steps):
        fn(100 + i)
    return (time.perf_counter() - t0) * 1e3 / steps


def main():
    eng = hw.Engine(device=0)
    c = eng.bond_curve(hw.Rng(1234, N))
    P, f = c["P"], c["f"]
    out = {"engine_ms": {}, "reference_ms": {}, "n_paths": N}
    out["engine_ms"]["q1_bond_curve"] = timed(lambda i: eng.bond_curve(hw.Rng(i, N)))
    out["engine_ms"]["q2b_zbc_cv"] = timed(lambda i: eng.zbc_cv(hw.Rng(i, N), P, f))
    out["engine_ms"]["q3_sequence"] = timed(lambda i: eng.vega(hw.Rng(i, N), P, f))
    out["engine_ms"]["q3_pathwise_only"] = timed(lambda i: eng.vega_pathwise(hw.Rng(i, N), P, f))
    out["engine_ms"]["zbc_validation_20_seeds"] = timed(
        lambda i: eng.zbc_cv_batch([i * 1000003 + r * 12345 for r in range(20)], N, P, f), steps=5, warmup=1)
    out["engine_ms"]["vega_validation_20_seeds"] = timed(
        lambda i: eng.vega_pathwise_batch([i * 1000003 + r * 982451653 for r in range(20)], N, P, f), steps=5, warmup=1)
    if os.path.exists(HARNESS):
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, "data"))
            for name, key, steps in (("q3seq", "q3_sequence", 10), ("zbc20", "zbc_validation_20_seeds", 3),
                                     ("vega20", "vega_validation_20_seeds", 3)):
                o = os.path.join(td, name + ".json")
                subprocess.run([HARNESS, "workload", name, str(steps), "1", o], check=True, cwd=td,
                               stdout=subprocess.DEVNULL, timeout=900)
                out["reference_ms"][key] = json.load(open(o))["wall_ms_per_step"]
            for q, key in (("q1", "q1_bond_curve"), ("q2", "q2b_zbc_cv"), ("q3", "q3_pathwise_only")):
                o = os.path.join(td, q + ".json")
                subprocess.run([HARNESS, "bench", q, "20", "3", o], check=True, cwd=td, stdout=subprocess.DEVNULL,
                               timeout=900)
                r = json.load(open(o))
                out["reference_ms"][key] = r["workload_ms_per_step"]
                out["reference_ms"][key + "_kernel_only"] = r["kernel_ms_per_step"]
    out["speedup"] = {k: out["reference_ms"][k] / v for k, v in out["engine_ms"].items() if k in out["reference_ms"]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
