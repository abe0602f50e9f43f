#!/usr/bin/env python3
"""Wave quantisation of the simulation grid: event-timed hw1f_bond_curve / hw1f_zbc_cv for chunk counts that are and
are not multiples of the resident-block capacity (148 SMs x 2 blocks = 296 chunks of 1024 subsequences per wave).
    python tools/wave_probe.py > gpurun_out/wave_probe.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hw1f_b200 as hw  # noqa: E402

eng = hw.Engine(device=0)
mkt = eng.bond_curve(hw.Rng(1234, 1 << 20))
out = {"chunk": 1024, "slots_per_wave": 296, "q1": {}, "zbc": {}}
for chunks in (148, 296, 444, 592, 888, 1024, 1036, 1184, 1480, 2048, 2072, 2368):
    n = chunks * 1024
    for key, fn in (("q1", lambda s: eng.bond_curve(hw.Rng(s, n))["sim_ms"]),
                    ("zbc", lambda s: eng.zbc_cv(hw.Rng(s, n), mkt["P"], mkt["f"], n_steps_S1=500)["sim_ms"])):
        ms = sorted(fn(100 + i) for i in range(12))[2:-2]
        t = sum(ms) / len(ms)
        out[key][chunks] = {"ms": round(t, 4), "waves": round(chunks / 296, 3), "us_per_chunk_wave": round(1e3 * t / (chunks / 296), 2)}
print(json.dumps(out, indent=1))
