"""Prints the measured pipe rates of this GPU (roofline denominators) as JSON.
Run on the GPU box: python tools/pipe_probe.py > gpurun_out/pipe_probe.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hw1f_b200 as hw  # noqa: E402

NAMES = ["ffma", "ffma2", "mufu_ex2", "lop3_shf", "i2fp(+lop3)", "mufu+i2fp", "hw1f_mix", "fmul2", "mufu_lg2", "mufu_sqrt",
         "mufu_sin(+fmul.rz)", "mufu_cos(+fmul.rz)", "mufu_box_muller_mix", "mufu_box_muller_mix_under_q1_load"]
eng = hw.Engine(device=0)
out = {}
for which, name in enumerate(NAMES):
    best = None
    for _ in range(3):
        ms, n = eng.pipe_probe(which, 256 if which not in (4, 5, 6, 13) else 64)
        r = n / (ms * 1e-3)
        best = r if best is None else max(best, r)
    out[name] = {"thread_instr_per_s": best, "per_sm_per_clk_at_1965MHz": best / (148 * 1.965e9)}
print(json.dumps(out, indent=1))
