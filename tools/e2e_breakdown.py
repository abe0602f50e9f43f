#!/usr/bin/env python3
"""host time of each call of the bench's end-to-end step (set_model, Rng, moments, finish), median over 200 steps"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hw1f_b200 as hw
N = 1 << 20
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng = hw.Engine(device=0, stream=stream.cuda_stream)
moments = torch.zeros(202, dtype=torch.float64, device="cuda")
for i in range(5):
    eng.bond_curve_moments(hw.Rng(i, N), moments.data_ptr()); eng.bond_curve_finish(moments.data_ptr(), N)
T = {k: [] for k in ("set_model", "rng", "moments", "finish", "total")}
for i in range(200):
    t0 = time.perf_counter(); eng.set_model(eng.params)
    t1 = time.perf_counter(); rng = hw.Rng(100 + i, N)
    t2 = time.perf_counter(); eng.bond_curve_moments(rng, moments.data_ptr())
    t3 = time.perf_counter(); out = eng.bond_curve_finish(moments.data_ptr(), N)
    t4 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
        T[k].append(v * 1e6)
for k, v in T.items():
    v.sort(); print(k.ljust(10), "median %.1f us" % v[len(v) // 2])
