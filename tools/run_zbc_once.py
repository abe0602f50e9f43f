#!/usr/bin/env python3
"""a few hw1f_zbc_cv / hw1f_vega_pathwise / hw1f_fused calls at 2^20 subsequences (target of ncu captures)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hw1f_b200 as hw
N = 1 << 20
eng = hw.Engine(device=0)
c = eng.bond_curve(hw.Rng(1234, N))
for i in range(3):
    z = eng.zbc_cv(hw.Rng(10 + i, N), c["P"], c["f"], n_steps_S1=500)
    v = eng.vega_pathwise(hw.Rng(20 + i, N), c["P"], c["f"], n_steps_S1=500)
    f = eng.fused(hw.Rng(30 + i, N), c["P"], c["f"], n_steps_S1=500)
print(z["price_cv"], v["vega_pathwise"], f["vega"]["vega_fd"])
