"""small driver used under ncu: a few ZBC / pathwise / FD / fused launches at full size"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hw1f_b200 as hw
N = 1 << 20
eng = hw.Engine(device=0)
c = eng.bond_curve(hw.Rng(1234, N))
for i in range(3):
    eng.zbc_cv(hw.Rng(10 + i, N), c["P"], c["f"], n_steps_S1=500)
    eng.vega(hw.Rng(20 + i, N), c["P"], c["f"], n_steps_S1=500)
print("ok", eng.launch_count)
