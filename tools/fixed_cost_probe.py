import sys, os, time
sys.path.insert(0, os.getcwd())
import hw1f_b200 as hw
N = 1 << 20
eng = hw.Engine(device=0)
c = eng.bond_curve(hw.Rng(1234, N))
P, f = c["P"], c["f"]
for n in (2, 10, 100, 500):
    ev = []
    for i in range(30):
        ev.append(eng.zbc_cv(hw.Rng(i, N), P, f, n_steps_S1=n)["sim_ms"])
    ev = sorted(ev[5:])
    print("zbc n_steps", n, "event ms median", ev[len(ev)//2])
for n in (2, 500):
    ev = []
    for i in range(30):
        ev.append(eng.vega_pathwise(hw.Rng(i, N), P, f, n_steps_S1=n)["ms_pathwise"])
    ev = sorted(ev[5:]); print("pathwise n_steps", n, ev[len(ev)//2])
