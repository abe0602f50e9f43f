import sys, os, time
sys.path.insert(0, os.getcwd())
import hw1f_b200 as hw
N = 1 << 20
eng = hw.Engine(device=0)
c = eng.bond_curve(hw.Rng(1234, N)); P, f = c["P"], c["f"]
for name, fn in [("fused", lambda i: eng.fused(hw.Rng(i, N), P, f)), ("vega", lambda i: eng.vega(hw.Rng(i, N), P, f)), ("fused again", lambda i: eng.fused(hw.Rng(i, N), P, f))]:
    ts = []
    for i in range(14):
        t0 = time.perf_counter(); fn(i); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, " ".join("%.3f" % t for t in ts))
