#!/bin/bash
# A/B of kernel build variants on the GPU box: Q1 bench ms/step (device-timed, end-to-end, sustained) plus the event
# times of the ZBC pass and the wall time of the Q3 sequence for each library under <pkg>/lib/variants/
PKG=monte-carlo-simulation-of-hull-white-model-and-sensitivities-computation_b200
for lib in $PKG/lib/libhw1f.so $PKG/lib/variants/*.so; do
  for rep in 1 2; do
    HW1F_LIB=$PWD/$lib python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-workloads --no-scaling-run 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); print('$lib'.split('/')[-1].ljust(22), 'ms/step %.4f' % l['ms_per_step'], 'value %.4e' % l['value'], 'e2e_ms %.4f' % l['e2e']['ms_per_step'], 'sustained %.4f' % l['sustained']['ms_per_step'], 'steady-state XU %.4f' % l['roofline']['steady_state']['xu_frac'], 'mhz', l['clocks']['sm_mhz'])"
  done
  HW1F_LIB=$PWD/$lib python - <<'PY'
import time, hw1f_b200 as hw
N = 1 << 20
eng = hw.Engine(device=0)
c = eng.bond_curve(hw.Rng(1234, N)); P, f = c["P"], c["f"]
z = sorted(eng.zbc_cv(hw.Rng(i, N), P, f, n_steps_S1=500)["sim_ms"] for i in range(25))[5:-5]
def wall(fn, n=20):
    for i in range(3): fn(i)
    t0 = time.perf_counter()
    for i in range(n): fn(50 + i)
    return (time.perf_counter() - t0) * 1e3 / n
q3 = wall(lambda i: eng.vega(hw.Rng(i, N), P, f, n_steps_S1=500))
pw = wall(lambda i: eng.vega_pathwise(hw.Rng(i, N), P, f, n_steps_S1=500))
fu = wall(lambda i: eng.fused(hw.Rng(i, N), P, f, n_steps_S1=500))
rc = wall(lambda i: eng.vega_fd_recalibrated(hw.Rng(i, N), n_steps_S1=500))
zb = wall(lambda i: eng.zbc_cv_batch(list(range(i, i + 20)), N, P, f, n_steps_S1=500), 5)
print("    zbc event ms %.4f   q3 sequence wall ms %.4f   pathwise %.4f   fused %.4f   recal FD %.4f   zbc x20 %.3f" % (sum(z) / len(z), q3, pw, fu, rc, zb))
PY
done
