#!/bin/bash
# A/B of kernel build variants on the GPU box: prints ms/step of the Q1 bench for each library
PKG=monte-carlo-simulation-of-hull-white-model-and-sensitivities-computation_b200
for lib in $PKG/lib/libhw1f.so $PKG/lib/variants/*.so; do
  for rep in 1 2; do
    HW1F_LIB=$PWD/$lib python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); print('$lib'.split('/')[-1].ljust(22), 'ms/step %.4f' % l['ms_per_step'], 'value %.4e' % l['value'], 'e2e_ms %.4f' % l['e2e']['ms_per_step'], 'mhz', l['clocks']['sm_mhz'])"
  done
done
