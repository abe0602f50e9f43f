#!/usr/bin/env python3
"""small pass over every decomposed / reference-order entry point (target of compute-sanitizer runs)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hw1f_b200 as hw
N = (1 << 12) + 77          # ragged: the last chunk is partly masked
for mode in (hw._ffi.MODE_DECOMPOSED, hw._ffi.MODE_REFERENCE_ORDER):
    eng = hw.Engine(device=0)
    eng.set_mode(mode)
    c = eng.bond_curve(hw.Rng(7, N))
    z = eng.zbc_cv(hw.Rng(8, N), c["P"], c["f"], n_steps_S1=500)
    z2 = eng.zbc_cv(hw.Rng(8, N).seek(499), c["P"], c["f"], n_steps_S1=499)
    v = eng.vega(hw.Rng(9, N), c["P"], c["f"], n_steps_S1=500)
    r = eng.vega_fd_recalibrated(hw.Rng(9, N), n_steps_S1=495)
    f = eng.fused(hw.Rng(10, N), c["P"], c["f"], n_steps_S1=500)
    b, _ = eng.zbc_cv_batch([1, 2, 3], N, c["P"], c["f"], n_steps_S1=500)
    s = eng.sample_paths(hw.Rng(11, N), 32)
    N2 = 1 << 14                      # >= 8 simulation blocks: the batch-means confidence intervals of f and theta
    eng.bond_curve(hw.Rng(12, N2))
    ci = eng.bond_curve_ci()
    vb, _ = eng.vega_pathwise_batch([4, 5], N, c["P"], c["f"], n_steps_S1=500)
    for m in range(4):
        eng.reduction_bench(hw.Rng(13, N), m, c["P"], c["f"], n_steps_S1=500, n_warmup=1, n_runs=1)
    p = hw.default_params(n_steps=500)   # odd save stride (5 steps per maturity): the ODD instantiations
    eng.set_model(p)
    c5 = eng.bond_curve(hw.Rng(14, N))
    r5 = eng.vega_fd_recalibrated(hw.Rng(15, N), n_steps_S1=250)
    print(mode, float(c["P"][-1]), z["price_cv"], v["vega_fd_recal"], f["vega"]["vega_fd"], len(b), s.shape)
    eng.close()
