import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import hw1f_b200 as hw
for over in (dict(n_steps=8184, n_mat=1024), dict(n_steps=1023, n_mat=1024, T_final=10.23), dict(n_steps=2, n_mat=3, T_final=0.5), dict(n_steps=8192, n_mat=513)):
  for mode in (0, 1):
    eng = hw.Engine(device=0, params=hw.default_params(**over)); eng.set_mode(mode)
    n = (1 << 13) + 5
    c = eng.bond_curve(hw.Rng(1, n))
    S1, S2 = 0.5 * eng.params.T_final, eng.params.T_final
    ns = eng.steps_to(S1)
    print(over, mode, "ns", ns, "P_end", c["P"][-1])
    for name, fn in (("zbc", lambda: eng.zbc_cv(hw.Rng(2, n), c["P"], c["f"], S1=S1, S2=S2, n_steps_S1=ns)["price_cv"]),
                     ("vega", lambda: {k: v for k, v in eng.vega(hw.Rng(3, n), c["P"], c["f"], S1=S1, S2=S2, n_steps_S1=ns).items() if k.startswith("vega") and not k.endswith("se")}),
                     ("fd_recal", lambda: eng.vega_fd_recalibrated(hw.Rng(3, n), S1=S1, S2=S2, n_steps_S1=ns)["vega_fd_recal"]),
                     ("fused", lambda: eng.fused(hw.Rng(4, n), c["P"], c["f"], S1=S1, S2=S2, n_steps_S1=ns)["vega"]["vega_fd"]),
                     ("ci", lambda: float(eng.bond_curve_ci()["f_se"].max())),
                     ("batch", lambda: eng.zbc_cv_batch([1, 2], n, c["P"], c["f"], S1=S1, S2=S2, n_steps_S1=ns)[0][0]),
                     ("paths", lambda: eng.sample_paths(hw.Rng(5, n), 4).shape)):
        try:
            print("   ", name, fn())
        except Exception as e:
            print("   ", name, "ERR", str(e)[:160])
    eng.close()
