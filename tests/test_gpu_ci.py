"""The engine's own confidence intervals (for f, theta, beta*, rho the parity criterion is "inside the
95 % CI", and the reference reports none of them): every standard error the engine returns is checked against the
empirical spread of the estimator over independent seeds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 1 << 16          # 64 simulation blocks = 64 batches
SEEDS = [900001 + 7919 * k for k in range(40)]


@pytest.fixture(scope="module")
def runs(engine, hw):
    out = []
    for s in SEEDS:
        c = engine.bond_curve(hw.Rng(s, N))
        ci = engine.bond_curve_ci()
        th = engine.theta_calibrate(c["f"])
        out.append((c, ci, th))
    return out


def _ratio(se_mean, samples):
    sd = samples.std(axis=0, ddof=1)
    ok = sd > 0
    return se_mean[ok] / sd[ok]


def _first_maturity(engine, hw):
    """reference-order arithmetic forms p0 - c in float32, whose ulp (2.4e-7) is the size of the deviation itself at
    T <= 0.2 (sd of p0 there: 4e-7 .. 1e-6): the spread is quantisation, not Monte Carlo noise, and no variance
    estimate can be right.  The decomposed form evaluates the deviation directly (z^2 p(z^2)) and is checked from T = 0.1."""
    return 1 if engine.mode == hw._ffi.MODE_DECOMPOSED else 3


def test_P_se_exact_and_batch_means(runs, engine, hw):
    m0 = _first_maturity(engine, hw)
    P = np.array([r[0]["P"] for r in runs], np.float64)
    exact = np.array([r[0]["P_se"] for r in runs], np.float64).mean(axis=0)
    batch = np.array([r[1]["P_se_batch"] for r in runs], np.float64).mean(axis=0)
    # SD of an SD estimated from 40 samples: +-11 %; batch means over 64 batches: +-9 % per run, averaged over 40 runs
    r_exact = _ratio(exact[m0:], P[:, m0:])
    r_batch = _ratio(batch[m0:], P[:, m0:])
    assert 0.7 < np.median(r_exact) < 1.3 and r_exact.min() > 0.55 and r_exact.max() < 1.6
    assert 0.7 < np.median(r_batch) < 1.3 and r_batch.min() > 0.55 and r_batch.max() < 1.6
    assert np.abs(batch[m0:] / exact[m0:] - 1).max() < 0.15


def test_f_se_matches_the_spread_over_seeds(runs, engine, hw):
    m0 = _first_maturity(engine, hw) + 1
    f = np.array([r[0]["f"] for r in runs], np.float64)[:, m0:]
    se = np.array([r[1]["f_se"] for r in runs], np.float64).mean(axis=0)[m0:]
    r = _ratio(se, f)
    assert 0.75 < np.median(r) < 1.25, np.median(r)
    assert r.min() > 0.55 and r.max() < 1.6, (r.min(), r.max())


def test_theta_se_matches_the_spread_over_seeds(runs, engine, hw):
    m0 = _first_maturity(engine, hw) + 2
    th = np.array([r[2]["theta_rec"] for r in runs], np.float64)[:, m0:]
    se = np.array([r[1]["theta_se"] for r in runs], np.float64).mean(axis=0)[m0:]
    r = _ratio(se, th)
    assert 0.75 < np.median(r) < 1.25, np.median(r)
    assert r.min() > 0.5 and r.max() < 1.7, (r.min(), r.max())


def test_ci_needs_a_curve_launch_and_enough_blocks(engine, hw):
    c = engine.bond_curve(hw.Rng(1, 1 << 12))           # 4 blocks
    with pytest.raises(hw.package.engine.HW1FError):
        engine.bond_curve_ci()
    c = engine.bond_curve(hw.Rng(1, N))
    engine.zbc_cv(hw.Rng(2, N), c["P"], c["f"], n_steps_S1=500)   # overwrites the block partials
    with pytest.raises(hw.package.engine.HW1FError):
        engine.bond_curve_ci()


def test_beta_and_rho_se_match_the_spread_over_seeds(engine, hw):
    c = engine.bond_curve(hw.Rng(4242, 1 << 18))
    res, _ = engine.zbc_cv_batch(SEEDS[:32], N, c["P"], c["f"], n_steps_S1=500)
    beta = np.array([r["beta_f64"] for r in res])
    rho = np.array([r["corr_f64"] for r in res])
    price = np.array([r["price_cv_f64"] for r in res])
    # one degree of freedom per antithetic pair is conservative: the stated SE may exceed the spread, never undercut it by much
    for name, x, se in (("beta", beta, np.mean([r["beta_se"] for r in res])), ("rho", rho, np.mean([r["corr_se"] for r in res])),
                        ("price", price, np.mean([r["se_cv"] for r in res]))):
        ratio = se / x.std(ddof=1)
        assert 0.7 < ratio < 2.5, (name, ratio)
