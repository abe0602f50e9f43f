"""Engine vs the UNMODIFIED reference kernels run on the same B200 with the same seeds
(oracle/_ref/ref_harness, built from /root/reference by oracle/Makefile).  This is the north-star
parity statement: integer RNG state bit-exact; P(0,T), theta, ZBC price, beta and vega within 1e-5
relative, or within the Monte Carlo CI where the reference's float-atomic summation dominates."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
SEED = 20251018
N = 1 << 20


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    out = tmp_path_factory.mktemp("ref") / "parity.json"
    subprocess.run([HARNESS, "parity", str(SEED), str(out)], check=True, stdout=subprocess.DEVNULL, timeout=600,
                   cwd=str(out.parent))
    with open(out) as f:
        r = json.load(f)
    keep = os.environ.get("HW1F_SAVE_REF")
    if keep:
        with open(keep, "w") as f:
            json.dump(r, f)
    return r


@pytest.fixture(scope="module")
def mine(engine, hw, ref):
    curve = engine.bond_curve(hw.Rng(SEED, N))
    return curve


def test_reference_rng_states_bit_exact(engine, hw, ref):
    rng = hw.Rng(SEED, N)
    for st in ref["states"]:
        state, _ = engine.debug_rng(rng, st["path"], 0)
        assert int(state[0]) == st["d"] and state[1:].tolist() == st["v"], st


def test_reference_curve(mine, ref, oracle):
    """The reference accumulates P_sum[m] with float32 atomics (market_data.cuh:63,73): 1024 block
    partials of ~2046 are added to a running sum of ~2e6 whose ulp is 0.125.  At short maturities the
    partials are nearly identical, the rounding is the same at every add, and the error is a BIAS of
    up to 1024*0.0625/2.1e6 = 3e-5 relative (observed ~1e-5 at T=0.1), far outside its own MC
    interval; f = -d ln P/dT amplifies it tenfold at T=0.  So: the engine must sit on the
    double-precision oracle (same paths, same seeds) to 1e-6, and the reference must sit within its
    own float32-accumulation error bound of both."""
    P, f = np.array(ref["P"], np.float32), np.array(ref["f"], np.float32)
    P_orc, f_orc = oracle.bond_curve(SEED, N)            # full size: a few seconds on the box's cores
    assert np.abs(mine["P"] / P_orc - 1).max() < 1e-6
    assert np.abs(mine["f"] - f_orc).max() < 2e-6
    assert np.abs(P / P_orc - 1).max() < 3e-5             # the reference's own float-atomic error
    assert np.abs(mine["P"] / P - 1).max() < 3e-5
    assert np.abs(mine["P"][20:] / P[20:] - 1).max() < 1e-5   # dispersed partials: north-star tolerance
    assert np.abs(mine["f"] - f).max() < 3e-4
    assert np.abs(mine["f"][20:] - f[20:]).max() < 3e-5


def test_reference_theta(engine, mine, ref):
    got = engine.theta_calibrate(np.array(ref["f"], np.float32))
    assert np.abs(got["theta_rec"] - np.array(ref["theta_rec"], np.float32)).max() < 1e-6
    assert (got["theta_ref"] == np.array(ref["theta_orig"], np.float32)).all()


def test_reference_zbc(engine, hw, ref):
    n_steps = engine.steps_to(5.0)
    P, f = np.array(ref["P"], np.float32), np.array(ref["f"], np.float32)
    z = engine.zbc_cv(hw.Rng(SEED + 54321, N), P, f, n_steps_S1=n_steps)
    assert np.allclose(z["mom"], ref["zbc_moments"], rtol=2e-5)
    assert z["mean_X"] == pytest.approx(ref["zbc_mean_X"], rel=1e-5)
    assert z["price_cv"] == pytest.approx(ref["zbc_price_cv"], rel=1e-5)
    # beta / rho come from E[XY]-E[X]E[Y] in float32: cancellation floor ~1e-4
    assert z["beta"] == pytest.approx(ref["zbc_beta"], rel=5e-4)
    assert z["corr"] == pytest.approx(ref["zbc_corr"], rel=5e-4)


def test_reference_vega(engine, hw, ref):
    n_steps = engine.steps_to(5.0)
    P, f = np.array(ref["P"], np.float32), np.array(ref["f"], np.float32)
    v = engine.vega(hw.Rng(SEED, N), P, f, n_steps_S1=n_steps)
    assert v["vega_pathwise"] == pytest.approx(ref["vega_pathwise"], rel=1e-5)
    # FD quotients amplify float32 price noise by 500: compare inside the MC standard error
    assert v["vega_fd"] == pytest.approx(ref["vega_fd"], abs=5e-4)
    assert v["vega_fd_recal"] == pytest.approx(ref["vega_fd_recal"], abs=5e-3)


def test_reference_sample_paths(engine, hw, ref):
    got = engine.sample_paths(hw.Rng(SEED, N).seek(1000), 32)
    assert np.abs(got[0] - np.array(ref["r_path0"], np.float32)).max() < 2e-6
    assert np.abs(got[31] - np.array(ref["r_path31"], np.float32)).max() < 2e-6
