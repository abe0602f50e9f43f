"""hw1f_multi_*: single-process multi-GPU front end (needs >= 2 GPUs; skipped on a 1-GPU box)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi(hw, engine):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    lib = hw._ffi.load()
    h = C.c_void_p()
    assert lib.hw1f_multi_create(2, C.byref(h)) == 0
    p = hw.default_params()
    assert lib.hw1f_multi_set_model(h, C.byref(p)) == 0, lib.hw1f_multi_last_error(h)
    assert lib.hw1f_multi_set_mode(h, engine.mode) == 0      # same arithmetic mode as the single-GPU engine
    yield lib, h
    lib.hw1f_multi_destroy(h)


def test_multi_bond_curve_equals_single_gpu(multi, engine, hw):
    lib, h = multi
    n = (1 << 20) + 7
    P, f, se = (np.zeros(101, np.float32) for _ in range(3))
    ms = C.c_float()
    st = lib.hw1f_multi_bond_curve(h, 1234, n, 0, P.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p),
                                   se.ctypes.data_as(C.c_void_p), C.byref(ms))
    assert st == 0, lib.hw1f_multi_last_error(h)
    one = engine.bond_curve(hw.Rng(1234, n))
    assert np.abs(P / one["P"] - 1).max() < 2e-7 and np.abs(f - one["f"]).max() < 1e-6
    assert np.allclose(se[1:], one["P_se"][1:], rtol=1e-3)


def test_multi_zbc_and_vega_equal_single_gpu(multi, engine, hw):
    lib, h = multi
    n = 1 << 20
    c = engine.bond_curve(hw.Rng(1, n))
    Pm, fm = np.ascontiguousarray(c["P"]), np.ascontiguousarray(c["f"])
    res = hw.package.ZbcResult()
    st = lib.hw1f_multi_zbc_cv(h, 77, n, 0, 5.0, 10.0, engine.K_DEFAULT, Pm.ctypes.data_as(C.c_void_p),
                               fm.ctypes.data_as(C.c_void_p), 500, C.byref(res))
    assert st == 0, lib.hw1f_multi_last_error(h)
    one = engine.zbc_cv(hw.Rng(77, n), Pm, fm, n_steps_S1=500)
    assert np.allclose(list(res.mom), one["mom"], rtol=1e-12)
    assert res.price_cv == pytest.approx(one["price_cv"], rel=1e-7)
    v, se = C.c_double(), C.c_double()
    st = lib.hw1f_multi_vega_pathwise(h, 78, n, 0, 5.0, 10.0, engine.K_DEFAULT, Pm.ctypes.data_as(C.c_void_p),
                                      fm.ctypes.data_as(C.c_void_p), 500, C.byref(v), C.byref(se))
    assert st == 0, lib.hw1f_multi_last_error(h)
    onev = engine.vega_pathwise(hw.Rng(78, n), Pm, fm, n_steps_S1=500)
    assert v.value == pytest.approx(onev["vega_pathwise_f64"], rel=1e-12)
    assert se.value == pytest.approx(onev["vega_pathwise_se"], rel=1e-9)


def test_multi_fused_equals_single_gpu(multi, engine, hw):
    """the scaling-run call: fused pass sharded over two devices, one all-reduce, finish on device 0"""
    lib, h = multi
    n = (1 << 20) + 5
    c = engine.bond_curve(hw.Rng(1, 1 << 20))
    Pm, fm = np.ascontiguousarray(c["P"]), np.ascontiguousarray(c["f"])
    P, f, se = (np.zeros(101, np.float32) for _ in range(3))
    z, v, ms = hw.package.ZbcResult(), hw.package.VegaResult(), C.c_float()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    st = lib.hw1f_multi_fused(h, 4321, n, 0, 5.0, 10.0, engine.K_DEFAULT, vp(Pm), vp(fm), 0.001, 500, vp(P), vp(f),
                              vp(se), C.byref(z), C.byref(v), C.byref(ms))
    assert st == 0, lib.hw1f_multi_last_error(h)
    one = engine.fused(hw.Rng(4321, n), Pm, fm, eps=0.001, n_steps_S1=500)
    assert np.abs(P / one["P"] - 1).max() < 2e-7 and np.abs(f - one["f"]).max() < 2e-6
    assert np.allclose(list(z.mom), one["zbc"]["mom"], rtol=1e-12)
    assert z.price_cv == pytest.approx(one["zbc"]["price_cv"], rel=1e-6)
    assert v.vega_pathwise_f64 == pytest.approx(one["vega"]["vega_pathwise_f64"], rel=1e-12)
    assert v.vega_fd == pytest.approx(one["vega"]["vega_fd"], rel=1e-4)
    assert ms.value > 0


def test_peer_allreduce_kernel_vs_nccl():
    """own NVLink peer-memory all-reduce (hw1f_comm_*) against NCCL, two ranks under torchrun"""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tools", "peer_allreduce_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "peer all-reduce ok on 2 GPUs" in res.stdout


def test_sharded_bond_curve_default_engine_stream():
    """parallel.sharded_bond_curve with an engine on its own stream (stream-ordering race)"""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534",
                          os.path.join(root, "tools", "sharded_curve_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "sharded_bond_curve ok on 2 GPUs" in res.stdout
