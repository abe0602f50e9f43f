"""The C-ABI library loads on a CPU-only box, exports every symbol include/hw1f.h declares,
and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hw1f.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hw1f_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(hw):
    declared = _declared()
    assert len(declared) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", hw.LIB_PATH], text=True)
    exported = set(re.findall(r"\b(hw1f_[a-z0-9_]+)\b", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    assert sorted(hw._ffi.SYMBOLS) == declared       # the Python binding covers the whole header


def test_sm100a_only(hw):
    out = subprocess.check_output(["cuobjdump", "-lelf", hw.LIB_PATH], text=True)
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_packed_fp32_in_sass(hw):
    """the hot loop really uses the Blackwell packed-FP32 pipe and no atomics"""
    sass = subprocess.check_output(["cuobjdump", "-sass", hw.LIB_PATH], text=True)
    assert "FFMA2" in sass and "MUFU.EX2" in sass
    body = sass.split("bond_curve_kernel")[1] if "bond_curve_kernel" in sass else sass
    assert "ATOM" not in body.split("Function :")[0]


def test_status_strings_and_params(hw):
    lib = hw._ffi.load()
    assert lib.hw1f_abi_version() == 1
    assert lib.hw1f_status_string(0) == b"ok"
    p = hw.default_params()
    assert (p.n_steps, p.n_mat) == (1000, 101)
    assert abs(p.sigma - 0.1) < 1e-8 and abs(p.theta_a1 - 0.019) < 1e-8 and abs(p.fd_theta_a1 - 0.014) < 1e-8


def test_rng_handle_semantics(hw):
    r = hw.Rng(1234, 1 << 20)
    assert r.tell() == 0
    r.seek(500)
    c = r.clone()
    assert c.tell() == 500 and c.info == {"seed": 1234, "first_path": 0, "n_paths": 1 << 20}
    r.seek(0)
    assert c.tell() == 500


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(hw):
    with pytest.raises(hw.HW1FError) as ei:
        hw.Engine(device=0)
    assert ei.value.status == hw._ffi.ERR_NO_DEVICE
