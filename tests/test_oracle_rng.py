"""Pins the CPU oracle's XORWOW (oracle/xorwow_ref.c, jump matrices rebuilt from the recurrence)
and the ENGINE's host jump algebra against golden vectors produced by cuRAND's own header
(tests/golden/xorwow_golden.json, generator: oracle/ref/gen_curand_golden.cu)."""
import json
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "xorwow_golden.json")) as f:
        return json.load(f)


def _fold(draws):
    x = 0
    for k, d in enumerate(draws):
        x ^= (int(d) << (k & 31))
    return x & (2 ** 64 - 1), int(np.sum(draws.astype(np.uint64)))


def test_golden_shape(golden):
    assert golden["sizeof_curandState"] == 48
    assert len(golden["cases"]) >= 20


def test_oracle_matches_curand_header(oracle, golden):
    for c in golden["cases"]:
        d = oracle.draws(c["seed"], c["subsequence"], c["offset"], 2000)
        assert d[:8].tolist() == c["draws_head"], c
        assert int(d[499]) == c["draw_499"] and int(d[999]) == c["draw_999"] and int(d[1999]) == c["draw_1999"]
        x, s = _fold(d)
        assert x == c["xor2000"] and s == c["sum2000"]


def test_engine_host_jump_matches_curand_header(hw, golden):
    eng_mod = hw.package.engine
    for c in golden["cases"]:
        if c["offset"] % 2:
            continue
        st = eng_mod.host_rng_state(c["seed"], c["subsequence"], c["offset"])
        assert int(st[0]) == c["d"], c
        assert st[1:].tolist() == c["v"], c


def test_offset_equals_stepping(oracle):
    a = oracle.draws(42, 777, 0, 1500)
    b = oracle.draws(42, 777, 1000, 500)
    assert (a[1000:] == b).all()


def test_normal_stream_offsets(oracle):
    """normal k of a path is the same whatever offset the stream is opened at (incl. odd offsets:
    the cached cos-branch value of curand_normal, curand_normal.h:313-326)"""
    full = oracle.normals(1234, 9, 0, 64)
    for off in (1, 2, 7, 31, 32):
        part = oracle.normals(1234, 9, off, 64 - off)
        assert (full[off:] == part).all(), off
    assert np.isfinite(full).all() and abs(float(full.mean())) < 0.5
