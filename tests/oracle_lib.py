"""ctypes binding of the CPU oracle (oracle/_build/libhw1f_oracle.so).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only.  The product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libhw1f_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [
        ("a", C.c_float), ("sigma", C.c_float), ("r0", C.c_float), ("T_final", C.c_float),
        ("n_steps", C.c_int), ("n_mat", C.c_int),
        ("theta_a0", C.c_float), ("theta_b0", C.c_float),
        ("theta_a1", C.c_float), ("theta_b1", C.c_float),
        ("theta_break", C.c_float), ("fd_theta_a1", C.c_float),
    ]


class OrcZbcResult(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "mean_X", "mean_Y", "var_Y", "var_X", "cov", "beta", "price_cv", "corr_single", "corr",
        "control_adjustment")]


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("hw1f_oracle.c", "xorwow_ref.c", "hw1f_oracle.h", "xorwow_ref.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_dt.restype = C.c_float
        _lib.orc_mat_spacing.restype = C.c_float
        _lib.orc_exp_adt.restype = C.c_float
        _lib.orc_sig_st.restype = C.c_float
        _lib.orc_sig_st.argtypes = [C.POINTER(OrcParams), C.c_float]
        _lib.orc_steps_to.argtypes = [C.POINTER(OrcParams), C.c_float]
        _lib.orc_drift_tables.argtypes = [C.POINTER(OrcParams), C.c_float, C.c_void_p, C.c_void_p]
        _lib.orc_shifted_drift_table.argtypes = [C.POINTER(OrcParams), C.c_float, C.c_float, C.c_void_p]
        _lib.orc_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        _lib.orc_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        _lib.orc_bond_curve_sums.argtypes = [C.POINTER(OrcParams), C.c_float, C.c_void_p, C.c_uint64, C.c_uint64,
                                             C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib.orc_curve_finalize.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.orc_theta.argtypes = [C.POINTER(OrcParams), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_zbc_moments.argtypes = [C.POINTER(OrcParams), C.c_float, C.c_float, C.c_void_p, C.c_uint64,
                                         C.c_uint64, C.c_int64, C.c_uint64, C.c_int, C.c_float, C.c_float,
                                         C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_zbc_algebra.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(OrcZbcResult)]
        _lib.orc_vega_pathwise_sums.argtypes = [C.POINTER(OrcParams), C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                                C.c_uint64, C.c_uint64, C.c_int64, C.c_uint64, C.c_int, C.c_float,
                                                C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_sample_paths.argtypes = [C.POINTER(OrcParams), C.c_float, C.c_void_p, C.c_uint64, C.c_uint64,
                                          C.c_int, C.c_uint64, C.c_void_p]
        _lib.orc_run_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Convenience wrapper: the reference's estimators on the CPU."""

    K_DEFAULT = float(np.exp(np.float32(-0.1)).astype(np.float32))  # expf(-0.1f), src/2:110

    def __init__(self, **overrides):
        self.L = lib()
        self.p = OrcParams()
        self.L.orc_default_params(C.byref(self.p))
        for k, v in overrides.items():
            setattr(self.p, k, v)

    # -- model constants --
    @property
    def dt(self):
        return self.L.orc_dt(C.byref(self.p))

    def sig_st(self, sigma=None):
        return self.L.orc_sig_st(C.byref(self.p), self.p.sigma if sigma is None else sigma)

    def steps_to(self, S1):
        return self.L.orc_steps_to(C.byref(self.p), S1)

    def drift_tables(self, sigma=None):
        n = self.p.n_steps
        d = np.zeros(n, np.float32)
        s = np.zeros(n, np.float32)
        self.L.orc_drift_tables(C.byref(self.p), self.p.sigma if sigma is None else sigma, _p(d), _p(s))
        return d, s

    def shifted_drift_table(self, sigma_new, sigma_old=None):
        d = np.zeros(self.p.n_steps, np.float32)
        self.L.orc_shifted_drift_table(C.byref(self.p), sigma_new, self.p.sigma if sigma_old is None else sigma_old,
                                       _p(d))
        return d

    # -- RNG --
    def draws(self, seed, subsequence, offset, n):
        out = np.zeros(n, np.uint32)
        self.L.orc_draws(seed, subsequence, offset, n, _p(out))
        return out

    def normals(self, seed, subsequence, offset_normals, n):
        out = np.zeros(n, np.float32)
        self.L.orc_normals(seed, subsequence, offset_normals, n, _p(out))
        return out

    # -- Q1 --
    def bond_curve_sums(self, seed, n_pairs, first_path=0, offset=0, sigma=None, drift=None):
        sigma = self.p.sigma if sigma is None else sigma
        if drift is None:
            drift, _ = self.drift_tables(sigma)
        s = np.zeros(self.p.n_mat, np.float64)
        q = np.zeros(self.p.n_mat, np.float64)
        self.L.orc_bond_curve_sums(C.byref(self.p), self.sig_st(sigma), _p(drift), seed, first_path, n_pairs,
                                   offset, _p(s), _p(q))
        return s, q

    def curve_finalize(self, sums, n_pairs):
        sf = np.asarray(sums, np.float64).astype(np.float32)
        P = np.zeros(self.p.n_mat, np.float32)
        f = np.zeros(self.p.n_mat, np.float32)
        self.L.orc_curve_finalize(C.byref(self.p), _p(sf), n_pairs, _p(P), _p(f))
        return P, f

    def bond_curve(self, seed, n_pairs, **kw):
        s, _ = self.bond_curve_sums(seed, n_pairs, **kw)
        return self.curve_finalize(s, n_pairs)

    def theta(self, f, sigma=None):
        f = np.ascontiguousarray(f, np.float32)
        n = self.p.n_mat
        rec, orig, Ts = (np.zeros(n, np.float32) for _ in range(3))
        self.L.orc_theta(C.byref(self.p), self.p.sigma if sigma is None else sigma, _p(f), _p(rec), _p(orig), _p(Ts))
        return rec, orig, Ts

    # -- Q2b --
    def zbc_moments(self, seed, n_pairs, P_mkt, f_mkt, S1=5.0, S2=10.0, K=None, n_steps_S1=None, first_path=0,
                    offset=0, sigma=None, drift=None):
        sigma = self.p.sigma if sigma is None else sigma
        if drift is None:
            drift, _ = self.drift_tables(sigma)
        K = self.K_DEFAULT if K is None else K
        n_steps_S1 = self.steps_to(S1) if n_steps_S1 is None else n_steps_S1
        P_mkt = np.ascontiguousarray(P_mkt, np.float32)
        f_mkt = np.ascontiguousarray(f_mkt, np.float32)
        mom = np.zeros(5, np.float64)
        self.L.orc_zbc_moments(C.byref(self.p), sigma, self.sig_st(sigma), _p(drift), seed, first_path, n_pairs,
                               offset, n_steps_S1, S1, S2, K, _p(P_mkt), _p(f_mkt), _p(mom))
        return mom

    def zbc_algebra(self, mom, n_total, P0S2):
        mf = np.asarray(mom, np.float64).astype(np.float32)
        r = OrcZbcResult()
        self.L.orc_zbc_algebra(_p(mf), n_total, P0S2, C.byref(r))
        return {n: getattr(r, n) for n, _ in OrcZbcResult._fields_}

    # -- Q3 --
    def vega_pathwise_sums(self, seed, n_paths, P_mkt, f_mkt, S1=5.0, S2=10.0, K=None, n_steps_S1=None,
                           first_path=0, offset=0, sigma=None):
        sigma = self.p.sigma if sigma is None else sigma
        drift, sdrift = self.drift_tables(sigma)
        K = self.K_DEFAULT if K is None else K
        n_steps_S1 = self.steps_to(S1) if n_steps_S1 is None else n_steps_S1
        P_mkt = np.ascontiguousarray(P_mkt, np.float32)
        f_mkt = np.ascontiguousarray(f_mkt, np.float32)
        s = C.c_double()
        q = C.c_double()
        self.L.orc_vega_pathwise_sums(C.byref(self.p), sigma, self.sig_st(sigma), _p(drift), _p(sdrift), seed,
                                      first_path, n_paths, offset, n_steps_S1, S1, S2, K, _p(P_mkt), _p(f_mkt),
                                      C.byref(s), C.byref(q))
        return s.value, q.value

    def sample_paths(self, seed, n_show, first_path=0, offset=0, sigma=None):
        sigma = self.p.sigma if sigma is None else sigma
        drift, _ = self.drift_tables(sigma)
        out = np.zeros((n_show, self.p.n_steps + 1), np.float32)
        self.L.orc_sample_paths(C.byref(self.p), self.sig_st(sigma), _p(drift), seed, first_path, n_show, offset,
                                _p(out))
        return out

    def run_stats(self, samples):
        s = np.ascontiguousarray(samples, np.float32)
        out = np.zeros(8, np.float32)
        self.L.orc_run_stats(_p(s), len(s), _p(out))
        return dict(zip(("mean", "var", "sd", "se", "moe", "lo", "hi", "cv_pct"), out.tolist()))

    def max_threads(self):
        return self.L.orc_max_threads()
