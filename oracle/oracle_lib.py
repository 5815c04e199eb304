"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by cosmopmc_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cosmopmc_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

_vp, _i, _i64, _u64, _u32, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_double
_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(HERE, "pmc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        sig = {
            "orc_cholesky": (_i, [_i, _vp]),
            "orc_mvdens_log_pdf": (_d, [_i, _i, _vp, _vp, _vp]),
            "orc_mix_log_pdf": (_d, [_i, _i, _i, _vp, _vp, _vp, _vp]),
            "orc_mix_log_pdf_batch": (None, [_i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
            "orc_philox4x32_10": (None, [_vp, _vp, _vp]),
            "orc_select_component": (_i, [_i, _vp, _d]),
            "orc_sample_draws": (None, [_u64, _u32, _i64, _i, _i, _pd, _vp, _pd]),
            "orc_simulate": (_i64, [_i64, _u64, _u32, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
            "orc_simulate_from_draws": (_i64, [_i64, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
            "orc_Esqr": (_d, [C.POINTER(A.Cosmo), _d, _i]),
            "orc_w": (_d, [C.POINTER(A.Cosmo), _d, _i, _pi, _pi]),
            "orc_f_K": (_d, [C.POINTER(A.Cosmo), _d, _i]),
            "orc_D_lum": (_d, [C.POINTER(A.Cosmo), _d, _pi]),
            "orc_loglike": (_d, [C.POINTER(A.Like), _vp, _pi]),
            "orc_posterior_log_pdf": (_d, [C.POINTER(A.Target), _vp, _pi]),
            "orc_posterior_log_pdf_batch": (None, [C.POINTER(A.Target), _i64, _vp, _vp, _vp, _i]),
            "orc_sn_mean_stages": (_d, [C.POINTER(A.Like), _vp]),
            "orc_map_params": (_i, [C.POINTER(A.Like), _vp, _vp]),
            "orc_importance_weights": (_i64, [C.POINTER(A.Target), _i64, _vp, _i, _i, _i, _vp, _vp, _vp, _d, _vp, _vp, _pd, _i]),
            "orc_normalize_weights": (_d, [_i64, _vp, _vp, _d, _pd]),
            "orc_perplexity_and_ess": (_d, [_i64, _vp, _vp, _pd]),
            "orc_enc": (_d, [_i, _vp]),
            "orc_update_prop_rb": (_i, [_i64, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
            "orc_iteration": (_i, [C.POINTER(A.Target), _i64, _u64, _u32, _d, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(A.Stats), _i]),
            "orc_iteration_mode_b": (_i, [C.POINTER(A.Target), _i64, _u64, _u32, _d, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(A.Stats), _i]),
        }
        for n, (r, a) in sig.items():
            f = getattr(L, n)
            f.restype, f.argtypes = r, a
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


# ---- numpy-friendly wrappers ---------------------------------------------------
def cholesky_stack(cov):
    cov = f64(cov).copy()
    K, d, _ = cov.shape
    for k in range(K):
        a = np.ascontiguousarray(cov[k])
        rc = lib().orc_cholesky(d, _p(a))
        if rc:
            raise ValueError("component %d not positive definite" % k)
        cov[k] = a
    return cov


def mix_log_pdf(X, wght, mean, chol, df=-1):
    X, wght, mean, chol = f64(X), f64(wght), f64(mean), f64(chol)
    N, d = X.shape
    out = np.empty(N)
    lib().orc_mix_log_pdf_batch(N, len(wght), d, df, _p(wght), _p(mean), _p(chol), _p(X), _p(out))
    return out


def simulate(N, seed, it, offset, wght, mean, chol, bmin, bmax, df=-1):
    wght, mean, chol = f64(wght), f64(mean), f64(chol)
    bmin, bmax = f64(bmin), f64(bmax)
    K, d = mean.shape
    X = np.empty((N, d))
    idx = np.empty(N, dtype=np.int32)
    flg = np.empty(N, dtype=np.int16)
    nok = lib().orc_simulate(N, seed, it, offset, K, d, df, _p(wght), _p(mean), _p(chol),
                             _p(bmin), _p(bmax), _p(X), _p(idx), _p(flg))
    return X, idx, flg, nok


def simulate_from_draws(u, z, wght, mean, chol, bmin, bmax):
    u, z, wght, mean, chol = f64(u), f64(z), f64(wght), f64(mean), f64(chol)
    bmin, bmax = f64(bmin), f64(bmax)
    N, d = z.shape
    X = np.empty((N, d))
    idx = np.empty(N, dtype=np.int32)
    flg = np.empty(N, dtype=np.int16)
    nok = lib().orc_simulate_from_draws(N, _p(u), _p(z), len(wght), d, _p(wght), _p(mean),
                                        _p(chol), _p(bmin), _p(bmax), _p(X), _p(idx), _p(flg))
    return X, idx, flg, nok


def posterior_log_pdf(spec, X, nthreads=0):
    X = f64(X)
    N = X.shape[0]
    out = np.empty(N)
    err = np.zeros(N, dtype=np.int32)
    lib().orc_posterior_log_pdf_batch(C.byref(spec.t), N, _p(X), _p(out), _p(err), nthreads)
    return out, err


def map_params(spec, idata, X):
    """mapped model of data set idata for every row of X: (N, 16) array + error flags"""
    X = f64(X)
    out = np.zeros((X.shape[0], 16))
    err = np.zeros(X.shape[0], dtype=np.int32)
    for n in range(X.shape[0]):
        row = np.zeros(16)
        err[n] = lib().orc_map_params(C.byref(spec.t.like[idata]), _p(np.ascontiguousarray(X[n])), _p(row))
        out[n] = row
    return out, err


def importance_weights(spec, X, flg, wght, mean, chol, beta=1.0, df=-1, nthreads=0):
    X, wght, mean, chol = f64(X), f64(wght), f64(mean), f64(chol)
    N, d = X.shape
    flg = np.ascontiguousarray(flg, dtype=np.int16).copy()
    logw = np.zeros(N)
    mx = C.c_double()
    nok = lib().orc_importance_weights(C.byref(spec.t), N, _p(X), len(wght), d, df, _p(wght),
                                       _p(mean), _p(chol), beta, _p(flg), _p(logw),
                                       C.byref(mx), nthreads)
    return logw, flg, mx.value, nok


def normalize_weights(logw, flg, maxW):
    w = f64(logw).copy()
    ls = C.c_double()
    s = lib().orc_normalize_weights(len(w), _p(flg), _p(w), maxW, C.byref(ls))
    return w, s, ls.value


def perplexity_and_ess(wbar, flg):
    ess = C.c_double()
    p = lib().orc_perplexity_and_ess(len(wbar), _p(flg), _p(f64(wbar)), C.byref(ess))
    return p, ess.value


def update_prop_rb(X, idx, flg, wbar, wght, mean, chol, df=-1):
    X, wbar = f64(X), f64(wbar)
    wght, mean, chol = f64(wght).copy(), f64(mean).copy(), f64(chol).copy()
    N, d = X.shape
    K = len(wght)
    cov = np.zeros((K, d, d))
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    flg = np.ascontiguousarray(flg, dtype=np.int16)
    nd = lib().orc_update_prop_rb(N, _p(X), _p(idx), _p(flg), _p(wbar), K, d, df, _p(wght),
                                  _p(mean), _p(chol), _p(cov))
    return wght, mean, chol, cov, nd


def iteration(spec, N, seed, it, beta, wght, mean, chol, df=-1, nthreads=0, mode_b=False):
    wght, mean, chol = f64(wght).copy(), f64(mean).copy(), f64(chol).copy()
    K, d = mean.shape
    X = np.empty((N, d))
    idx = np.empty(N, dtype=np.int32)
    flg = np.empty(N, dtype=np.int16)
    w = np.empty(N)
    st = A.Stats()
    fn = lib().orc_iteration_mode_b if mode_b else lib().orc_iteration
    rc = fn(C.byref(spec.t), N, seed, it, beta, K, d, df, _p(wght), _p(mean),
                             _p(chol), _p(X), _p(idx), _p(flg), _p(w), C.byref(st), nthreads)
    return dict(rc=rc, wght=wght, mean=mean, chol=chol, X=X, idx=idx, flg=flg, w=w,
                stats=st.as_dict())
