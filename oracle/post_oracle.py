"""CPU oracle of the weighted post-processing of a PMC sample.

TEST INFRASTRUCTURE ONLY (see pmc_oracle.h): imported by tests/ only.

numpy restatement of the reference's in-tree host code,
    mean_from_psim / estimate_param_covar_weight   call sites exec/exec_helper.c:79, :332 (pmclib)
    median_from_psim                               exec/exec_helper.c:164-199
    sigma_from_psim                                exec/exec_helper.c:201-275
    acc_histogram                                  tools/src/nhist.c:87-162
PINNED: unlike the PMC iteration (pmc_oracle.c, parity unpinned), these functions exist in the
reference tree, so `ref()` below loads the reference's own compiled code
(oracle/_ref/libref_post.so, recipe oracle/build_ref_post.py) and tests/test_post_cpu.py checks
this restatement against it and against tests/golden/post_ref.json (generated from it).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libref_post.so")
CONF_123_HALF = (0.6827 / 2.0, 0.9545 / 2.0, 0.9973 / 2.0)    # include/pmctools/maths.h conf_68/95/99


def _flagged(X, w, flg, a):
    X = np.asarray(X, dtype=np.float64)
    m = np.ones(len(X), bool) if flg is None else (np.asarray(flg) != 0)
    ww = np.ones(len(X)) if w is None else np.asarray(w, dtype=np.float64)
    return X[m, a], ww[m]


def moments(X, w=None, flg=None):
    """weighted mean (mean_from_psim) and covariance about it (estimate_param_covar_weight:
    second pass centred on the first pass's mean, normalised by the sum of weights)"""
    X = np.asarray(X, dtype=np.float64)
    m = np.ones(len(X), bool) if flg is None else (np.asarray(flg) != 0)
    ww = (np.ones(len(X)) if w is None else np.asarray(w, dtype=np.float64))[m]
    Xf = X[m]
    s = ww.sum()
    mean = (ww[:, None] * Xf).sum(0) / s
    dX = Xf - mean
    cov = (ww[:, None, None] * dX[:, :, None] * dX[:, None, :]).sum(0) / s
    return mean, cov


def sigma(X, w, flg, a, center, conf=CONF_123_HALF):
    """sigma_from_psim: sort the flagged (x_a, w) pairs; from the first element >= center walk
    right (left) adding weights while the sum is <= conf[j]; report the distance of the NEXT
    element from center, -1 if the sample ends first."""
    par, ww = _flagged(X, w, flg, a)
    o = np.argsort(par, kind="stable")
    par, ww = par[o], ww[o]
    n = len(par)
    out = np.full(6, -1.0)
    if n == 0:
        return out
    i = 0
    while par[i] < center:           # exec_helper.c:221-225
        i += 1
        if i == n - 1:
            break
    imean = i
    for j in range(3):
        s, i = 0.0, imean
        while s <= conf[j] and i < n:
            s += ww[i]
            i += 1
        out[j] = -1.0 if i == n else par[i] - center
        s, i = 0.0, imean - 1
        while s <= conf[j] and i >= 0:
            s += ww[i]
            i -= 1
        out[3 + j] = -1.0 if i == -1 else center - par[i]
    return out


def median(X, w, flg, a):
    """median_from_psim: first sorted element where the running weight reaches 0.5"""
    par, ww = _flagged(X, w, flg, a)
    o = np.argsort(par, kind="stable")
    par, ww = par[o], ww[o]
    m = 0.0
    for i in range(len(par)):
        m += ww[i]
        if m == 0.5:
            return par[i]
        if m > 0.5:
            return 0.5 * (par[i] + par[i - 1]) if i > 0 else par[i]
    return float("nan")


def histogram(X, w, flg, pidx, nbins, limits):
    """acc_histogram on a fresh histogram: per bin (count, sum w, sum w^2) and the reference's
    data[] = sum w / nsamples, var[] = sum (w - data)^2 / (nsamples (nsamples - 1)); the last
    axis runs fastest; samples on or outside a limit are dropped."""
    X = np.asarray(X, dtype=np.float64)
    N = len(X)
    m = np.ones(N, bool) if flg is None else (np.asarray(flg) != 0)
    ww = np.ones(N) if w is None else np.asarray(w, dtype=np.float64)
    nd = len(pidx)
    tdim = int(np.prod(nbins))
    pos = np.zeros(N, dtype=np.int64)
    mul = 1
    valid = m.copy()
    for ap in range(nd):
        ip = nd - ap - 1
        lo, hi = limits[2 * ip], limits[2 * ip + 1]
        stp = (hi - lo) / nbins[ip]
        vp = X[:, pidx[ip]]
        valid &= ~((vp <= lo) | (vp >= hi))
        with np.errstate(invalid="ignore"):
            nb = np.where(valid, (vp - lo) / stp, 0.0).astype(np.int64)
        nb = np.where(nb == nbins[ip], nb - 1, nb)
        pos += nb * mul
        mul *= nbins[ip]
    pos, wv = pos[valid], ww[valid]
    count = np.bincount(pos, minlength=tdim).astype(np.float64)
    sumw = np.bincount(pos, weights=wv, minlength=tdim)
    sumw2 = np.bincount(pos, weights=wv * wv, minlength=tdim)
    return count, sumw, sumw2


def hist_data_var(count, sumw, sumw2, nsamples):
    """the reference's data[] and var[] from the three per-bin sums (nhist.c:104-108,131,157)"""
    data = sumw / nsamples
    var = (sumw2 - 2.0 * data * sumw + count * data * data) / ((nsamples - 1.0) * nsamples)
    return data, var


# ---- the reference's own compiled code (oracle/_ref) ------------------------------------------
class _NdHist(C.Structure):      # tools/include/nhist.h:20-31
    _fields_ = [("ndim", C.c_size_t), ("tdim", C.c_size_t), ("total", C.c_double), ("volume", C.c_double),
                ("nsamples", C.c_double), ("isLog", C.c_int), ("nbins", C.POINTER(C.c_size_t)),
                ("limits", C.POINTER(C.c_double)), ("stps", C.POINTER(C.c_double)), ("data", C.POINTER(C.c_double)),
                ("var", C.POINTER(C.c_double)), ("buf", C.c_void_p), ("lvol", C.POINTER(C.c_double))]


_ref = None


def ref():
    """reference functions, or None where oracle/_ref was not built (no /root/reference)"""
    global _ref
    if _ref is None and os.path.exists(REF_LIB):
        L = C.CDLL(REF_LIB)
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        L.sigma_from_psim.restype = None
        L.sigma_from_psim.argtypes = [vp, vp, vp, i, i, i, d, vp, vp, vp]
        L.median_from_psim.restype = d
        L.median_from_psim.argtypes = [vp, vp, vp, i, i, i, vp]
        L.init_nd_histogram.restype = C.POINTER(_NdHist)
        L.init_nd_histogram.argtypes = [C.c_size_t, vp, vp, vp]
        L.acc_histogram.restype = None
        L.acc_histogram.argtypes = [C.c_size_t, C.c_size_t, vp, vp, vp, C.POINTER(_NdHist), vp]
        _ref = L
    return _ref


def _arrs(X, w, flg):
    X = np.ascontiguousarray(X, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    flg = np.ascontiguousarray(flg, dtype=np.int16)
    return X, w, flg


def ref_sigma(X, w, flg, a, center, conf=CONF_123_HALF):
    L = ref()
    X, w, flg = _arrs(X, w, flg)
    out = np.zeros(6)
    cf = np.array(conf, dtype=np.float64)
    err = C.c_void_p(None)
    L.sigma_from_psim(X.ctypes.data, w.ctypes.data, flg.ctypes.data, len(X), X.shape[1], a, center,
                      out.ctypes.data, cf.ctypes.data, C.addressof(err))
    assert not err.value, "reference raised an error"
    return out


def ref_median(X, w, flg, a):
    L = ref()
    X, w, flg = _arrs(X, w, flg)
    err = C.c_void_p(None)
    v = L.median_from_psim(X.ctypes.data, w.ctypes.data, flg.ctypes.data, len(X), X.shape[1], a, C.addressof(err))
    assert not err.value, "reference raised an error"
    return v


def ref_histogram(X, w, pidx, nbins, limits):
    """acc_histogram of the reference on a fresh nd_histogram: returns data[], var[], total"""
    L = ref()
    X = np.ascontiguousarray(X, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    nb = np.array(nbins, dtype=np.uintp)
    lim = np.array(limits, dtype=np.float64)
    pi = np.array(pidx, dtype=np.uintp)
    err = C.c_void_p(None)
    h = L.init_nd_histogram(len(pidx), nb.ctypes.data, lim.ctypes.data, C.addressof(err))
    assert not err.value
    L.acc_histogram(X.shape[1], len(X), X.ctypes.data, w.ctypes.data, pi.ctypes.data, h, C.addressof(err))
    assert not err.value, "reference raised an error"
    tdim = h.contents.tdim
    data = np.array([h.contents.data[i] for i in range(tdim)])
    var = np.array([h.contents.var[i] for i in range(tdim)])
    return data, var, h.contents.total
