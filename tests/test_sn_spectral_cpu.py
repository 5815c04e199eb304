"""CPU pin of the spectral SN quadrature (cosmopmc_b200/csrc/sn_spectral.cuh, tables built in pmcb200.cu build_sn):
a numpy restatement of the construction -- Chebyshev points of [a(z_max), 1], the reference's stage-5 Romberg
functional W[z][m] and its error functional D[z][m] applied to T_m, the two per-sample certificates -- against the
oracle's node-by-node NR qromb (orc_w) at every redshift of the Union sample.

Checked: (1) a certified sample's distances agree with the oracle's Romberg to 1e-12 (the Romberg truncation error,
~1e-7, is reproduced, not removed); (2) certified => the oracle stops at stage 5 at every redshift (the certificate is
a sufficient condition for the reference's stopping rule); (3) samples the oracle integrates beyond stage 5 are never
certified."""
import ctypes as C

import numpy as np
import pytest

from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T
from oracle import oracle_lib as O

M = 28
TAIL_TOL, EPS = 1.0e-12, 1.0e-6
RW = [3937.0 / 103275.0, 3062.0 / 80325.0, 27728.0 / 722925.0, 22016.0 / 722925.0, 65536.0 / 722925.0]
RD = [-31.0 / 206550.0, -73.0 / 481950.0, -67.0 / 722925.0, -424.0 / 722925.0, 256.0 / 722925.0]


def romberg5_nodes(a):
    """nodes / weights of stages 1..5 of NR qromb on [a, 1] (trapzd: x = a + del/2, x += del)"""
    h = 1.0 - a
    xs, w, e = [a, 1.0], [0.5 * RW[0], 0.5 * RW[0]], [0.5 * RD[0], 0.5 * RD[0]]
    for j in range(1, 5):
        it = 1 << (j - 1)
        dl = h / it
        x = a + 0.5 * dl
        for _ in range(it):
            xs.append(x); w.append(RW[j]); e.append(RD[j])
            x += dl
    return np.array(xs), h * np.array(w), np.array(e)


def build_tables(zs):
    L = np.longdouble
    az = 1.0 / (1.0 + zs)
    alo = min(az.min(), 0.999)
    mid, half = L(0.5) * (1 + L(alo)), L(0.5) * (1 - L(alo))
    j = np.arange(M, dtype=L)
    nodes = np.asarray(mid + half * np.cos(np.pi * (j + L(0.5)) / M), dtype=np.float64)
    dct = np.array([[(1.0 if m == 0 else 2.0) / M * np.cos(np.pi * m * (jj + 0.5) / M) for jj in range(M)] for m in range(M)])
    W, dmax = np.zeros((len(zs), M)), np.zeros(M)
    for k, a in enumerate(az):
        x, w, e = romberg5_nodes(a)
        th = np.arccos(np.clip((x.astype(L) - mid) / half, -1, 1))
        Tm = np.cos(np.outer(np.arange(M, dtype=L), th)) / np.sqrt(x.astype(L))[None, :]
        W[k] = np.asarray(Tm @ w.astype(L), dtype=np.float64)
        dmax = np.maximum(dmax, np.abs(np.asarray(Tm @ e.astype(L), dtype=np.float64)))
    return nodes, dct, W, dmax


def spectral_distances(tab, Om, Ode, w0, w1):
    """(certified, w[z]) the way k_like_sn_spec forms them: q = (a^3 E^2 / Ode)^-1/2 at the nodes, c = DCT q, ss = W c"""
    nodes, dct, W, dmax = tab
    OK = 1.0 - Om - Ode
    Q = Om / Ode + OK / Ode * nodes + nodes ** (-3.0 * (w0 + w1)) * np.exp(-3.0 * w1 * (1.0 - nodes))
    if not (Ode > 0 and np.all(Q > 0)):
        return False, None
    q = 1.0 / np.sqrt(Q)
    c = dct @ q
    tail = np.abs(c[-3:]).sum()
    ok = tail <= TAIL_TOL * abs(c[0]) and float(dmax @ np.abs(c)) <= 0.25 * EPS * q.min()
    return bool(ok), 2997.92458 / np.sqrt(Ode) * (W @ c)


@pytest.fixture(scope="module")
def tab():
    tabz, _, _ = T.load_sn_table(T.SN_FIXTURE)
    zs = np.unique(tabz[:, 0])
    return zs, build_tables(zs)


def oracle_distances(zs, Om, Ode, w0, w1):
    lib = O.lib()
    c = A.Cosmo(Omega_m=Om, Omega_de=Ode, w0_de=w0, w1_de=w1, h_100=0.7, Omega_b=0.045, Omega_nu_mass=0.0,
                Neff_nu_mass=0.0, de_param=A.DE["linder"])
    w, st, bad = np.zeros(len(zs)), np.zeros(len(zs), dtype=int), False
    for k, z in enumerate(zs):
        ns, e = C.c_int(0), C.c_int(0)
        w[k] = lib.orc_w(C.byref(c), 1.0 / (1.0 + z), 0, C.byref(ns), C.byref(e))
        st[k], bad = ns.value, bad or e.value != 0
    return w, st, bad


def test_spectral_distances_match_the_oracle_romberg(tab):
    zs, tb = tab
    rng = np.random.default_rng(5)
    w_, mean, cov = T.proposal_sn(10)
    cases = []
    for _ in range(60):       # the benchmark proposal (flat wCDM)
        x = rng.multivariate_normal(mean[rng.integers(10)], cov[0])
        cases.append((x[0], 1.0 - x[0], x[1], 0.0))
    for _ in range(60):       # curved, w0-wa
        cases.append((rng.uniform(0.1, 0.6), rng.uniform(0.4, 1.1), rng.uniform(-2.0, -0.5), rng.uniform(-1.0, 0.8)))
    for _ in range(60):       # the whole prior box of Demo/MC_Demo/SN
        Om = rng.uniform(0.0, 1.2)
        cases.append((Om, 1.0 - Om, rng.uniform(-3.5, 0.5), 0.0))
    ncert = 0
    for Om, Ode, w0, w1 in cases:
        ok, ws = spectral_distances(tb, Om, Ode, w0, w1)
        if ws is None:
            continue
        wo, st, bad = oracle_distances(zs, Om, Ode, w0, w1)
        if ok:
            ncert += 1
            assert not bad
            assert np.all(st == 5), (Om, Ode, w0, w1, st.max())          # the stopping-rule certificate holds
            assert np.max(np.abs(ws / wo - 1.0)) < 1e-12, (Om, Ode, w0, w1)
        if (not bad) and st.max() > 5:
            assert not ok                                                 # deeper Romberg stages are never certified
    assert ncert >= 100, ncert        # the certificates are not vacuous: most of these samples pass


def test_romberg_truncation_is_reproduced_not_removed(tab):
    """the spectral value agrees with the 17-node Romberg value far better than either agrees with the integral"""
    from scipy.integrate import quad
    zs, tb = tab
    Om, w0 = 0.3, -1.8
    ok, ws = spectral_distances(tb, Om, 1 - Om, w0, 0.0)
    wo, st, bad = oracle_distances(zs, Om, 1 - Om, w0, 0.0)
    assert ok and not bad
    k = len(zs) - 1
    a = 1.0 / (1.0 + zs[k])
    exact = 2997.92458 * quad(lambda x: 1.0 / np.sqrt(Om * x + (1 - Om) * x ** (1 - 3 * w0)), a, 1.0, epsabs=0, epsrel=1e-13)[0]
    trunc = abs(wo[k] / exact - 1.0)
    assert 1e-11 < trunc < 1e-6
    assert abs(ws[k] / wo[k] - 1.0) < 1e-3 * trunc
