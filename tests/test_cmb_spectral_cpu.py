"""CPU pin of the spectral form of the comoving distance to a* (cosmo.cuh cmb_spec_w; reference: likeli_CMBDistPrior,
wrappers/src/wmap.c:1027-1035 -> nicaea w(a*) -> pmclib sm2_qromberg = NR qromb, 11 stages for this integrand).

The library's tables (pmcb200_cmb_spectral_tables: host only, no device needed) are applied in numpy to the integrand
1/sqrt(a^4 E^2) of the oracle and compared with the oracle's node-by-node Romberg (orc_w):
  * certified samples: the stage-11 functional reproduces orc_w to 1e-13 relative, and orc_w did stop at stage 11;
  * the certificate mirrors the reference's stopping rule: a sample for which the oracle stops at another stage is never
    certified, and the predicted |dss_11| / |ss_11| tells which;
  * the reference's value is NOT the integral (truncation 2e-5): the functional reproduces the rule, not the integral."""
import ctypes as C

import numpy as np

from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T

EPS = 1.0e-6        # ROMB_EPS / ORC_ROMB_EPS
R_H = 2997.92458


def tables():
    lib = A.load_library()
    nrow = C.c_int(0)
    m = lib.pmcb200_cmb_spectral_tables(None, None, C.byref(nrow))
    tk = np.zeros(m); th = np.zeros((nrow.value, m))
    assert lib.pmcb200_cmb_spectral_tables(tk.ctypes.data, th.ctypes.data, None) == m
    return tk, th


def certified(th, fk, tail_tol=5.0e-12):
    acc = th @ fk
    ok = abs(acc[15]) + abs(acc[16]) + abs(acc[17]) <= tail_tol * abs(acc[14]) and acc[0] > 0
    ok = ok and abs(acc[1]) <= (1 - 1e-4) * EPS * abs(acc[0])
    for j in range(6):
        ok = ok and abs(acc[8 + j]) > (1 + 1e-4) * EPS * abs(acc[2 + j])
    return ok, acc


def test_tables_shape_and_nodes():
    tk, th = tables()
    assert len(tk) == 56 and th.shape == (18, 56)
    assert np.all(tk > 0) and np.all(tk < 1) and np.all(np.diff(tk) < 0)       # Chebyshev points of v = ln(t + tau), descending
    assert abs(th[0].sum() - 1.0) < 1e-12          # the rule integrates a constant exactly: sum of the weights = interval length
    assert abs(th[1].sum()) < 1e-12                # and its error estimate of a constant is zero
    assert abs(th[14].sum() - 1.0) < 1e-12         # c_0 of a constant


def test_functional_reproduces_the_reference_romberg(oracle):
    L = oracle.lib()
    L.orc_z_star.restype = C.c_double; L.orc_z_star.argtypes = [C.POINTER(A.Cosmo)]
    tk, th = tables()
    spec = T.target_cmb_bao_sn()
    c0 = A.Cosmo.from_buffer_copy(bytes(spec.t.like[0].model))
    rng = np.random.default_rng(11)
    ncert = nother = 0
    worst = 0.0
    for trial in range(60):
        cc = A.Cosmo.from_buffer_copy(bytes(c0))
        wide = trial >= 30
        cc.Omega_m = 0.27 + (0.06 if wide else 0.02) * rng.normal()
        cc.Omega_de = 0.73 + (0.06 if wide else 0.02) * rng.normal()
        cc.h_100 = 0.71 + (0.05 if wide else 0.02) * rng.normal()
        cc.w0_de = -1.0 + (0.3 if wide else 0.1) * rng.normal()
        cc.Omega_b = 0.045 + 0.003 * rng.normal()
        if not (cc.Omega_m > 0.05 and cc.Omega_de > 0.05 and cc.h_100 > 0.3 and cc.w0_de < -0.4):
            continue
        a0 = 1.0 / (1.0 + L.orc_z_star(C.byref(cc)))
        ns, er = C.c_int(0), C.c_int(0)
        w_ref = L.orc_w(C.byref(cc), a0, 1, C.byref(ns), C.byref(er))
        assert er.value == 0
        fk = np.array([1.0 / np.sqrt(a ** 4 * L.orc_Esqr(C.byref(cc), float(a), 1)) for a in a0 + (1.0 - a0) * tk])
        ok, acc = certified(th, fk)
        if ok:
            ncert += 1
            assert ns.value == 11, (trial, ns.value)                      # certified => the reference stopped at stage 11
            worst = max(worst, abs(R_H * (1.0 - a0) * acc[0] / w_ref - 1.0))
        else:
            nother += 1
        if ns.value != 11:
            assert not ok
            # the reason is visible in the predicted stopping ratios
            r11 = abs(acc[1]) / abs(acc[0]); r10 = abs(acc[13]) / abs(acc[7])
            assert (ns.value > 11 and r11 > (1 - 1e-4) * EPS) or (ns.value < 11 and r10 <= (1 + 1e-4) * EPS), (ns.value, r10, r11)
    assert ncert >= 30 and worst < 1e-13, (ncert, nother, worst)


def test_the_rule_is_reproduced_not_the_integral(oracle):
    L = oracle.lib()
    L.orc_z_star.restype = C.c_double; L.orc_z_star.argtypes = [C.POINTER(A.Cosmo)]
    tk, th = tables()
    spec = T.target_cmb_bao_sn()
    cc = A.Cosmo.from_buffer_copy(bytes(spec.t.like[0].model))
    a0 = 1.0 / (1.0 + L.orc_z_star(C.byref(cc)))
    f = lambda a: 1.0 / np.sqrt(a ** 4 * L.orc_Esqr(C.byref(cc), float(a), 1))
    xs, ws = np.polynomial.legendre.leggauss(200)
    u = 0.5 * (0.0 - np.log(a0)) * xs + 0.5 * np.log(a0)
    exact = 0.5 * (0.0 - np.log(a0)) * np.sum(ws * np.array([np.exp(t) * f(np.exp(t)) for t in u]))
    fk = np.array([f(a) for a in a0 + (1.0 - a0) * tk])
    ss11 = (1.0 - a0) * (th[0] @ fk)
    ns, er = C.c_int(0), C.c_int(0)
    w_ref = L.orc_w(C.byref(cc), a0, 1, C.byref(ns), C.byref(er)) / R_H
    assert abs(ss11 / w_ref - 1.0) < 1e-13
    assert 5e-6 < abs(w_ref / exact - 1.0) < 1e-4        # the reference's own truncation error, reproduced


def _polint0(xa, ya):
    """Numerical Recipes polint at x = 0 (as oracle/pmc_oracle.c polint0): value and last correction"""
    n = len(xa)
    c, d = list(ya), list(ya)
    ns = int(np.argmin(np.abs(xa)))
    y = ya[ns]; ns -= 1
    dy = 0.0
    for m in range(1, n):
        for i in range(n - m):
            ho, hp = xa[i], xa[i + m]
            w = c[i + 1] - d[i]
            den = w / (ho - hp)
            d[i] = hp * den
            c[i] = ho * den
        if 2 * (ns + 1) < n - m:
            dy = c[ns + 1]
        else:
            dy = d[ns]; ns -= 1
        y += dy
    return y, dy


def test_every_stage_functional_against_a_node_by_node_tableau(oracle):
    """Rows 0..13 of the table (value and error estimate of stages 5..11) against NR's trapzd / polint tableau run node
    by node in numpy on the same integrand: the stage values to 1e-13, the error estimates to 1e-7 of themselves (they are
    differences six orders below the value)."""
    L = oracle.lib()
    L.orc_z_star.restype = C.c_double; L.orc_z_star.argtypes = [C.POINTER(A.Cosmo)]
    tk, th = tables()
    spec = T.target_cmb_bao_sn()
    cc = A.Cosmo.from_buffer_copy(bytes(spec.t.like[0].model))
    cc.Omega_m, cc.w0_de = 0.29, -0.93
    a0 = 1.0 / (1.0 + L.orc_z_star(C.byref(cc)))
    f = lambda a: 1.0 / np.sqrt(a ** 4 * L.orc_Esqr(C.byref(cc), float(a), 1))
    h = 1.0 - a0
    # NR trapzd stages 1..11
    s, hh = [], [1.0]
    st = 0.5 * h * (f(a0) + f(1.0))
    s.append(st)
    for j in range(1, 11):
        it = 1 << (j - 1)
        dl = h / it
        x = a0 + 0.5 * dl
        tot = 0.0
        for _ in range(it):
            tot += f(x); x += dl
        st = 0.5 * (st + h * tot / it)
        s.append(st); hh.append(0.25 * hh[-1])
    fk = np.array([f(a) for a in a0 + h * tk])
    acc = h * (th @ fk)
    for j in range(5, 12):            # stage j uses s[j-5 .. j-1]
        ss, dss = _polint0(np.array(hh[j - 5:j]), s[j - 5:j])
        row_s = 0 if j == 11 else 2 + (j - 5)
        row_d = 1 if j == 11 else 8 + (j - 5)
        assert abs(acc[row_s] / ss - 1.0) < 1e-13, (j, acc[row_s], ss)
        assert abs(acc[row_d] / dss - 1.0) < 1e-7, (j, acc[row_d], dss)
