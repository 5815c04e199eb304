#!/usr/bin/env python
"""Prototype (CPU, numpy): the reference's Romberg value of the comoving distance to a* (11 stages, 1025 equidistant
nodes in a) as a fixed linear functional, applied to a Chebyshev interpolant of the integrand in ln a.
Questions: (1) does the functional reproduce orc_w to rounding?  (2) how many Chebyshev points for 1e-13?
(3) how far are the stopping tests of stages 5..11 from their threshold?"""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle_lib as O
from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T

L = O.lib()
K = 5
x = np.array([4.0 ** (-q) for q in range(K)], dtype=np.longdouble)      # h_k ratios (only ratios matter)


def lagr0(xs):
    lam = np.ones(len(xs), dtype=np.longdouble)
    for k in range(len(xs)):
        for m in range(len(xs)):
            if m != k:
                lam[k] *= (0 - xs[m]) / (xs[k] - xs[m])
    return lam


lam5 = lagr0(x)                       # ss  = sum lam5[q] s_{j-4+q}
lam4 = np.concatenate([[0], lagr0(x[1:])])
mu5 = lam5 - lam4                     # dss = P_{0..4}(0) - P_{1..4}(0)


def stage_weights(j, jmax):
    """weights on the 2^(jmax-1)+1 nodes of stage jmax for ss_j, dss_j (j <= jmax), unit interval length"""
    n = 2 ** (jmax - 1)
    ws, wd = np.zeros(n + 1, dtype=np.longdouble), np.zeros(n + 1, dtype=np.longdouble)
    for q in range(K):
        k = j - 4 + q                  # trapezoid stage k: 2^(k-1) intervals
        nk = 2 ** (k - 1); step = n // nk
        t = np.zeros(n + 1, dtype=np.longdouble)
        t[::step] = 1.0 / nk; t[0] *= 0.5; t[-1] *= 0.5
        ws += lam5[q] * t; wd += mu5[q] * t
    return ws, wd


def model(rng, c0):
    cc = A.Cosmo.from_buffer_copy(bytes(c0))
    cc.Omega_m = 0.27 + 0.02 * rng.normal(); cc.Omega_de = 0.73 + 0.02 * rng.normal(); cc.h_100 = 0.71 + 0.02 * rng.normal()
    cc.w0_de = -1 + 0.1 * rng.normal(); cc.Omega_b = 0.045 + 0.003 * rng.normal()
    return cc


spec = T.target_cmb_bao_sn()
c0 = A.Cosmo.from_buffer_copy(bytes(spec.t.like[0].model))
rng = np.random.default_rng(2)
JM = 11
W = {j: stage_weights(j, JM) for j in range(5, JM + 1)}
for trial in range(8):
    cc = model(rng, c0)
    zs = L.orc_z_star(C.byref(cc)) if hasattr(L, "orc_z_star") else 1091.0
    L.orc_z_star.restype = C.c_double; L.orc_z_star.argtypes = [C.POINTER(A.Cosmo)]
    zs = L.orc_z_star(C.byref(cc))
    a0 = 1.0 / (1.0 + zs)
    ns, er = C.c_int(0), C.c_int(0)
    w_ref = L.orc_w(C.byref(cc), a0, 1, C.byref(ns), C.byref(er)) / 2997.92458
    f = lambda a: 1.0 / np.sqrt(a ** 4 * L.orc_Esqr(C.byref(cc), float(a), 1))
    n = 2 ** (JM - 1)
    a = a0 + (1.0 - a0) * np.arange(n + 1) / n
    fv = np.array([f(t) for t in a], dtype=np.longdouble)
    h = np.longdouble(1.0 - a0)
    ratios = []
    for j in range(5, JM + 1):
        ss, dss = h * np.sum(W[j][0] * fv), h * np.sum(W[j][1] * fv)
        ratios.append(float(abs(dss) / abs(ss)))
    ss11 = float(h * np.sum(W[JM][0] * fv))
    line = "z*=%.2f stages %d  functional/ref-1 = %.2e  |dss/ss| j=5..11: %s" % (zs, ns.value, ss11 / w_ref - 1, " ".join("%.1e" % r for r in ratios))
    # Chebyshev interpolant in u = ln a
    res = []
    for M in (24, 28, 32, 36, 40, 48):
        jn = np.arange(M)
        xc = np.cos(np.pi * (jn + 0.5) / M)
        lo, hi = np.log(a0), 0.0
        u = 0.5 * (hi - lo) * xc + 0.5 * (hi + lo)
        g = np.array([f(np.exp(t)) for t in u])
        cf = np.polynomial.chebyshev.chebfit(xc, g, M - 1)
        xa = (2 * np.log(a.astype(np.float64)) - (hi + lo)) / (hi - lo)
        gi = np.polynomial.chebyshev.chebval(xa, cf)
        ss_i = float(h * np.sum(W[JM][0] * gi.astype(np.longdouble)))
        res.append("M=%d %.1e (tail %.0e)" % (M, ss_i / ss11 - 1, abs(cf[-3:]).sum() / abs(cf[0])))
    print(line); print("    interpolant: " + "  ".join(res))


# ---- value-space functionals on fixed nodes in v = ln(t + tau), t = (a - a*) / (1 - a*) ------------------------------
def theta_tables(M, tau, JM=11):
    ld = np.longdouble
    vlo, vhi = np.log(ld(tau)), np.log(ld(1) + ld(tau))
    k = np.arange(M)
    xk = np.cos(np.pi * (k + ld(0.5)) / M).astype(ld)
    vk = (vhi - vlo) / 2 * xk + (vhi + vlo) / 2
    tk = np.exp(vk) - ld(tau)
    n = 2 ** (JM - 1)
    ti = np.arange(n + 1, dtype=ld) / n
    xi = (2 * np.log(ti + ld(tau)) - (vhi + vlo)) / (vhi - vlo)
    xi = np.clip(xi, -1, 1)
    th = np.arccos(xi)
    m = np.arange(M)
    B = np.cos(np.outer(th, m)).astype(ld)                 # T_m(x_i)
    Tk = np.cos(np.outer(np.arccos(xk), m)).astype(ld)     # T_m(x_k)
    sc = np.full(M, ld(2) / M); sc[0] = ld(1) / M
    card = (B * sc) @ Tk.T                                 # l_k(x_i): [n+1][M]
    rows = {}
    for j in range(5, JM + 1):
        ws, wd = stage_weights(j, JM)
        rows[("ss", j)] = ws @ card
        rows[("dss", j)] = wd @ card
    # Chebyshev coefficients c_0, c_(M-3..M-1) as functionals of the node values
    for mm in (0, M - 3, M - 2, M - 1):
        rows[("c", mm)] = sc[mm] * Tk[:, mm]
    return tk.astype(np.float64), {k_: v.astype(np.float64) for k_, v in rows.items()}


tau = (1 / 1091.0) / (1 - 1 / 1091.0)
rng = np.random.default_rng(5)
for M in (48, 56, 64):
    tk, rows = theta_tables(M, tau)
    worst = 0.0; worst_tail = 0.0; rmax = 0.0; rmin10 = 1.0
    for trial in range(12):
        cc = model(rng, c0)
        if trial >= 8:      # wider
            cc.Omega_m = 0.27 + 0.06 * rng.normal(); cc.w0_de = -1 + 0.3 * rng.normal(); cc.h_100 = 0.71 + 0.05 * rng.normal()
        zs = L.orc_z_star(C.byref(cc)); a0 = 1 / (1 + zs)
        ns, er = C.c_int(0), C.c_int(0)
        w_ref = L.orc_w(C.byref(cc), a0, 1, C.byref(ns), C.byref(er)) / 2997.92458
        f = lambda a: 1.0 / np.sqrt(a ** 4 * L.orc_Esqr(C.byref(cc), float(a), 1))
        fk = np.array([f(a0 + (1 - a0) * t) for t in tk])
        h = 1 - a0
        ss11 = h * rows[("ss", 11)] @ fk
        r = [abs(rows[("dss", j)] @ fk) / abs(rows[("ss", j)] @ fk) for j in range(5, 12)]
        tail = (abs(rows[("c", M - 1)] @ fk) + abs(rows[("c", M - 2)] @ fk) + abs(rows[("c", M - 3)] @ fk)) / abs(rows[("c", 0)] @ fk)
        if ns.value == 11:
            worst = max(worst, abs(ss11 / w_ref - 1)); worst_tail = max(worst_tail, tail)
        rmax = max(rmax, r[-1]); rmin10 = min(rmin10, r[-2])
        if M == 56:
            print("  z*=%.1f stages %d  ss11/ref-1 = %.1e  r10 = %.2e r11 = %.3e tail %.1e" % (zs, ns.value, ss11 / w_ref - 1, r[-2], r[-1], tail))
    print("M = %d: worst |ss11/ref - 1| = %.1e, worst tail %.1e, max r11 %.3e, min r10 %.2e, sum|theta| %.3f" % (M, worst, worst_tail, rmax, rmin10, np.abs(rows[("ss", 11)]).sum()))
