"""CPU tests of the oracle (oracle/pmc_oracle.c): pins against scipy, against the
reference's in-tree restatements (perl formulas), against documented known
answers, and against the committed golden vectors.  No GPU needed."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest
from scipy import integrate, stats

from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def cosmo(**kw):
    d = dict(T.COSMO_SN)
    d.update(kw)
    return A.Cosmo(d["Omega_m"], d["Omega_de"], d["w0_de"], d["w1_de"], d["h_100"], d["Omega_b"],
                   d["Omega_nu_mass"], d["Neff_nu_mass"], d["de_param"], 0)


def test_cholesky_and_logpdf_vs_scipy(oracle):
    rng = np.random.default_rng(0)
    for d in (1, 2, 5, 20, 32):
        Aa = rng.standard_normal((d, d))
        cov = Aa @ Aa.T + d * np.eye(d)
        ch = oracle.cholesky_stack(cov[None])[0]
        assert np.allclose(ch, np.linalg.cholesky(cov), rtol=1e-13, atol=1e-14)
        mean = rng.standard_normal(d)
        X = mean + rng.standard_normal((50, d)) * 2
        got = oracle.mix_log_pdf(X, [1.0], mean[None], ch[None])
        ref = stats.multivariate_normal(mean, cov).logpdf(X)
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
        got_t = oracle.mix_log_pdf(X, [1.0], mean[None], ch[None], df=3)
        ref_t = stats.multivariate_t(mean, cov, df=3).logpdf(X)
        assert np.allclose(got_t, ref_t, rtol=1e-12, atol=1e-12)
    assert oracle.lib().orc_cholesky(2, np.array([[1.0, 2.0], [2.0, 1.0]]).ctypes.data) != 0


def test_mixture_logpdf_semantics(oracle):
    """log sum_k alpha_k phi_k without max-shift; zero-weight components skipped; far tail -> -inf"""
    w = np.array([0.3, 0.0, 0.7])
    mean = np.array([[0.0, 0.0], [5.0, 5.0], [1.0, -1.0]])
    cov = np.stack([np.eye(2), np.eye(2) * 1e-12, np.eye(2) * 2])
    ch = oracle.cholesky_stack(cov)
    X = np.array([[0.1, 0.2], [5.0, 5.0], [1e3, 1e3]])
    got = oracle.mix_log_pdf(X, w, mean, ch)
    ref = np.log(0.3 * stats.multivariate_normal(mean[0], cov[0]).pdf(X[:2])
                 + 0.7 * stats.multivariate_normal(mean[2], cov[2]).pdf(X[:2]))
    assert np.allclose(got[:2], ref, rtol=1e-13)
    assert got[2] == -np.inf


def test_component_selection(oracle):
    w = np.array([0.25, 0.0, 0.5, 0.25])
    f = lambda u: oracle.lib().orc_select_component(4, w.ctypes.data, u)
    assert f(0.0) == 0 and f(np.nextafter(0.25, 0)) == 0
    assert f(0.25) == 2                       # u < cw strictly; dead component 1 skipped
    assert f(np.nextafter(0.75, 0)) == 2 and f(0.75) == 3
    assert f(np.nextafter(1.0, 0)) == 3 and f(1.0) == 3   # fallback: last live component


def test_philox_known_answer(oracle):
    """Random123 known-answer vectors for Philox4x32-10."""
    def ph(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        oracle.lib().orc_philox4x32_10(c, k, o)
        return list(o)
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sampler_statistics(oracle):
    spec = T.target_gauss2d()
    mean = np.array([[0.3, 0.6], [0.7, 0.2]])
    cov = np.array([[[0.01, 0.004], [0.004, 0.02]], [[0.02, -0.01], [-0.01, 0.03]]])
    ch = oracle.cholesky_stack(cov)
    X, idx, flg, nok = oracle.simulate(400000, 7, 0, 0, [0.25, 0.75], mean, ch, *spec.box)
    assert abs((idx == 1).mean() - 0.75) < 5e-3
    for k in (0, 1):
        Xk = X[idx == k]
        assert np.allclose(Xk.mean(0), mean[k], atol=2e-3)
        assert np.allclose(np.cov(Xk.T), cov[k], atol=1e-3)
    inside = np.all((X >= 0) & (X <= 1), axis=1)
    assert np.array_equal(flg != 0, inside) and nok == inside.sum()
    # 1-D normality of the whitened draws of component 0
    z = np.linalg.solve(ch[0], (X[idx == 0] - mean[0]).T).T
    assert stats.kstest(z[:, 0], "norm").pvalue > 1e-3


def test_romberg_against_quad(oracle):
    """comoving / luminosity distance against scipy.quad to the Romberg tolerance (1e-6)."""
    e = C.c_int(0)
    for kw in (dict(), dict(Omega_m=0.3, Omega_de=0.6), dict(Omega_m=0.25, Omega_de=0.9),
               dict(w0_de=-0.8, w1_de=0.3), dict(w0_de=-1.2, w1_de=-0.5, de_param=A.DE["jassal"])):
        c = cosmo(**kw)
        OK = 1 - c.Omega_m - c.Omega_de
        for z in (0.015, 0.3, 1.0, 1.55):
            a = 1 / (1 + z)
            def E(zz):
                aa = 1 / (1 + zz)
                if c.de_param == A.DE["linder"]:
                    fde = aa ** (-3 * (1 + c.w0_de + c.w1_de)) * np.exp(-3 * c.w1_de * (1 - aa))
                else:
                    fde = aa ** (-3 * (1 + c.w0_de)) * np.exp(1.5 * c.w1_de * (1 - aa) ** 2)
                return np.sqrt(c.Omega_m / aa ** 3 + OK / aa ** 2 + c.Omega_de * fde)
            w = 2997.92458 * integrate.quad(lambda zz: 1 / E(zz), 0, z, epsabs=0, epsrel=1e-12)[0]
            if abs(OK) < 1e-8:
                fk = w
            else:
                sk = np.sqrt(abs(OK)) / 2997.92458
                fk = np.sinh(sk * w) / sk if OK > 0 else np.sin(sk * w) / sk
            dl = oracle.lib().orc_D_lum(C.byref(c), a, C.byref(e))
            assert e.value == 0
            assert abs(dl / (fk * (1 + z)) - 1) < 2e-6
    # the stopping rule is NR qromb: at least 5 stages
    ns = C.c_int(0)
    oracle.lib().orc_w(C.byref(cosmo()), 0.5, 0, C.byref(ns), C.byref(e))
    assert ns.value == 5


def test_sn_likelihood_against_independent_numpy(oracle):
    """SN chi^2 re-derived in numpy from Manual/manual.tex:1290-1325 with quad distances."""
    spec = T.target_sn_demo()
    tab, sig_int, v_pec = T.load_sn_table(T.SN_FIXTURE)
    x = np.array([0.3, -0.9, 19.3, 1.4, -2.2])
    Om, w0, M, al, be = x
    z = tab[:, 0]
    E = lambda zz: np.sqrt(Om * (1 + zz) ** 3 + (1 - Om) * (1 + zz) ** (3 * (1 + w0)))
    dl = np.array([2997.92458 * integrate.quad(lambda t: 1 / E(t), 0, zi, epsrel=1e-11)[0] * (1 + zi) for zi in z])
    mu_th = 5 * np.log10(dl / 0.7) + 25
    mu_obs = tab[:, 1] + M + al * (tab[:, 2] - 1) + be * tab[:, 3]
    sig2 = (tab[:, 4] + al ** 2 * tab[:, 5] + be ** 2 * tab[:, 6]
            + 2 * (al * tab[:, 7] + be * tab[:, 8] + al * be * tab[:, 9])
            + (5 / np.log(10) * v_pec / 299792.458 / z) ** 2 + sig_int ** 2)
    logl = -0.5 * np.sum((mu_obs - mu_th) ** 2 / sig2)
    logpr = -np.sum(np.log(spec.box[1] - spec.box[0]))
    got, err = oracle.posterior_log_pdf(spec, x[None])
    assert err[0] == 0
    assert abs(got[0] - (logl + logpr)) < 2e-4 * abs(logl)      # Romberg truncation (1e-6 in D_L)


def test_golden_logposterior(oracle):
    g = json.load(open(os.path.join(HERE, "golden", "sn_logpost.json")))
    got, err = oracle.posterior_log_pdf(T.target_sn_demo(), np.array(g["x"]))
    assert err.sum() == 0
    assert np.allclose(got, g["logpost"], rtol=1e-13, atol=0)


def test_golden_test_suite_fiducials(oracle):
    """the fiducial points of bin/test_suite_cosmo_pmc.pl:51-58 for SN, BAO (A, d_z), WMAP distance priors and the
    joint SN + BAO set against the committed goldens (generator: tests/golden/make_test_suite_fixture.py)"""
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_test_suite_fixture.py"))
    mk = importlib.util.module_from_spec(spec_); spec_.loader.exec_module(mk)
    gold = json.load(open(os.path.join(HERE, "golden", "test_suite_logpost.json")))["cases"]
    for name, (spec, fid) in mk.suite().items():
        lp, err = oracle.posterior_log_pdf(spec, np.array([fid]))
        assert err[0] == 0 and fid == gold[name]["fid"]
        assert abs(lp[0] - gold[name]["logpost"]) <= 1e-12 * abs(gold[name]["logpost"]), name


def test_parameter_mapping_rules(oracle):
    """set_base_parameters (param.c:1544-1661): flat closure, physical densities, error cases."""
    spec = T.target_sn_demo()
    tab = T.load_sn_table(T.SN_FIXTURE)[0]
    # Omega_m given => Omega_de = 1 - Omega_m (flat): equals the curved target on the flat line
    curved = T.target_sn_curved()
    x = np.array([[0.31, -1.0, 19.3, 1.5, -2.0]])
    xc = np.array([[0.31, 0.69, 19.3, 1.5, -2.0]])
    a, _ = oracle.posterior_log_pdf(spec, x)
    b, _ = oracle.posterior_log_pdf(curved, xc)
    off = np.sum(np.log(curved.box[1] - curved.box[0])) - np.sum(np.log(spec.box[1] - spec.box[0]))
    assert abs((a[0]) - (b[0] + off)) < 1e-9
    # physical densities: omega_m = Omega_m h^2 with h_100 sampled
    phys = T.TargetSpec(["omega_m", "h_100", "M", "alpha", "beta"], [0.05, 0.5, 19.1, 0.5, -3.5],
                        [0.3, 0.9, 19.8, 2.6, -0.8]).add_snia()
    nonp = T.TargetSpec(["Omega_m", "h_100", "M", "alpha", "beta"], [0.05, 0.5, 19.1, 0.5, -3.5],
                        [1.0, 0.9, 19.8, 2.6, -0.8]).add_snia()
    h = 0.7
    p, e1 = oracle.posterior_log_pdf(phys, np.array([[0.3 * h * h, h, 19.3, 1.5, -2.0]]))
    q, e2 = oracle.posterior_log_pdf(nonp, np.array([[0.3, h, 19.3, 1.5, -2.0]]))
    offp = np.sum(np.log(nonp.box[1] - nonp.box[0])) - np.sum(np.log(phys.box[1] - phys.box[0]))
    assert e1[0] == 0 and e2[0] == 0 and abs(p[0] - (q[0] + offp)) < 1e-9
    # mixing physical and non-physical parameters is an error (param.c:1554-1556)
    mixed = T.TargetSpec(["Omega_m", "omega_b", "M", "alpha", "beta"], [0.05, 0.01, 19.1, 0.5, -3.5],
                         [1.0, 0.05, 19.8, 2.6, -0.8]).add_snia()
    _, e3 = oracle.posterior_log_pdf(mixed, np.array([[0.3, 0.02, 19.3, 1.5, -2.0]]))
    assert e3[0] != 0
    # overdetermined total density (param.c:1577-1578)
    over = T.TargetSpec(["Omega_m", "Omega_de", "Omega_K"], [0.05, 0.1, -0.5], [1.0, 1.0, 0.5]).add_snia()
    _, e4 = oracle.posterior_log_pdf(over, np.array([[0.3, 0.7, 0.0]]))
    assert e4[0] != 0


def test_de_conservative_prior(oracle):
    """special prior de_conservative: volume term of wrappers/src/param.c:1072-1094 and the probes'
    hard cut, which returns log L = 0 for a violating model (sn.c:263-274)."""
    lo, hi = [0.0, -3.5, 19.1, 0.5, -3.5], [1.2, 0.5, 19.8, 2.6, -0.8]
    names = ["Omega_m", "w_0_de", "M", "alpha", "beta"]
    plain = T.TargetSpec(names, lo, hi).add_snia()
    cons = T.TargetSpec(names, lo, hi).add_snia(special="de_conservative")
    X = np.array([[0.3, -0.9, 19.3, 1.5, -2.0], [0.3, -1.2, 19.3, 1.5, -2.0], [0.3, -0.2, 19.3, 1.5, -2.0]])
    p0, e0 = oracle.posterior_log_pdf(plain, X)
    p1, e1 = oracle.posterior_log_pdf(cons, X)
    assert not e0.any() and not e1.any()
    vol = np.log(hi[1] - lo[1]) - np.log(2.0 / 3.0)
    logpr = -np.sum(np.log(np.array(hi) - np.array(lo)))
    assert abs((p1[0] - p0[0]) - vol) < 1e-12                  # inside [-1, -1/3]: only the volume changes
    assert np.allclose(p1[1:], logpr + vol, rtol=0, atol=1e-12)   # outside: log L = 0
    # w0-wa: both w(1) and w(a_acc = 2/3) must lie inside; two-triangle volume as written in the reference
    names2 = ["Omega_m", "w_0_de", "w_1_de", "M", "alpha", "beta"]
    lo2, hi2 = [0.0, -2.0, -3.0, 19.1, 0.5, -3.5], [1.2, 0.0, 2.0, 19.8, 2.6, -0.8]
    cons2 = T.TargetSpec(names2, lo2, hi2).add_snia(special="de_conservative")
    plain2 = T.TargetSpec(names2, lo2, hi2).add_snia()
    X2 = np.array([[0.3, -0.9, 0.6, 19.3, 1.5, -2.0],      # w(2/3) = -0.7: inside
                   [0.3, -0.9, 2.0, 19.3, 1.5, -2.0],      # w(2/3) = -0.233: outside
                   [0.3, -0.5, -1.8, 19.3, 1.5, -2.0]])    # w(2/3) = -1.1: outside
    q0, _ = oracle.posterior_log_pdf(plain2, X2)
    q1, _ = oracle.posterior_log_pdf(cons2, X2)
    vol2 = np.log(2.0) + np.log(5.0) - np.log(0.5 * 4.0 / 9.0 / (1.0 - 2.0 / 3.0)) - np.log(0.5 * 4.0 / 9.0)
    logpr2 = -np.sum(np.log(np.array(hi2) - np.array(lo2)))
    assert abs((q1[0] - q0[0]) - vol2) < 1e-12
    assert np.allclose(q1[1:], logpr2 + vol2, rtol=0, atol=1e-12)
    # a w0 range narrower than the prior is refused (param.c:1081-1083)
    bad = T.TargetSpec(names, [0.0, -0.8, 19.1, 0.5, -3.5], hi).add_snia(special="de_conservative")
    _, eb = oracle.posterior_log_pdf(bad, X[:1])
    assert eb.all()


def test_bao_cmb_sanity(oracle):
    """BAO / CMB distance-prior model values at the WMAP7 best fit are close to the data."""
    spec = T.TargetSpec(["Omega_m", "Omega_de", "h_100", "Omega_b"], [0.1, 0.3, 0.5, 0.02],
                        [0.6, 1.1, 0.9, 0.08]).add_cmbdp()
    x = np.array([[0.272, 0.728, 0.704, 0.0456]])
    lp, err = oracle.posterior_log_pdf(spec, x)
    assert err[0] == 0
    # Gaussian normalisation of the 3-d data + box prior; chi^2 at the best fit must be small
    cov = np.linalg.inv(np.array(T.WMAP7_DP["covinv"]))
    norm = -0.5 * (3 * np.log(2 * np.pi) + np.linalg.slogdet(cov)[1]) - np.sum(np.log(spec.box[1] - spec.box[0]))
    chi2 = -2 * (lp[0] - norm)
    assert 0 <= chi2 < 25
    spec2 = T.TargetSpec(["Omega_m", "Omega_de"], [0.05, 0.2], [0.8, 1.2]).add_bao(T.BAO_REID10_A)
    lp2, err2 = oracle.posterior_log_pdf(spec2, np.array([[0.27, 0.73], [0.6, 0.4]]))
    assert err2.sum() == 0 and lp2[0] > lp2[1]


def test_bao_D_V_ratio_against_quad(oracle):
    """distance_D_V_ratio (bao.c:56,170-171): D_V(z_2i) / D_V(z_2i+1), D_V = [f_K^2 c z / H]^(1/3)
    (Manual/manual.tex:1707-1735), against scipy quadrature for a curved w0 model."""
    from scipy.integrate import quad
    data = dict(method="distance_D_V_ratio", mean=[1.736, 1.52], covinv=[[260.0, -40.0], [-40.0, 400.0]],
                z=[0.35, 0.2, 0.57, 0.35])
    spec = T.TargetSpec(["Omega_m", "Omega_de", "w_0_de"], [0.05, 0.2, -2.0], [0.8, 1.2, -0.4]).add_bao(data)
    x = np.array([[0.3, 0.6, -0.9], [0.25, 0.85, -1.2]])
    lp, err = oracle.posterior_log_pdf(spec, x)
    assert err.sum() == 0
    cov = np.linalg.inv(np.array(data["covinv"]))
    for i, (Om, Ode, w0) in enumerate(x):
        OK = 1.0 - Om - Ode
        E = lambda z: np.sqrt(Om * (1 + z) ** 3 + OK * (1 + z) ** 2 + Ode * (1 + z) ** (3 * (1 + w0)))
        def DV(z):
            w = quad(lambda t: 1.0 / E(t), 0.0, z, epsabs=1e-13, epsrel=1e-13)[0]
            sk = np.sqrt(abs(OK))
            fk = np.sinh(sk * w) / sk if OK > 0 else np.sin(sk * w) / sk
            return (fk * fk * z / E(z)) ** (1.0 / 3.0)
        model = np.array([DV(0.35) / DV(0.2), DV(0.57) / DV(0.35)])
        r = model - np.array(data["mean"])
        ref = (-0.5 * r @ np.array(data["covinv"]) @ r - 0.5 * (2 * np.log(2 * np.pi) + np.linalg.slogdet(cov)[1])
               - np.sum(np.log(spec.box[1] - spec.box[0])))
        assert abs(lp[i] - ref) < 2e-5 * max(1.0, abs(ref))      # Romberg EPS 1e-6 on each distance


def test_cmbdp_de_conservative_is_an_error(oracle):
    """wmap.c:1041-1044: a model outside the de_conservative range raises wmap_de_prior (point dropped),
    unlike SN / BAO which return log L = 0 (sn.c:263-274, bao.c:154-176)."""
    names, lo, hi = ["Omega_b", "Omega_m", "Omega_de", "h_100", "w_0_de"], [0.02, 0.1, 0.3, 0.5, -1.5], [0.08, 0.6, 1.1, 0.9, 0.0]
    x = np.array([[0.045, 0.27, 0.73, 0.71, -0.8], [0.045, 0.27, 0.73, 0.71, -1.2], [0.045, 0.27, 0.73, 0.71, -0.2]])
    lp, err = oracle.posterior_log_pdf(T.TargetSpec(names, lo, hi).add_cmbdp(special="de_conservative"), x)
    assert list(err != 0) == [False, True, True]
    lp, err = oracle.posterior_log_pdf(T.TargetSpec(names, lo, hi).add_bao(T.BAO_BOSS12_DZ, special="de_conservative"), x)
    assert err.sum() == 0 and lp[1] == lp[2] and lp[0] != lp[1]


def test_weights_normalisation_perplexity_ess(oracle):
    rng = np.random.default_rng(3)
    N = 5000
    logw = rng.normal(size=N) * 3 - 700.0            # would underflow without the max-shift
    flg = (rng.random(N) > 0.1).astype(np.int16)
    mx = logw[flg != 0].max()
    w, s, logsum = oracle.normalize_weights(logw, flg, mx)
    ref = np.where(flg != 0, np.exp(logw - mx), 0.0)
    assert np.allclose(w, ref / ref.sum(), rtol=1e-13) and abs(w.sum() - 1) < 1e-12
    assert abs(logsum - (np.log(ref.sum()) + mx)) < 1e-12
    perp, ess = oracle.perplexity_and_ess(w, flg)
    wp = w[w > 0]
    assert abs(perp - np.exp(-np.sum(wp * np.log(wp))) / N) < 1e-14
    assert abs(ess - 1 / np.sum(w ** 2)) < 1e-9 * ess
    # uniform weights: perplexity = (number of live points)/N, ESS = number of live points
    w, _, _ = oracle.normalize_weights(np.zeros(N), flg, 0.0)
    perp, ess = oracle.perplexity_and_ess(w, flg)
    nlive = int((flg != 0).sum())
    assert abs(perp - nlive / N) < 1e-12 and abs(ess - nlive) < 1e-6


def test_em_update_recovers_gaussian(oracle):
    """With uniform weights and one component the RB update is the sample mean / covariance;
    with a two-component proposal and a two-mode sample it separates the modes."""
    rng = np.random.default_rng(5)
    N = 40000
    true_mean = np.array([1.0, -2.0, 0.5])
    Aa = rng.standard_normal((3, 3)) * 0.3
    true_cov = Aa @ Aa.T + 0.2 * np.eye(3)
    X = rng.multivariate_normal(true_mean, true_cov, N)
    flg = np.ones(N, np.int16); idx = np.zeros(N, np.int32); wbar = np.full(N, 1.0 / N)
    ch = oracle.cholesky_stack(np.eye(3)[None] * 4)
    w, m, ch2, cov, nd = oracle.update_prop_rb(X, idx, flg, wbar, [1.0], np.zeros((1, 3)), ch)
    assert nd == 0 and abs(w[0] - 1) < 1e-15
    assert np.allclose(m[0], X.mean(0), rtol=1e-12)
    assert np.allclose(cov[0], np.cov(X.T, bias=True), rtol=1e-10)
    assert np.allclose(ch2[0] @ ch2[0].T, cov[0], rtol=1e-12)
    # flagged-out samples do not contribute
    flg2 = flg.copy(); flg2[::2] = 0
    w2, m2, _, _, _ = oracle.update_prop_rb(X, idx, flg2, np.where(flg2 != 0, 2.0 / N, 0), [1.0], np.zeros((1, 3)), ch)
    assert np.allclose(m2[0], X[1::2].mean(0), rtol=1e-12)
    # dead components: alpha < 1/N or < MINCOUNT draws
    X2 = np.concatenate([X, X[:10] + 50.0])
    idx2 = np.concatenate([idx, np.ones(10, np.int32)])
    N2 = len(X2)
    ch3 = oracle.cholesky_stack(np.stack([np.eye(3) * 4, np.eye(3)]))
    w3, m3, _, _, nd3 = oracle.update_prop_rb(X2, idx2, np.ones(N2, np.int16), np.full(N2, 1 / N2), [0.5, 0.5],
                                              np.array([[0, 0, 0], true_mean + 50.0]), ch3)
    assert nd3 == 1 and w3[1] == 0.0 and abs(w3[0] - 1) < 1e-15


@pytest.mark.parametrize("df", [-1, 4])
def test_em_update_against_independent_numpy(oracle, df):
    """The oracle's Rao-Blackwellised update (the checker of the GPU EM kernels) against a plain numpy / scipy
    restatement of Cappe et al. 2008 eqs. 12-14 (Student-t: gamma = (nu + d) / (nu + m), mu' = B / G,
    Sigma' = sum w rho gamma (x - mu')(x - mu')^T / alpha'): multi-component, non-uniform weights."""
    rng = np.random.default_rng(17 + df)
    N, K, d = 6000, 4, 3
    mean = rng.normal(size=(K, d)) * 1.5
    Aa = rng.normal(size=(K, d, d)) * 0.4
    cov = np.einsum("kij,klj->kil", Aa, Aa) + 0.5 * np.eye(d)
    alpha = np.array([0.4, 0.3, 0.2, 0.1])
    comp = rng.choice(K, size=N, p=alpha)
    X = np.stack([rng.multivariate_normal(mean[k], cov[k]) for k in comp]) + 0.3
    wbar = rng.random(N) ** 3
    flg = (rng.random(N) > 0.05).astype(np.int16)
    wbar[flg == 0] = 0.0
    wbar /= wbar.sum()
    ch = oracle.cholesky_stack(cov)
    w2, m2, ch2, cov2, nd = oracle.update_prop_rb(X, comp.astype(np.int32), flg, wbar, alpha, mean, ch, df=df)
    assert nd == 0
    if df > 0:
        pdf = np.stack([stats.multivariate_t(mean[k], cov[k], df=df).pdf(X) for k in range(K)], 1)
        icov = np.linalg.inv(cov)
        dx = X[:, None, :] - mean[None]
        maha = np.einsum("nki,kij,nkj->nk", dx, icov, dx)
        gam = (df + d) / (df + maha)
    else:
        pdf = np.stack([stats.multivariate_normal(mean[k], cov[k]).pdf(X) for k in range(K)], 1)
        gam = np.ones((N, K))
    rho = alpha * pdf
    rho /= rho.sum(1, keepdims=True)
    wr = wbar[:, None] * rho
    A_ = wr.sum(0)
    G_ = (wr * gam).sum(0)
    mu = np.einsum("nk,ni->ki", wr * gam, X) / G_[:, None]
    dxn = X[:, None, :] - mu[None]
    Sig = np.einsum("nk,nki,nkj->kij", wr * gam, dxn, dxn) / A_[:, None, None]
    assert np.allclose(w2, A_ / A_.sum(), rtol=1e-11)
    assert np.allclose(m2, mu, rtol=1e-10, atol=1e-12)
    assert np.allclose(cov2, Sig, rtol=1e-9, atol=1e-12)


def test_fisher_stencil_of_a_gaussian_posterior(oracle):
    """go_fishing.c:37-85 restated (the checker of pmcb200_fisher_host): NR (5.7.10) central second differences of
    the oracle's posterior, with the reference's diagonal shortcut, recover Sigma^-1 of a Gaussian target."""
    d = 4
    rng = np.random.default_rng(2)
    Aa = rng.normal(size=(d, d)) * 0.4
    cov = Aa @ Aa.T + np.eye(d)
    lo, hi = -6.0 * np.ones(d), 6.0 * np.ones(d)
    spec = T.TargetSpec(["dummy%d" % j for j in range(d)], lo, hi).add_mix([1.0], [0.1 * np.ones(d)], [cov])
    pos, h = 0.05 * np.arange(d), 0.01 * (hi - lo)
    diff = [(+1, +1), (+1, -1), (-1, +1), (-1, -1)]
    F = np.zeros((d, d))
    for a in range(d):
        for b in range(a, d):
            c = []
            for j in range(4):
                if j == 2 and a == b:
                    c.append(c[1]); continue
                p = pos.copy(); p[a] += diff[j][0] * h[a]; p[b] += diff[j][1] * h[b]
                lp, err = oracle.posterior_log_pdf(spec, p[None])
                assert not err.any()
                c.append(lp[0])
            F[a, b] = F[b, a] = -(c[0] - c[1] - c[2] + c[3]) / (4.0 * h[a] * h[b])
    assert np.max(np.abs(F - np.linalg.inv(cov))) < 1e-8


def test_evidence_known_answer_tempering_demo(oracle):
    """Demo/tempering/README.md:11-36: Gaussian target on the unit square => evidence ~ 1."""
    spec = T.target_gauss2d()
    K = 4
    rng = np.random.default_rng(2)
    mean = 0.5 + 0.2 * (rng.random((K, 2)) - 0.5)
    ch = oracle.cholesky_stack(np.repeat((np.eye(2) * 0.05)[None], K, 0))
    w = np.full(K, 1 / K)
    st = None
    for it in range(5):
        o = oracle.iteration(spec, 30000, 11, it, 1.0, w, mean, ch, nthreads=0)
        w, mean, ch, st = o["wght"], o["mean"], o["chol"], o["stats"]
    assert abs(np.exp(st["ln_evidence"]) - 1) < 0.02
    assert st["perplexity"] > 0.9 and st["ess"] > 0.8 * 30000 * 0.9


@pytest.mark.skipif(not (os.path.isdir(REF) and shutil.which("perl")), reason="needs the reference tree + perl")
def test_formats_against_reference_perl_restatements(oracle, tmp_path):
    """bin/evidence.pl and bin/neff_proposal.pl are in-tree restatements of the evidence / ENC
    formulas on the pmcsim / proposal file formats: run them on files written from oracle output."""
    spec = T.target_gauss2d()
    K = 3
    mean = np.array([[0.4, 0.5], [0.5, 0.6], [0.6, 0.4]])
    cov = np.repeat((np.eye(2) * 0.03)[None], K, 0)
    ch = oracle.cholesky_stack(cov)
    N = 20000
    o = oracle.iteration(spec, N, 3, 0, 1.0, np.array([0.2, 0.3, 0.5]), mean, ch)
    st = o["stats"]
    # pmcsim: log w (unnormalised), -component, params  (exec/exec_helper.c:378-424)
    ok = o["flg"] != 0
    logw = np.log(o["w"][ok]) + st["logSum"]
    with open(tmp_path / "pmcsim", "w") as f:
        f.write("# npar = 2, n_ded = 0\n#         weight            chi2\n")
        for lw, i, x in zip(logw, o["idx"][ok], o["X"][ok]):
            f.write("%16.9g%16.9g%16.9g%16.9g\n" % (lw, -float(i), x[0], x[1]))
    out = subprocess.run(["perl", os.path.join(REF, "bin/evidence.pl"), str(tmp_path / "pmcsim")],
                         capture_output=True, text=True).stdout.split("\n")[1].split()
    # evidence.pl divides by the number of lines (flagged samples); ours by all N draws
    ln_e_perl = float(out[1]) + np.log(ok.sum() / N)
    assert abs(ln_e_perl - st["ln_evidence"]) < 1e-4
    # proposal file (mix_mvdens format, Manual/manual.tex:3204-3255)
    with open(tmp_path / "proposal", "w") as f:
        f.write("%d %d\n" % (K, 2))
        covn = o["chol"] @ o["chol"].transpose(0, 2, 1)
        for k in range(K):
            f.write("%.10g\n2 -1 2 0\n" % o["wght"][k])
            f.write(" ".join("%.10g" % v for v in o["mean"][k]) + "\n")
            for r in covn[k]:
                f.write(" ".join("%.10g" % v for v in r) + "\n")
    out = subprocess.run(["perl", os.path.join(REF, "bin/neff_proposal.pl"), str(tmp_path / "proposal")],
                         capture_output=True, text=True).stdout.split("\n")[1].split()
    assert abs(float(out[1]) - st["enc"]) < 2e-3      # the script prints %.3f
