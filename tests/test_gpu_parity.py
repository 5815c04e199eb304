"""GPU parity tests: every stage of the PMC iteration through the C-ABI
(libpmc_b200.so) against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): log-likelihoods and log-weights 1e-10
relative; updated proposal (alpha, mu, Sigma), perplexity, ENC 1e-8 relative;
component selection bit-exact for identical uniforms.
"""
import numpy as np
import pytest
import torch

from cosmopmc_b200 import targets as T

pytestmark = pytest.mark.gpu

RTOL_LOG = 1e-10
RTOL_EM = 1e-8


def rel(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)) if a.size else 0.0


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def sn_setup(oracle, pmc, K=10, spec=None, df=-1):
    spec = spec or T.target_sn_demo()
    w, m, cov = T.proposal_sn(K)
    ch = oracle.cholesky_stack(cov)
    pmc.set_target(spec)
    pmc.set_proposal(w, m, chol=ch, df=df)
    return spec, w, m, ch


# ---------------------------------------------------------------- sampler ----
def test_component_selection_bit_exact(oracle, pmc_factory):
    pmc = pmc_factory()
    spec, w, m, ch = sn_setup(oracle, pmc)
    w = np.array([0.05, 0.0, 0.2, 0.15, 0.0, 0.1, 0.1, 0.25, 0.05, 0.1])
    w /= w.sum()
    pmc.set_proposal(w, m, chol=ch)
    rng = np.random.default_rng(5)
    N = 200000
    u = rng.random(N)
    cw = np.cumsum(w)
    # adversarial uniforms: exactly on, just below and just above the cumulative sums
    edge = np.concatenate([cw, np.nextafter(cw, 0), np.nextafter(cw, 2), [0.0, np.nextafter(1.0, 0)]])
    u[:edge.size] = np.clip(edge, 0.0, np.nextafter(1.0, 0))
    z = rng.standard_normal((N, 5))
    X0, idx0, flg0, _ = oracle.simulate_from_draws(u, z, w, m, ch, *spec.box)
    b = pmc.alloc(N)
    pmc.simulate_from_draws(dev(u), dev(z), b["X"], b["idx"], b["flg"])
    assert np.array_equal(b["idx"].cpu().numpy(), idx0)          # bit-exact
    assert not np.isin(idx0, [1, 4]).any()                        # dead components never drawn
    assert rel(b["X"].cpu().numpy(), X0, 1e-2) < 1e-12
    assert np.array_equal(b["flg"].cpu().numpy(), flg0)


@pytest.mark.parametrize("df", [-1, 3])
def test_philox_sampler_matches_oracle(oracle, pmc_factory, df):
    pmc = pmc_factory()
    spec, w, m, ch = sn_setup(oracle, pmc, df=df)
    N, seed, it, off = 50000, 20090903, 3, 123456789012
    X0, idx0, flg0, nok0 = oracle.simulate(N, seed, it, off, w, m, ch, *spec.box, df=df)
    b = pmc.alloc(N)
    pmc.simulate_mix_mvdens(N, seed, it, off, b["X"], b["idx"], b["flg"])
    X1 = b["X"].cpu().numpy()
    assert np.array_equal(b["idx"].cpu().numpy(), idx0)
    assert rel(X1, X0, 1e-2) < 1e-11
    assert (b["flg"].cpu().numpy() != flg0).sum() == 0
    # shard independence: the draws of a sub-range do not depend on the shard layout
    b2 = pmc.alloc(1000)
    pmc.simulate_mix_mvdens(1000, seed, it, off + 777, b2["X"], b2["idx"], b2["flg"])
    assert torch.equal(b2["X"], b["X"][777:1777])


def test_sampler_moments(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_gauss2d()
    pmc.set_target(spec)
    mean = np.array([[0.3, 0.6]])
    cov = np.array([[[0.01, 0.004], [0.004, 0.02]]])
    pmc.set_proposal([1.0], mean, cov=cov)
    N = 2000000
    b = pmc.alloc(N)
    pmc.simulate_mix_mvdens(N, 7, 0, 0, b["X"], b["idx"], b["flg"])
    X = b["X"].cpu().numpy()
    assert np.allclose(X.mean(0), mean[0], atol=5e-4)
    assert np.allclose(np.cov(X.T), cov[0], atol=2e-4)


# ------------------------------------------------------------- log-pdf -------
@pytest.mark.parametrize("K,d,df", [(10, 5, -1), (9, 5, 3), (30, 8, -1), (20, 20, -1), (3, 2, -1), (4, 32, -1)])
def test_mix_mvdens_log_pdf(oracle, pmc_factory, K, d, df):
    pmc = pmc_factory()
    rng = np.random.default_rng(K * 100 + d)
    A = rng.standard_normal((K, d, d)) * 0.3
    cov = A @ A.transpose(0, 2, 1) + np.eye(d)[None] * 0.5
    mean = rng.standard_normal((K, d))
    w = rng.random(K); w[K // 2] = 0.0; w /= w.sum()
    ch = oracle.cholesky_stack(cov)
    pmc.set_proposal(w, mean, chol=ch, df=df)
    N = 20000
    X = mean[rng.integers(0, K, N)] + rng.standard_normal((N, d)) * 1.5
    X[:10] += 1e3                      # far tail -> log 0 = -inf on both sides
    ref = oracle.mix_log_pdf(X, w, mean, ch, df)
    got = pmc.mix_mvdens_log_pdf(dev(X)).cpu().numpy()
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    assert (~fin).sum() >= (10 if df < 0 else 0)
    assert rel(got[fin], ref[fin]) < RTOL_LOG


def test_empty_batches(oracle, pmc_factory):
    pmc = pmc_factory()
    sn_setup(oracle, pmc)
    X = torch.empty((0, 5), dtype=torch.float64, device="cuda")
    assert pmc.mix_mvdens_log_pdf(X).numel() == 0
    lp, err = pmc.posterior_log_pdf(X)
    assert lp.numel() == 0 and err.numel() == 0


# --------------------------------------------------------- likelihoods -------
def box_samples(spec, N, seed, shrink=0.0):
    lo, hi = spec.box
    rng = np.random.default_rng(seed)
    return lo + (shrink + (1 - 2 * shrink) * rng.random((N, len(lo)))) * (hi - lo)


class environ:
    """set environment variables for the duration of a with block (the library reads its switches per call)"""

    def __init__(self, kv):
        self.kv = kv

    def __enter__(self):
        import os
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        import os
        for k, v in self.old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def check_posterior(oracle, pmc, spec, X):
    ref, eref = oracle.posterior_log_pdf(spec, X)
    got, egot = pmc.posterior_log_pdf(dev(X))
    got, egot = got.cpu().numpy(), egot.cpu().numpy()
    assert np.array_equal(egot != 0, eref != 0)
    ok = eref == 0
    assert ok.sum() > 0.5 * len(X)
    r = rel(got[ok], ref[ok])
    assert r < RTOL_LOG, r
    if any(spec.t.like[i].kind == 6 for i in range(spec.t.ndata)):
        # CMB distance priors: the distance to a* comes from the spectral form of the 11-stage Romberg rule where it is
        # certified (cosmo.cuh cmb_spec_w), node by node elsewhere; PMCB200_CMB_EXACT=1 takes every sample node by node
        with environ({"PMCB200_CMB_EXACT": "1"}):
            pmc.counters()
            gotx, egotx = pmc.posterior_log_pdf(dev(X))
            assert pmc.counters()["cmb_spec"] == 0
        gotx, egotx = gotx.cpu().numpy(), egotx.cpu().numpy()
        assert np.array_equal(egotx != 0, eref != 0)
        assert rel(gotx[ok], ref[ok]) < RTOL_LOG
        assert rel(got[ok], gotx[ok]) < 1e-11
    if any(spec.t.like[i].kind in (6, 7) for i in range(spec.t.ndata)):
        # BAO / CMB distance priors: the round-1 kernels (libdevice integrand, one lane per integral) are kept behind
        # PMCB200_LIKE_V1 for A/B measurements -- same integrals, same error flags
        with environ({"PMCB200_LIKE_V1": "1"}):
            got1, egot1 = pmc.posterior_log_pdf(dev(X))
        got1, egot1 = got1.cpu().numpy(), egot1.cpu().numpy()
        assert np.array_equal(egot1 != 0, eref != 0)
        assert rel(got1[ok], ref[ok]) < RTOL_LOG
    if any(spec.t.like[i].kind == 3 for i in range(spec.t.ndata)):
        # the SN likelihood has four kernels: spectral form of the quadrature on the FP64 tensor cores (large batches;
        # hands what it cannot certify to the warp kernel), the same one sample per thread (chi2_betaz, add_logdetCov),
        # one sample per warp (small batches, node by node), one sample per thread (node by node).  This batch through all.
        small = len(X) <= 4096
        for env in ({"PMCB200_SN_WARP_MAX": "0" if small else "1000000000"},
                    {"PMCB200_SN_WARP_MAX": "0", "PMCB200_SN_SPEC_V1": "1"},
                    {"PMCB200_SN_WARP_MAX": "0", "PMCB200_SN_TAIL32": "1"},      # opt-in TF32 coefficient tail
                    {"PMCB200_SN_WARP_MAX": "0", "PMCB200_SN_EXACT": "1"}):
            with environ(env):
                got2, egot2 = pmc.posterior_log_pdf(dev(X))
            got2, egot2 = got2.cpu().numpy(), egot2.cpu().numpy()
            assert np.array_equal(egot2 != 0, eref != 0), env
            assert rel(got2[ok], ref[ok]) < RTOL_LOG, env
            # same per-redshift result to ~1e-15, the sum over redshifts associated differently; log pi = -chi2/2 +
            # constants can come out near zero, which amplifies the last bits of chi2 in this relative measure
            assert rel(got2[ok], got[ok]) < 1e-11, env
    return r


def test_posterior_sn_golden_fiducial(oracle, pmc_factory):
    """log-posterior at the reference test-suite's fiducial point
    (bin/test_suite_cosmo_pmc.pl:51) and at the manual's posterior mean,
    against the committed golden values generated from the oracle."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sn_logpost.json")))
    pmc = pmc_factory()
    spec = T.target_sn_demo()
    pmc.set_target(spec)
    X = np.array(g["x"])
    got, err = pmc.posterior_log_pdf(dev(X))
    assert int(err.sum()) == 0
    assert rel(got.cpu().numpy(), np.array(g["logpost"])) < RTOL_LOG


def test_posterior_sn_box(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_sn_demo()
    pmc.set_target(spec)
    check_posterior(oracle, pmc, spec, box_samples(spec, 3000, 11))


def test_posterior_sn_curved(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_sn_curved()
    pmc.set_target(spec)
    X = box_samples(spec, 3000, 12)
    X[:50, 1] = 1.0 - X[:50, 0]          # exactly flat rows exercise the |Omega_K| < eps branch
    check_posterior(oracle, pmc, spec, X)


@pytest.mark.parametrize("mode,logdet", [("chi2_no_sc", 0), ("chi2_betaz", 1), ("chi2_Theta2_denom_fixed", 1)])
def test_posterior_sn_modes(oracle, pmc_factory, mode, logdet):
    pmc = pmc_factory()
    spec = T.TargetSpec(["Omega_m", "w_0_de", "M", "alpha", "beta", "beta_z"],
                        [0.1, -2.0, 19.1, 0.5, -3.5, -1.0], [0.9, -0.3, 19.8, 2.6, -0.8, 1.0])
    spec.add_snia(chi2mode=mode, add_logdetCov=logdet, Theta2_denom=(0.0, 1.5, -2.0))
    pmc.set_target(spec)
    check_posterior(oracle, pmc, spec, box_samples(spec, 1000, 13))


def test_posterior_sn_bao_w0wa(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_sn_bao_w0wa()
    pmc.set_target(spec)
    check_posterior(oracle, pmc, spec, box_samples(spec, 1500, 14))


def test_posterior_de_conservative(oracle, pmc_factory):
    """special prior de_conservative on every probe: volume term (param.c:1072-1094); the SN and BAO probes
    return log L = 0 for a violating model (sn.c:263-274, bao.c:154-176), likeli_CMBDistPrior raises
    wmap_de_prior (wmap.c:1041-1044) so the point is dropped with zero weight."""
    pmc = pmc_factory()
    names = ["Omega_b", "Omega_m", "Omega_de", "h_100", "w_0_de", "w_1_de", "M", "alpha", "beta"]
    lo, hi = [0.02, 0.1, 0.3, 0.5, -1.5, -1.5, 19.1, 0.5, -3.5], [0.08, 0.6, 1.1, 0.9, 0.0, 1.5, 19.8, 2.6, -0.8]
    spec = (T.TargetSpec(names, lo, hi).add_bao(T.BAO_BOSS12_DZ, special="de_conservative")
            .add_snia(cosmo=T.COSMO_DP, special="de_conservative"))
    pmc.set_target(spec)
    X = box_samples(spec, 1500, 21)
    w_now, w_acc = X[:, 4], X[:, 4] + X[:, 5] / 3.0
    cutm = (w_now < -1) | (w_now > -1 / 3) | (w_acc < -1) | (w_acc > -1 / 3)
    assert 0.2 < cutm.mean() < 0.95
    check_posterior(oracle, pmc, spec, X)
    got, err = pmc.posterior_log_pdf(dev(X))
    got, err = got.cpu().numpy(), err.cpu().numpy()
    assert np.ptp(got[cutm & (err == 0)]) < 1e-12   # both probes returned 0: only the constant prior is left
    assert (cutm & (err == 0)).sum() > 0.9 * cutm.sum()
    # with the CMB distance priors in the set a violating model is an error (flag 0), never log L = 0
    spec3 = (T.TargetSpec(names, lo, hi).add_cmbdp(special="de_conservative")
             .add_bao(T.BAO_BOSS12_DZ, special="de_conservative").add_snia(cosmo=T.COSMO_DP, special="de_conservative"))
    pmc.set_target(spec3)
    ref, eref = oracle.posterior_log_pdf(spec3, X)
    got, err = pmc.posterior_log_pdf(dev(X))
    got, err = got.cpu().numpy(), err.cpu().numpy()
    assert np.array_equal(err != 0, eref != 0) and np.all(err[cutm] != 0) and (err[~cutm] == 0).sum() > 50
    assert rel(got[eref == 0], ref[eref == 0]) < RTOL_LOG
    # narrower w0 range than the prior: refused like the reference does
    bad = T.TargetSpec(["Omega_m", "w_0_de", "M", "alpha", "beta"], [0.0, -0.8, 19.1, 0.5, -3.5],
                       [1.2, 0.5, 19.8, 2.6, -0.8]).add_snia(special="de_conservative")
    with pytest.raises(Exception):
        pmc.set_target(bad)


def test_posterior_cmb_bao_sn(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_cmb_bao_sn()
    pmc.set_target(spec)
    check_posterior(oracle, pmc, spec, box_samples(spec, 1000, 15))


def test_posterior_bao_A_and_prior(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.TargetSpec(["Omega_m", "Omega_de", "h_100"], [0.05, 0.2, 0.5], [0.8, 1.2, 0.9])
    spec.add_bao(T.BAO_REID10_A)
    spec.set_prior([0.3, 0.72], [[0.01, 0.0002], [0.0002, 0.0064]], indprior=[1, 0, 1])   # (Omega_m, h_100)
    pmc.set_target(spec)
    check_posterior(oracle, pmc, spec, box_samples(spec, 1500, 16))


def test_posterior_bao_D_V_ratio(oracle, pmc_factory):
    """distance_D_V_ratio (bao.c:56,170-171): model_i = D_V(z_2i) / D_V(z_2i+1) against a 2-point Gaussian
    (Percival et al. 2010-like ratios D_V(0.35)/D_V(0.2) and D_V(0.57)/D_V(0.35)), curved w0 cosmology."""
    pmc = pmc_factory()
    spec = T.TargetSpec(["Omega_m", "Omega_de", "w_0_de", "h_100"], [0.05, 0.2, -2.0, 0.5], [0.8, 1.2, -0.4, 0.9])
    spec.add_bao(dict(method="distance_D_V_ratio", mean=[1.736, 1.52], covinv=[[260.0, -40.0], [-40.0, 400.0]],
                      z=[0.35, 0.2, 0.57, 0.35]))
    pmc.set_target(spec)
    r = check_posterior(oracle, pmc, spec, box_samples(spec, 2000, 22))
    # the ratio is scale free: h_100 must not matter (property of the reference formula)
    X = box_samples(spec, 200, 23)
    X2 = X.copy(); X2[:, 3] = 0.55
    a, _ = pmc.posterior_log_pdf(dev(X)); b, _ = pmc.posterior_log_pdf(dev(X2))
    assert rel(a.cpu().numpy(), b.cpu().numpy()) < 1e-12


def test_posterior_bao_deep_romberg_stages(oracle, pmc_factory):
    """w0 + w1 > 1/3: dark energy dominates the early universe, the sound-horizon integrand is singular at a -> 0
    and the Romberg rule runs deep (up to 2^19 nodes, or 'too many steps').  On the device those stages are
    evaluated by the whole warp for one lane at a time (romberg_warp); values and error flags must still be the
    oracle's, for the stragglers and for their well-behaved warp neighbours."""
    pmc = pmc_factory()
    spec = T.target_sn_bao_w0wa()
    pmc.set_target(spec)
    rng = np.random.default_rng(31)
    N = 96
    X = np.array([0.28, 0.72, -1.0, 0.0, 19.31, 1.4, -2.4])[None, :] + rng.normal(size=(N, 7)) * np.array([0.04, 0.06, 0.15, 0.2, 0.03, 0.1, 0.1])
    hard = [3, 17, 40, 41, 77]                      # a few lanes of three different warps
    X[hard, 2] = [-0.4, -0.2, -0.6, -0.3, -0.1]
    X[hard, 3] = [0.9, 0.7, 1.2, 1.5, 0.5]
    ref, eref = oracle.posterior_log_pdf(spec, X)
    got, egot = pmc.posterior_log_pdf(dev(X))
    got, egot = got.cpu().numpy(), egot.cpu().numpy()
    assert np.array_equal(egot != 0, eref != 0)
    ok = eref == 0
    assert ok.sum() >= N - len(hard)
    assert rel(got[ok], ref[ok]) < RTOL_LOG
    c = pmc.counters()
    assert c["gen_evals"] > 2 * N * 17              # the deep stages were walked (and counted)


def test_posterior_banana_and_mixture(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_banana(20)
    pmc.set_target(spec)
    check_posterior(oracle, pmc, spec, box_samples(spec, 5000, 17))
    spec = T.TargetSpec(["dummy0", "dummy1"], [0, 0], [1, 1]).add_mix(
        [0.5, 0.5], [[0.3, 0.3], [0.7, 0.7]], [np.eye(2) * 0.01, [[0.02, 0.005], [0.005, 0.01]]],
        special="unity")
    pmc.set_target(spec)
    check_posterior(oracle, pmc, spec, box_samples(spec, 5000, 18))


def test_likelihood_error_policy(oracle, pmc_factory):
    """Errors inside a likelihood give the sample zero weight (manual.tex:507-512):
    unphysical cosmologies (negative a^4 E^2) must be flagged on both sides."""
    pmc = pmc_factory()
    spec = T.TargetSpec(["Omega_m", "Omega_de", "M", "alpha", "beta"],
                        [0.0, 0.0, 19.1, 0.5, -3.5], [1.2, 4.0, 19.8, 2.6, -0.8]).add_snia()
    pmc.set_target(spec)
    X = box_samples(spec, 2000, 19)
    X[:, 1] = 1.5 + 2.5 * np.random.default_rng(3).random(2000)    # bouncing / no-big-bang region
    ref, eref = oracle.posterior_log_pdf(spec, X)
    got, egot = pmc.posterior_log_pdf(dev(X))
    egot = egot.cpu().numpy()
    assert (eref != 0).sum() > 100 and (eref == 0).sum() > 100
    assert np.array_equal(egot != 0, eref != 0)
    ok = eref == 0
    assert rel(got.cpu().numpy()[ok], ref[ok]) < RTOL_LOG


# ------------------------------------------------- weights + EM update -------
def run_both(oracle, pmc, spec, w, m, ch, N, seed=1, it=0, beta=1.0, df=-1):
    o = oracle.iteration(spec, N, seed, it, beta, w, m, ch, df=df, nthreads=8)
    hX = np.empty((N, len(m[0]))); hidx = np.empty(N, np.int32)
    hflg = np.empty(N, np.int16); hw = np.empty(N)
    st = pmc.iteration_host(N, seed, it, beta, hX, hidx, hflg, hw)
    return o, st, hX, hidx, hflg, hw


def check_iteration(o, st, pmc, hX, hidx, hflg, hw, rtol_w=1e-9):
    so = o["stats"]
    assert st["nok_box"] == so["nok_box"] and st["nok"] == so["nok"]
    assert np.array_equal(hidx, o["idx"]) and np.array_equal(hflg, o["flg"])
    assert rel(hX, o["X"], 1e-2) < 1e-11
    assert abs(st["maxW"] - so["maxW"]) < RTOL_LOG * abs(so["maxW"])
    assert abs(st["logSum"] - so["logSum"]) < RTOL_LOG * abs(so["logSum"])
    for k in ("perplexity", "ess", "ln_evidence", "enc", "sum_shift"):
        assert abs(st[k] - so[k]) <= RTOL_EM * abs(so[k]), (k, st[k], so[k])
    assert st["ndead"] == so["ndead"]
    ok = o["flg"] != 0
    big = ok & (o["w"] > 1e-12 * o["w"].max())
    assert rel(hw[big], o["w"][big]) < rtol_w
    assert np.all(hw[~ok] == 0.0)
    wg, mg, chg, covg = pmc.get_proposal()
    assert rel(wg, o["wght"], 1e-300) < RTOL_EM
    live = o["wght"] > 0
    assert np.array_equal(wg > 0, live)
    assert rel(mg[live], o["mean"][live], 1e-3) < RTOL_EM
    covo = o["chol"] @ o["chol"].transpose(0, 2, 1)
    scale = np.sqrt(np.einsum("kii,kjj->kij", covo, covo))
    assert np.max(np.abs(covg[live] - covo[live]) / scale[live]) < RTOL_EM


def test_iteration_sn_demo(oracle, pmc_factory):
    """C1: the SN demo shape (d=5, K=10, N=10^4), full iteration vs oracle."""
    pmc = pmc_factory()
    spec, w, m, ch = sn_setup(oracle, pmc)
    o, st, *h = run_both(oracle, pmc, spec, w, m, ch, 10000, seed=20090903)
    check_iteration(o, st, pmc, *h)
    assert 0.0 < st["perplexity"] <= 1.0 and 1.0 <= st["ess"] <= 10000


def test_iteration_dead_components_and_tempering(oracle, pmc_factory):
    pmc = pmc_factory()
    spec, w, m, ch = sn_setup(oracle, pmc)
    m = m.copy(); m[3] += [0.5, 2.0, 0.4, 1.0, 1.0]     # a component far off: dies (alpha < 1/N)
    w = w.copy(); w[7] = 1e-4; w /= w.sum()               # a component with < MINCOUNT draws
    pmc.set_proposal(w, m, chol=ch)
    o, st, *h = run_both(oracle, pmc, spec, w, m, ch, 20000, seed=5, it=2, beta=0.7)
    assert o["stats"]["ndead"] >= 2
    check_iteration(o, st, pmc, *h)


def test_iteration_student_t(oracle, pmc_factory):
    pmc = pmc_factory()
    spec, w, m, ch = sn_setup(oracle, pmc, K=6, df=3)
    o, st, *h = run_both(oracle, pmc, spec, w, m, ch, 8000, seed=9, df=3)
    check_iteration(o, st, pmc, *h)


def test_iteration_banana(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_banana(20)
    w, m, cov = T.proposal_banana(10, 20)
    ch = oracle.cholesky_stack(cov)
    pmc.set_target(spec)
    pmc.set_proposal(w, m, chol=ch)
    o, st, *h = run_both(oracle, pmc, spec, w, m, ch, 30000, seed=3)
    check_iteration(o, st, pmc, *h)


@pytest.mark.parametrize("K,d,df,N", [(10, 20, -1, 30011), (20, 12, 5, 20000), (32, 11, -1, 25000),
                                       (3, 18, 4, 9000), (8, 10, -1, 4096), (4, 32, 7, 12000),
                                       (10, 5, -1, 20011), (30, 8, -1, 15000), (6, 5, 3, 9000), (9, 2, -1, 7000),
                                       (5, 3, 6, 5000), (12, 7, 4, 8000), (17, 1, -1, 6000), (16, 32, 5, 6000),
                                       (8, 9, -1, 7000), (24, 6, -1, 9000), (11, 4, 3, 5000)])
def test_iteration_em_tensor_core_shapes(oracle, pmc_factory, K, d, df, N, monkeypatch):
    """K <= 32 runs the EM statistics on the FP64 tensor cores (d <= 10: k_em_stats_mma_ws, samples sliced over the warps;
    else k_em_stats_mma, feature tiles over the warps): every component-tile count MT = 1..4, every padded dimension class
    (odd, even, 9 -> 10, 11 -> 12, 18 -> 20), Gaussian and Student-t, with and without the E-step cache, ragged last tile
    (and K = 16, d = 32 Student-t, whose staging exceeds the shared-memory budget: shared-memory kernel); against the
    oracle, against the shared-memory kernel (PMCB200_EM_NO_MMA=1) and, for d <= 10, against the tile-sliced tensor-core
    kernel (PMCB200_EM_NO_WS=1) on the same sample."""
    pmc = pmc_factory()
    rng = np.random.default_rng(100 * K + d)
    lo, hi = -6.0 * np.ones(d), 6.0 * np.ones(d)
    tc = np.diag(0.5 + rng.random(d))
    spec = T.TargetSpec(["dummy%d" % j for j in range(d)], lo, hi).add_mix([0.6, 0.4], [np.zeros(d), 0.5 * np.ones(d)],
                                                                         [tc, 0.7 * tc])
    mean = rng.normal(size=(K, d)) * 0.4
    A_ = rng.normal(size=(K, d, d)) * 0.15
    cov = np.einsum("kij,klj->kil", A_, A_) + np.eye(d)[None] * (1.0 + rng.random((K, 1, 1)))
    w = rng.random(K) + 0.2; w /= w.sum()
    ch = oracle.cholesky_stack(cov)
    pmc.set_target(spec)
    pmc.set_proposal(w, mean, chol=ch, df=df)
    o, st, *h = run_both(oracle, pmc, spec, w, mean, ch, N, seed=11, df=df)
    check_iteration(o, st, pmc, *h)
    wg, mg, chg, covg = pmc.get_proposal()
    monkeypatch.setenv("PMCB200_EM_NO_MMA", "1")
    pmc.set_proposal(w, mean, chol=ch, df=df)
    o2, st2, *h2 = run_both(oracle, pmc, spec, w, mean, ch, N, seed=11, df=df)
    w2, m2, ch2, cov2 = pmc.get_proposal()
    assert rel(wg, w2) < 1e-11 and rel(mg, m2, 1e-3) < 1e-10
    assert np.max(np.abs(covg - cov2)) < 1e-10 * np.max(np.abs(cov2))
    assert abs(st["perplexity"] - st2["perplexity"]) < 1e-12 * st2["perplexity"]
    if d <= 10:
        monkeypatch.delenv("PMCB200_EM_NO_MMA")
        monkeypatch.setenv("PMCB200_EM_NO_WS", "1")
        pmc.set_proposal(w, mean, chol=ch, df=df)
        o3, st3, *h3 = run_both(oracle, pmc, spec, w, mean, ch, N, seed=11, df=df)
        w3, m3, ch3, cov3 = pmc.get_proposal()
        assert rel(wg, w3) < 1e-11 and rel(mg, m3, 1e-3) < 1e-10
        assert np.max(np.abs(covg - cov3)) < 1e-10 * np.max(np.abs(cov3))


@pytest.mark.parametrize("cfg", ["sn", "banana", "student"])
def test_build_variants_behind_switches_agree(oracle, pmc_factory, cfg):
    """Every kernel variant the library keeps behind a run-time switch (A/B measurements) must deliver what the default
    path delivers: weight-kernel variants PMCB200_ESTEP = 0 (one sample per thread), 1, 3 (samples per thread), 4
    (column-packed staging); sampler PMCB200_SIM_STAGED = 0 / 1; no E-step cache (PMCB200_EM_NO_RHO); tile-sliced
    tensor-core EM kernel (PMCB200_EM_NO_WS).  Same seed => same samples, flags, indices bit for bit; weights and the updated
    proposal to 1e-12 (the variants re-associate nothing in a sample's own arithmetic; the EM variants sum in another order)."""
    pmc = pmc_factory()
    if cfg == "sn":
        spec = T.target_sn_demo(); w, m, cov = T.proposal_sn(10); df = -1
    elif cfg == "banana":
        spec = T.target_banana(20); w, m, cov = T.proposal_banana(10, 20); df = -1
    else:
        spec = T.target_banana(8); w, m, cov = T.proposal_banana(12, 8); df = 4
    ch = oracle.cholesky_stack(cov)
    N, d = 24001, len(m[0])
    pmc.set_target(spec)

    def run(env):
        with environ(env):
            pmc.set_proposal(w, m, chol=ch, df=df)
            hX = np.empty((N, d)); hidx = np.empty(N, np.int32); hflg = np.empty(N, np.int16); hw = np.empty(N)
            st = pmc.iteration_host(N, 5, 1, 1.0, hX, hidx, hflg, hw)
            return st, hX, hidx, hflg, hw, pmc.get_proposal()
    st0, X0, i0, f0, w0, p0 = run({})
    for env in ({"PMCB200_ESTEP": "0"}, {"PMCB200_ESTEP": "1"}, {"PMCB200_ESTEP": "3"}, {"PMCB200_ESTEP": "4"},
                {"PMCB200_SIM_STAGED": "0"}, {"PMCB200_SIM_STAGED": "1"}, {"PMCB200_EM_NO_RHO": "1"},
                {"PMCB200_EM_NO_WS": "1"}, {"PMCB200_EM_NO_MMA": "1"}):
        st, X, i_, f_, w_, p = run(env)
        assert np.array_equal(i_, i0) and np.array_equal(f_, f0), env
        assert rel(X, X0, 1e-2) < 1e-13, env
        assert st["nok"] == st0["nok"] and abs(st["perplexity"] - st0["perplexity"]) <= 1e-11 * st0["perplexity"], env
        big = w0 > 1e-12 * w0.max()
        assert rel(w_[big], w0[big]) < 1e-11, env
        assert rel(p[0], p0[0], 1e-300) < 1e-10 and rel(p[1], p0[1], 1e-3) < 1e-10, env
        assert np.max(np.abs(p[3] - p0[3])) < 1e-10 * np.max(np.abs(p0[3])), env


def test_iteration_cmb_bao_sn(oracle, pmc_factory):
    pmc = pmc_factory()
    spec = T.target_cmb_bao_sn()
    center = [0.045, 0.27, 0.73, 0.71, -1.0, 19.31, 1.4, -2.4]
    sigma = [0.003, 0.02, 0.02, 0.02, 0.1, 0.03, 0.1, 0.1]
    w, m, cov = T.proposal_generic(spec, 30, 4, center, sigma)
    ch = oracle.cholesky_stack(cov)
    pmc.set_target(spec)
    pmc.set_proposal(w, m, chol=ch)
    pmc.counters()
    o, st, *h = run_both(oracle, pmc, spec, w, m, ch, 6000, seed=4)
    check_iteration(o, st, pmc, *h)
    cnt = pmc.counters()
    assert cnt["cmb_spec"] > 0.9 * st["nok_box"], cnt      # on this proposal nearly every sample is certified for the spectral form


def test_iteration_sn_bao_w0wa(oracle, pmc_factory):
    """C4 (BASELINE.json configs[3]): SN Ia + BAO d_z, w0-wa, d = 7, K = 10 -- the bench's proposal, full iteration."""
    pmc = pmc_factory()
    spec = T.target_sn_bao_w0wa()
    w, m, cov = T.proposal_generic(spec, 10, 4, [0.28, 0.72, -1.0, 0.0, 19.31, 1.4, -2.4],
                                   [0.04, 0.06, 0.15, 0.2, 0.03, 0.1, 0.1])
    ch = oracle.cholesky_stack(cov)
    pmc.set_target(spec)
    pmc.set_proposal(w, m, chol=ch)
    o, st, *h = run_both(oracle, pmc, spec, w, m, ch, 8000, seed=6)
    check_iteration(o, st, pmc, *h)
    assert st["nok"] > 7000


def test_pipelined_host_delivery_matches_blocking_call(oracle, pmc_factory):
    """pmcb200_iteration_host_begin / pmcb200_host_wait: three chained iterations with the copies of one draining
    under the next give the same host arrays, statistics and proposals as three blocking pmcb200_iteration_host calls."""
    spec = T.target_sn_demo()
    w, m, cov = T.proposal_sn(10)
    ch = oracle.cholesky_stack(cov)
    N = 30000
    def host():
        return (torch.empty((N, 5), dtype=torch.float64).pin_memory(), torch.empty(N, dtype=torch.int32).pin_memory(),
                torch.empty(N, dtype=torch.int16).pin_memory(), torch.empty(N, dtype=torch.float64).pin_memory())
    a = pmc_factory(); a.set_target(spec); a.set_proposal(w, m, chol=ch)
    ref = []
    for it in range(3):
        h = host()
        st = a.iteration_host(N, 77, it, 1.0, *h)
        ref.append((st, [t.clone() for t in h], a.get_proposal()))
    b = pmc_factory(); b.set_target(spec); b.set_proposal(w, m, chol=ch)
    hs, sts = [host() for _ in range(3)], []
    for it in range(3):
        sts.append(b.iteration_host_begin(N, 77, it, 1.0, *hs[it]))
        b.host_wait(1)                               # everything but the iteration just begun is on the host
        if it > 0:
            for t, r in zip(hs[it - 1], ref[it - 1][1]):
                assert torch.equal(t, r)
    b.host_wait(0)
    for it in range(3):
        assert sts[it] == ref[it][0]
        for t, r in zip(hs[it], ref[it][1]):
            assert torch.equal(t, r)
    for x, y in zip(b.get_proposal(), ref[2][2]):
        assert np.array_equal(x, y)


def test_lazy_sample_delivery(oracle, pmc_factory):
    """hX = NULL in the iteration call leaves the sample array on the device; pmcb200_samples_host_begin fetches it on
    request: same statistics, same weights / flags / indices, and the fetched X equals the delivered one."""
    spec = T.target_sn_demo()
    w, m, cov = T.proposal_sn(10)
    ch = oracle.cholesky_stack(cov)
    N = 30000
    def host():
        return (torch.empty((N, 5), dtype=torch.float64).pin_memory(), torch.empty(N, dtype=torch.int32).pin_memory(),
                torch.empty(N, dtype=torch.int16).pin_memory(), torch.empty(N, dtype=torch.float64).pin_memory())
    a = pmc_factory(); a.set_target(spec); a.set_proposal(w, m, chol=ch)
    b = pmc_factory(); b.set_target(spec); b.set_proposal(w, m, chol=ch)
    for it in range(2):
        ha, hb = host(), host()
        sa = a.iteration_host(N, 5, it, 1.0, *ha)
        sb = b.iteration_host_begin(N, 5, it, 1.0, None, None, hb[2], hb[3])
        b.samples_host_begin(N, hb[0], hb[1])
        b.host_wait(0)
        assert sa == sb
        for x, y in zip(ha, hb):
            assert torch.equal(x, y)


def test_multi_iteration_convergence_gauss2d(pmc_factory):
    """Statistical known answer of the reference: Demo/tempering/README.md:11-36,
    2-D Gaussian on the unit square => evidence consistent with 1 and the
    perplexity climbs towards 1."""
    pmc = pmc_factory()
    spec = T.target_gauss2d()
    pmc.set_target(spec)
    rng = np.random.default_rng(1)
    K = 5
    mean = 0.5 + 0.2 * (rng.random((K, 2)) - 0.5)
    cov = np.repeat((np.eye(2) * 0.05)[None], K, 0)
    pmc.set_proposal(np.full(K, 1 / K), mean, cov=cov)
    st = None
    for it in range(6):
        st = pmc.iteration_host(100000, 42, it)
    assert abs(np.exp(st["ln_evidence"]) - 1.0) < 0.01
    assert st["perplexity"] > 0.9


def test_sharded_blocks_combine_like_one_rank(oracle, pmc_factory):
    """(e) multi-GPU logic on one device: 4 shards' stat blocks combined by
    em_finish must give the same update as one rank over all samples."""
    spec = T.target_sn_demo()
    w, m, cov = T.proposal_sn(10)
    ch = oracle.cholesky_stack(cov)
    N, world = 40000, 4
    one = pmc_factory(); one.set_target(spec); one.set_proposal(w, m, chol=ch)
    blk = torch.empty(one.stat_block_len(), dtype=torch.float64, device="cuda")
    one.iteration_local(N, 11, 0, 0, 1.0, blk)
    s1 = one.update_prop_rb(1, blk, N)
    many = pmc_factory(); many.set_target(spec); many.set_proposal(w, m, chol=ch)
    allb = torch.empty((world, many.stat_block_len()), dtype=torch.float64, device="cuda")
    per = N // world
    for r in range(world):
        many.iteration_local(per, 11, 0, r * per, 1.0, allb[r])
    s2 = many.update_prop_rb(world, allb, N)
    for k in ("nok", "nok_box", "ndead"):
        assert s1[k] == s2[k]
    for k in ("maxW", "logSum", "perplexity", "ess", "enc"):
        assert abs(s1[k] - s2[k]) <= 1e-12 * abs(s1[k])
    p1, p2 = one.get_proposal(), many.get_proposal()
    for a, b in zip(p1, p2):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-14)


def test_iteration_host_multi_contexts(oracle, pmc_factory):
    """(e) several contexts in one process (pmcb200_iteration_host_multi): 3 shards on one
    device (uneven: 10001 = 3334 + 3334 + 3333) against the single-context call and the oracle.
    Samples, indices and flags are bit-identical (Philox counter = global index); the update
    agrees to the parity tolerance; every context ends with the identical proposal."""
    from cosmopmc_b200.pmc import iteration_host_multi
    spec = T.target_sn_demo()
    w, m, cov = T.proposal_sn(10)
    ch = oracle.cholesky_stack(cov)
    N, seed = 10001, 77
    one = pmc_factory(); one.set_target(spec); one.set_proposal(w, m, chol=ch)
    X1 = np.empty((N, 5)); i1 = np.empty(N, np.int32); f1 = np.empty(N, np.int16); w1 = np.empty(N)
    s1 = one.iteration_host(N, seed, 2, 0.8, X1, i1, f1, w1)
    many = [pmc_factory() for _ in range(3)]
    for p in many:
        p.set_target(spec); p.set_proposal(w, m, chol=ch)
    X2 = np.empty((N, 5)); i2 = np.empty(N, np.int32); f2 = np.empty(N, np.int16); w2 = np.empty(N)
    s2 = iteration_host_multi(many, N, seed, 2, 0.8, X2, i2, f2, w2)
    assert np.array_equal(X1, X2) and np.array_equal(i1, i2) and np.array_equal(f1, f2)
    # the SN kernel is chosen by shard size (10001 samples: spectral tensor-core kernel, 3334: one sample per
    # warp, node by node); the two agree to ~2e-14 of log L ~ -180, i.e. ~4e-12 of a weight: parity tolerance
    assert np.allclose(w1, w2, rtol=1e-10, atol=0)
    for k in ("nok", "nok_box", "ndead", "nsamples"):
        assert s1[k] == s2[k]
    for k in ("maxW", "logSum", "perplexity", "ess", "enc", "ln_evidence"):
        assert abs(s1[k] - s2[k]) <= 1e-10 * abs(s1[k]), k
    ref = one.get_proposal()
    props = [p.get_proposal() for p in many]
    for a, b in zip(ref, props[0]):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-14)
    for q in props[1:]:
        for a, b in zip(props[0], q):
            assert np.array_equal(a, b)          # fixed-order combine: bitwise identical on every shard
    o = oracle.iteration(spec, N, seed, 2, 0.8, w, m, ch, nthreads=0)
    assert np.array_equal(i2, o["idx"]) and np.array_equal(f2, o["flg"])
    assert np.allclose(props[0][0], o["wght"], rtol=1e-8, atol=0)
    assert np.allclose(props[0][1], o["mean"], rtol=1e-8, atol=0)
    # more shards than samples: empty shards are legal
    # (N = 2: fewer than MINCOUNT draws per component, so every component dies -- on every shard alike)
    from cosmopmc_b200.pmc import PMCError
    from cosmopmc_b200 import _abi as A
    with pytest.raises(PMCError) as e:
        iteration_host_multi(many, 2, seed, 3, 1.0)
    assert e.value.code == A.ERR["NOSAMPLE"]


def test_iteration_host_multi_on_two_devices(oracle):
    """the same on two physical GPUs (skipped on a one-GPU box): contexts on cuda:0 and cuda:1,
    statistics blocks cross NVLink by cudaMemcpyPeerAsync; result identical to one context"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from cosmopmc_b200.pmc import PMC, iteration_host_multi
    spec = T.target_sn_demo()
    w, m, cov = T.proposal_sn(10)
    ch = oracle.cholesky_stack(cov)
    N, seed = 200001, 5
    res = []
    for devs in ((0,), (0, 1)):
        pmcs = [PMC(g, use_torch_stream=False) for g in devs]
        for p in pmcs:
            p.set_target(spec); p.set_proposal(w, m, chol=ch)
        hX = torch.empty((N, 5), dtype=torch.float64).pin_memory(); hi = torch.empty(N, dtype=torch.int32).pin_memory()
        hf = torch.empty(N, dtype=torch.int16).pin_memory(); hw = torch.empty(N, dtype=torch.float64).pin_memory()
        st = iteration_host_multi(pmcs, N, seed, 0, 1.0, hX, hi, hf, hw)
        res.append((st, hX.numpy().copy(), hi.numpy().copy(), hf.numpy().copy(), hw.numpy().copy(),
                    [p.get_proposal() for p in pmcs]))
        for p in pmcs:
            p.close()
    (s1, X1, i1, f1, w1, p1), (s2, X2, i2, f2, w2, p2) = res
    assert np.array_equal(X1, X2) and np.array_equal(i1, i2) and np.array_equal(f1, f2)
    assert np.allclose(w1, w2, rtol=1e-12, atol=0) and s1["nok"] == s2["nok"]
    assert abs(s1["perplexity"] - s2["perplexity"]) <= 1e-12 * s1["perplexity"]
    for a, b in zip(p1[0], p2[0]):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-14)
    for a, b in zip(p2[0], p2[1]):
        assert np.array_equal(a, b)


def test_sn_spectral_large_batch(oracle, pmc_factory):
    """The spectral SN kernel on a batch drawn from the benchmark proposal (the layout large batches take by default):
    log-posteriors against the oracle's node-by-node Romberg, against the exact device kernel, and the split between
    the spectral kernel and the samples it handed to the exact one."""
    for spec, lo_frac in ((T.target_sn_demo(), 0.99), (T.target_sn_curved(), 0.5)):
        pmc = pmc_factory()
        pmc.set_target(spec)
        w, m, cov = T.proposal_sn(10)
        rng = np.random.default_rng(77)
        N = 60000
        k = rng.integers(0, 10, N)
        X = m[k] + np.einsum("nij,nj->ni", np.linalg.cholesky(cov)[k], rng.normal(size=(N, 5)))
        if spec.spar[1] == "Omega_de":
            X[:, 1] = 0.7 + 0.35 * rng.normal(size=N)      # curved: Omega_de instead of w0
        lo, hi = spec.box
        X = X[((X >= lo) & (X <= hi)).all(axis=1)]
        pmc.counters()
        got, egot = pmc.posterior_log_pdf(dev(X))
        cnt = pmc.counters()
        assert cnt["sn_spec"] + cnt["sn_exact"] == len(X)
        assert cnt["sn_spec"] >= lo_frac * len(X), cnt
        with environ({"PMCB200_SN_EXACT": "1"}):
            ex, eex = pmc.posterior_log_pdf(dev(X))
            cnt2 = pmc.counters()
        assert cnt2["sn_spec"] == 0
        got, egot, ex, eex = got.cpu().numpy(), egot.cpu().numpy(), ex.cpu().numpy(), eex.cpu().numpy()
        assert np.array_equal(egot != 0, eex != 0)
        ok = eex == 0
        assert rel(got[ok], ex[ok]) < 1e-11
        sub = np.arange(0, len(X), 3)
        ref, eref = oracle.posterior_log_pdf(spec, X[sub])
        assert np.array_equal(egot[sub] != 0, eref != 0)
        oks = eref == 0
        assert rel(got[sub][oks], ref[oks]) < RTOL_LOG


@pytest.mark.parametrize("zs", [[0.5], [0.3] * 9, [0.02, 0.02, 0.4, 0.4, 0.4, 0.9, 1.4],
                                [0.05 + 0.06 * i for i in range(17)] + [0.11, 0.11, 0.65, 0.65, 0.65, 0.65]])
def test_sn_tile_layouts_synthetic_tables(oracle, pmc_factory, tmp_path, zs):
    """Primary / secondary supernova tiles of the tensor-core SN kernel on synthetic tables: one supernova; nine at one
    redshift (one primary tile, eight secondary ones with a single live column); multiplicities 2, 3, 1, 1; 19 redshifts
    (three primary tiles, the last one ragged) with multiplicities up to 5.  Spectral tensor-core kernel (forced for any
    batch size) against the oracle and the node-by-node device kernel."""
    rng = np.random.default_rng(len(zs))
    path = tmp_path / "sn.txt"
    with open(path, "w") as f:
        f.write("@sig_int 0.15\n@v_pec 300\n")
        for z in zs:
            mu = 5.0 * np.log10((1 + z) * z * 4283.0) + 25.0 - 19.3 + 0.15 * rng.normal()      # roughly the Hubble diagram
            s_, c_ = 1.0 + 0.1 * rng.normal(), 0.1 * rng.normal()
            v = [0.01 + 0.01 * rng.random(), 0.002 + 0.002 * rng.random(), 0.003 + 0.002 * rng.random(), 5e-4, 4e-4, 1e-3]
            f.write(" ".join("%.17g" % t for t in [z, mu, s_, c_] + v) + "\n")
    spec = T.TargetSpec(["Omega_m", "w_0_de", "M", "alpha", "beta"], [0.0, -3.5, 19.1, 0.5, -3.5],
                        [1.2, 0.5, 19.8, 2.6, -0.8]).add_snia(table=str(path))
    pmc = pmc_factory()
    pmc.set_target(spec)
    w, m, cov = T.proposal_sn(10)
    N = 6000
    k = rng.integers(0, 10, N)
    X = m[k] + np.einsum("nij,nj->ni", np.linalg.cholesky(cov)[k], rng.normal(size=(N, 5)))
    lo, hi = spec.box
    X = X[((X >= lo) & (X <= hi)).all(axis=1)]
    ref, eref = oracle.posterior_log_pdf(spec, X)
    with environ({"PMCB200_SN_WARP_MAX": "0"}):
        pmc.counters()
        got, egot = pmc.posterior_log_pdf(dev(X))
        cnt = pmc.counters()
    assert cnt["sn_spec"] > 0.9 * len(X), cnt
    with environ({"PMCB200_SN_EXACT": "1", "PMCB200_SN_WARP_MAX": "0"}):
        ex, eex = pmc.posterior_log_pdf(dev(X))
    got, egot, ex, eex = got.cpu().numpy(), egot.cpu().numpy(), ex.cpu().numpy(), eex.cpu().numpy()
    assert np.array_equal(egot != 0, eref != 0) and np.array_equal(eex != 0, eref != 0)
    ok = eref == 0
    # log pi = -chi2 / 2 + constants of a handful of supernovae can come out near zero: absolute floor of 1e-12 on top
    assert np.max(np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-2)) < RTOL_LOG
    assert np.max(np.abs(got[ok] - ex[ok]) / np.maximum(np.abs(ex[ok]), 1e-2)) < 1e-11


def test_sn_fast_path_matches_libdevice_path(oracle, pmc_factory, tmp_path):
    """The SN kernel's table-based exp2 / MUFU-seeded rsqrt path against the same
    kernel forced through libdevice exp (PMCB200_SN_FORCE_SLOW=1, separate
    process): agreement far inside the parity tolerance."""
    import os, subprocess, sys
    spec = T.target_sn_bao_w0wa()       # w0-wa + curvature: exercises HASQ and !FLAT
    X = box_samples(spec, 4000, 21)
    np.save(tmp_path / "X.npy", X)
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from cosmopmc_b200 import targets as T\n"
        "from cosmopmc_b200.pmc import PMC\n"
        "spec = T.target_sn_bao_w0wa(); pmc = PMC(0); pmc.set_target(spec)\n"
        "X = torch.from_numpy(np.load(%r)).cuda()\n"
        "lp, err = pmc.posterior_log_pdf(X)\n"
        "np.save(%r, np.stack([lp.cpu().numpy(), err.cpu().numpy().astype(float)]))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "X.npy"),
         str(tmp_path / "slow.npy"))
    env = dict(os.environ, PMCB200_SN_FORCE_SLOW="1")
    subprocess.check_call([sys.executable, "-c", code], env=env)
    slow = np.load(tmp_path / "slow.npy")
    pmc = pmc_factory()
    pmc.set_target(spec)
    lp, err = pmc.posterior_log_pdf(dev(X))
    lp, err = lp.cpu().numpy(), err.cpu().numpy()
    assert np.array_equal(err != 0, slow[1] != 0)
    ok = err == 0
    assert ok.sum() > 2000
    assert rel(lp[ok], slow[0][ok]) < 1e-12


# ------------------------------------------------- Fisher matrix (go_fishing.c) -------
def _fisher_stencil(pos, h, diag_only):
    """The reference's stencil order (go_fishing.c:37-85, elements (a, b >= a))."""
    diff = [(+1, +1), (+1, -1), (-1, +1), (-1, -1)]
    pts, elems = [], []
    d = len(pos)
    for a in range(d):
        for b in range(a, d):
            if diag_only and a != b:
                continue
            idx = []
            for j in range(4):
                if j == 2 and a == b:
                    idx.append(idx[1]); continue
                p = np.array(pos, dtype=np.float64)
                p[a] += diff[j][0] * h[a]
                p[b] += diff[j][1] * h[b]
                idx.append(len(pts)); pts.append(p)
            elems.append((a, b, idx))
    return np.array(pts), elems


@pytest.mark.parametrize("diag_only", [False, True])
def test_fisher_matrix_batched(oracle, pmc_factory, diag_only):
    """pmcb200_fisher_host: all 4 d(d+1)/2 stencil points of go_fishing.c's fisher_element in one launch.
    (1) Gaussian target: F = Sigma^-1 (central differences are exact on a quadratic); (2) SN Ia demo at the
    test suite's fiducial, fh = 0.001 (mkmax.h:33): against the same stencil evaluated by the oracle."""
    pmc = pmc_factory()
    d = 6
    rng = np.random.default_rng(5)
    A_ = rng.normal(size=(d, d)) * 0.3
    cov = A_ @ A_.T + np.eye(d)
    lo, hi = -8.0 * np.ones(d), 8.0 * np.ones(d)
    spec = T.TargetSpec(["dummy%d" % j for j in range(d)], lo, hi).add_mix([1.0], [0.2 * np.ones(d)], [cov])
    pmc.set_target(spec)
    pos, h = 0.1 * np.arange(d), 0.01 * (hi - lo)
    F = pmc.fisher_matrix(pos, h, diag_only)
    ref = np.linalg.inv(cov)
    if diag_only:
        ref = np.diag(np.diag(ref))
    assert np.allclose(F, F.T, rtol=0, atol=0)
    assert np.max(np.abs(F - ref)) < 1e-6 * np.max(np.abs(ref))
    # SN Ia
    spec = T.target_sn_demo()
    pmc.set_target(spec)
    pos = np.array([0.27, -1.0, 19.31, 1.6, -1.8])
    lo, hi = spec.box
    h = 0.001 * (hi - lo)
    F = pmc.fisher_matrix(pos, h, diag_only)
    pts, elems = _fisher_stencil(pos, h, diag_only)
    lp, err = oracle.posterior_log_pdf(spec, pts)
    assert not err.any()
    ref = np.zeros((5, 5))
    for a, b, idx in elems:
        c = lp[idx]
        ref[a, b] = ref[b, a] = -(c[0] - c[1] - c[2] + c[3]) / (4.0 * h[a] * h[b])
    scale = np.sqrt(np.abs(np.outer(np.diag(ref), np.diag(ref))))
    assert np.max(np.abs(F - ref) / scale) < 1e-5
    assert np.all(np.diag(F) > 0)
    # a stencil point outside the physical region fails loudly, as the reference forwards the error
    from cosmopmc_b200.pmc import PMCError
    bad = pos.copy(); bad[1] = 2.9
    with pytest.raises(PMCError):
        pmc.fisher_matrix(bad, h * 500.0, diag_only)
