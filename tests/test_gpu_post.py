"""GPU parity of the device post-processing (cosmopmc_b200/csrc/k_post.cu through the C-ABI:
pmcb200_post_moments / _sigma / _histogram) against the reference's own golden vectors
(tests/golden/post_ref.json), the numpy oracle, and -- where oracle/_ref travelled to this
machine -- the reference's compiled code (exec/exec_helper.c, tools/src/nhist.c)."""
import json
import os

import numpy as np
import pytest
import torch

from cosmopmc_b200 import targets as T
from oracle import post_oracle as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_ref.json")
unhex = float.fromhex


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


def test_post_golden_vectors_of_the_reference(pmc_factory):
    g = json.load(open(GOLD))
    X = np.array(g["X"]); w = np.array([unhex(v) for v in g["w"]]); flg = np.array(g["flg"], dtype=np.int16)
    pmc = pmc_factory()
    dX, dw, df = dev(X), dev(w), dev(flg)
    mean, cov = pmc.post_moments(dX, df, dw)
    center = np.array([unhex(v) for v in g["center"]])
    assert np.allclose(mean, center, rtol=1e-13, atol=0)
    m0, c0 = P.moments(X, w, flg)
    assert np.allclose(cov, c0, rtol=1e-10, atol=1e-18)
    for a in range(X.shape[1]):
        sig, med, nf = pmc.post_sigma(dX, df, dw, a, center[a], P.CONF_123_HALF)
        assert nf == int(flg.sum())
        assert np.array_equal(sig, np.array([unhex(v) for v in g["sigma"][a]]))      # bit-exact
        assert med == unhex(g["median"][a])
    n = int(flg.sum())
    for h in g["hist"]:
        cnt, sw, sw2 = pmc.post_histogram(dX, df, dw, h["pidx"], h["nbins"], h["limits"])
        data, var = P.hist_data_var(cnt, sw, sw2, n)
        rd = np.array([unhex(v) for v in h["data"]]); rv = np.array([unhex(v) for v in h["var"]])
        assert np.allclose(data, rd, rtol=1e-12, atol=0) and np.allclose(var, rv, rtol=1e-11, atol=0)
        assert cnt.sum() == np.count_nonzero(rd) or cnt.sum() >= np.count_nonzero(rd)


@pytest.mark.parametrize("d,N", [(5, 200001), (8, 50000), (22, 20000), (1, 1000)])
def test_post_against_oracle_random(pmc_factory, d, N):
    rng = np.random.default_rng(d * 1000 + 7)
    X = rng.standard_normal((N, d)) * rng.uniform(0.05, 2.0, d) + rng.uniform(-1, 20, d)
    w = rng.random(N) ** 3
    flg = (rng.random(N) > 0.2).astype(np.int16)
    X[flg == 0] = np.nan                      # unflagged rows may hold anything
    w[flg == 0] = 0.0
    w /= w.sum()
    pmc = pmc_factory()
    dX, dw, df = dev(X), dev(w), dev(flg)
    mean, cov = pmc.post_moments(dX, df, dw)
    m0, c0 = P.moments(X, w, flg)
    assert np.allclose(mean, m0, rtol=1e-12, atol=0)
    assert np.allclose(cov, c0, rtol=1e-9, atol=1e-12 * np.abs(c0).max())
    Xf, wf = X[flg != 0], w[flg != 0]
    ones = np.ones(len(Xf), np.int16)
    for a in sorted({0, d // 2, d - 1}):
        for center in (m0[a], float(Xf[:, a].min()) - 1.0, float(np.median(Xf[:, a]))):
            sig, med, nf = pmc.post_sigma(dX, df, dw, a, center, P.CONF_123_HALF)
            ref = P.ref_sigma(Xf, wf, ones, a, center) if P.ref() is not None else P.sigma(Xf, wf, ones, a, center)
            assert nf == len(Xf)
            # the interval ends are sample points: equal unless a prefix sum lands within rounding of the
            # confidence level (measure zero); compare exactly
            assert np.array_equal(sig, ref), (a, center, sig, ref)
        assert med == P.median(Xf, wf, ones, a)
    lo, hi = np.nanmin(X, 0), np.nanmax(X, 0)
    specs = [([0], [64], [lo[0] + 0.01, hi[0] - 0.01])]
    if d > 1:
        specs.append(([0, d - 1], [64, 64], [lo[0] - 1, hi[0] + 1, lo[d - 1] + 0.01, hi[d - 1] - 0.01]))
    for pidx, nb, lim in specs:
        cnt, sw, sw2 = pmc.post_histogram(dX, df, dw, pidx, nb, lim)
        c0_, s0_, q0_ = P.histogram(X, w, flg, pidx, nb, lim)
        assert np.array_equal(cnt, c0_)                                           # bin assignment is bit-exact
        assert np.allclose(sw, s0_, rtol=1e-12, atol=0) and np.allclose(sw2, q0_, rtol=1e-12, atol=0)


def test_post_on_a_pmc_iteration_sample(oracle, pmc_factory):
    """the use it is meant for: mean / covariance / intervals of the final weighted sample, on the
    device arrays the iteration left behind (no host round trip of X)"""
    spec = T.target_sn_demo()
    w, m, cov = T.proposal_sn(10)
    pmc = pmc_factory(); pmc.set_target(spec); pmc.set_proposal(w, m, cov=cov)
    N = 100000
    b = pmc.alloc(N)
    blk = torch.empty(pmc.stat_block_len(), dtype=torch.float64, device="cuda")
    pmc.iteration_local(N, 3, 0, 0, 1.0, blk, b)
    pmc.update_prop_rb(1, blk, N)
    pmc.normalize_importance_weight(b["flg"], b["logw"])
    mean, cv = pmc.post_moments(b["X"], b["flg"], b["logw"])
    X, wb, flg = b["X"].cpu().numpy(), b["logw"].cpu().numpy(), b["flg"].cpu().numpy()
    m0, c0 = P.moments(X, wb, flg)
    assert np.allclose(mean, m0, rtol=1e-12) and np.allclose(cv, c0, rtol=1e-9, atol=1e-16)
    sig, med, nf = pmc.post_sigma(b["X"], b["flg"], b["logw"], 0, mean[0], P.CONF_123_HALF)
    assert np.array_equal(sig, P.sigma(X, wb, flg, 0, mean[0])) and nf == int(flg.sum())
    assert 0.0 < sig[0] < sig[1] < sig[2] and sig[3] > 0.0       # 68% inside 95% inside 99% (the lower side may end at the box: -1)
