"""CPU tests of the post-processing oracle (oracle/post_oracle.py, numpy) against the
reference's OWN code: the committed golden vectors (tests/golden/post_ref.json, generated from
oracle/_ref/libref_post.so = exec/exec_helper.c + tools/src/nhist.c compiled unchanged) and,
where oracle/_ref exists, the compiled reference itself on random samples."""
import json
import os

import numpy as np
import pytest

from oracle import post_oracle as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_ref.json")
unhex = float.fromhex


def load_gold():
    g = json.load(open(GOLD))
    X = np.array(g["X"])
    w = np.array([unhex(v) for v in g["w"]])
    flg = np.array(g["flg"], dtype=np.int16)
    return g, X, w, flg


def test_oracle_matches_reference_golden_vectors():
    g, X, w, flg = load_gold()
    center = [unhex(v) for v in g["center"]]
    mean, cov = P.moments(X, w, flg)
    assert np.allclose(mean, center, rtol=1e-15)
    assert np.allclose(cov, np.cov(X[flg != 0].T, aweights=w[flg != 0], bias=True), rtol=1e-12)
    for a in range(X.shape[1]):
        ref = np.array([unhex(v) for v in g["sigma"][a]])
        assert np.array_equal(P.sigma(X, w, flg, a, center[a]), ref)          # bit-exact, incl. the -1 boundary code
        assert P.median(X, w, flg, a) == unhex(g["median"][a])
    n = int(flg.sum())
    for h in g["hist"]:
        cnt, sw, sw2 = P.histogram(X[:n], w[:n], None, h["pidx"], h["nbins"], h["limits"])
        data, var = P.hist_data_var(cnt, sw, sw2, n)
        rd = np.array([unhex(v) for v in h["data"]]); rv = np.array([unhex(v) for v in h["var"]])
        assert np.allclose(data, rd, rtol=1e-13, atol=0) and np.allclose(var, rv, rtol=1e-12, atol=0)
        assert abs(sw.sum() - unhex(h["total"])) < 1e-14


@pytest.mark.skipif(P.ref() is None, reason="oracle/_ref/libref_post.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_compiled_reference(seed):
    rng = np.random.default_rng(seed)
    N, d, nflag = 3000, 4, 2500 + 100 * seed
    X = rng.standard_normal((N, d)) * rng.uniform(0.1, 3.0, d) + rng.uniform(-2, 2, d)
    X[5] = X[6]                                   # ties
    w = rng.random(N) ** (2 + seed)
    flg = np.zeros(N, np.int16); flg[:nflag] = 1
    w[nflag:] = 0.0
    w /= w.sum()
    mean, _ = P.moments(X, w, flg)
    for a in range(d):
        for center in (mean[a], X[:, a].min() - 1.0, X[:nflag, a].max() + 1.0, np.median(X[:nflag, a])):
            assert np.array_equal(P.sigma(X, w, flg, a, center), P.ref_sigma(X, w, flg, a, center))
        assert P.median(X, w, flg, a) == P.ref_median(X, w, flg, a)
    lim = [X[:, 0].min() + 0.1, X[:, 0].max() - 0.1, X[:, 2].min() - 0.1, X[:, 2].max() + 0.1]
    for pidx, nb, li in (([0], [32], lim[:2]), ([0, 2], [12, 9], lim), ([2, 0], [9, 12], lim[2:] + lim[:2])):
        cnt, sw, sw2 = P.histogram(X[:nflag], w[:nflag], None, pidx, nb, li)
        data, var = P.hist_data_var(cnt, sw, sw2, nflag)
        rd, rv, tot = P.ref_histogram(X[:nflag], w[:nflag], pidx, nb, li)
        assert np.allclose(data, rd, rtol=1e-13, atol=0) and np.allclose(var, rv, rtol=1e-12, atol=0)
        assert abs(tot - sw.sum()) < 1e-13


def test_sigma_edge_cases():
    X = np.array([[0.0], [1.0], [2.0], [3.0]])
    w = np.full(4, 0.25); flg = np.ones(4, np.int16)
    # centre above every point: imean sticks at n-1 (exec_helper.c:221-225), right side hits the boundary
    s = P.sigma(X, w, flg, 0, 10.0, (0.2, 0.3, 0.6))
    assert list(s[:3]) == [-1.0, -1.0, -1.0] and s[3] == 10.0 - 1.0
    # a single flagged point: everything is a boundary hit
    s = P.sigma(X, w, np.array([0, 1, 0, 0], np.int16), 0, 1.0)
    assert np.all(s == -1.0)
    assert np.isnan(P.median(X, np.full(4, 0.1), flg, 0))       # weights never reach 0.5: err_median
