"""GPU tests at BASELINE.json's full size (N = 10^7, SN Ia C2 shape) through
size-independent properties, plus edge cases of the C-ABI (minimum / maximum
sizes, empty shards, all-rejected samples, error codes)."""
import numpy as np
import pytest
import torch

from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T
from cosmopmc_b200.pmc import PMCError

pytestmark = pytest.mark.gpu


def chol(cov):
    return np.stack([np.linalg.cholesky(c) for c in cov])


def test_full_size_iteration_properties(pmc_factory):
    """N = 10^7: determinism, shard invariance of the statistics, normalisation, consistency of
    the diagnostics with the returned weights, idempotent proposal read-back."""
    N = 10_000_000
    spec = T.target_sn_demo()
    w, m, cov = T.proposal_sn(10)
    ch = chol(cov)
    pmc = pmc_factory(); pmc.set_target(spec); pmc.set_proposal(w, m, chol=ch)
    bufs = pmc.alloc(N)
    blk = torch.zeros(pmc.stat_block_len(), dtype=torch.float64, device="cuda")
    pmc.iteration_local(N, 1, 0, 0, 1.0, blk, bufs)
    b1 = blk.clone()
    X1 = bufs["X"].clone(); lw1 = bufs["logw"].clone(); f1 = bufs["flg"].clone()
    pmc.iteration_local(N, 1, 0, 0, 1.0, blk, bufs)
    assert torch.equal(blk, b1) and torch.equal(bufs["X"], X1) and torch.equal(bufs["logw"], lw1)   # bit-deterministic
    st = pmc.update_prop_rb(1, blk, N)
    # box flags and counters
    lo, hi = spec.box
    inside = ((X1 >= torch.tensor(lo, device="cuda")) & (X1 <= torch.tensor(hi, device="cuda"))).all(dim=1)
    assert int(inside.sum()) == st["nok_box"] and int((f1 != 0).sum()) == st["nok"] <= st["nok_box"]
    # weights: max, log-sum-exp, normalisation, perplexity, ESS recomputed with torch from the returned log-weights
    ok = f1 != 0
    lw = lw1[ok]
    M = lw.max().item()
    e = torch.exp(lw - M)
    S = e.sum().item()
    assert st["maxW"] == M
    assert abs(st["logSum"] - (np.log(S) + M)) < 1e-11 * abs(st["logSum"])
    wbar = e / S
    H = -(wbar[wbar > 0] * torch.log(wbar[wbar > 0])).sum().item()
    assert abs(st["perplexity"] - np.exp(H) / N) < 1e-9 * st["perplexity"]
    assert abs(st["ess"] - 1.0 / (wbar * wbar).sum().item()) < 1e-9 * st["ess"]
    assert abs(st["ln_evidence"] - (st["logSum"] - np.log(N))) < 1e-12
    pmc.normalize_importance_weight(bufs["flg"], bufs["logw"], N)
    assert abs(bufs["logw"].sum().item() - 1.0) < 1e-10 and float(bufs["logw"][~ok].abs().max()) == 0.0
    # the updated proposal is a normalised mixture of PD components; ENC matches
    wg, mg, chg, covg = pmc.get_proposal()
    assert abs(wg.sum() - 1) < 1e-12 and abs(st["enc"] - 1 / np.sum(wg ** 2)) < 1e-12 * st["enc"]
    assert np.all(np.linalg.eigvalsh(covg[wg > 0]) > 0)
    # weighted moments of the sample reproduce the mixture's total mean (law of total expectation)
    xm = (wbar[:, None] * X1[ok]).sum(0).cpu().numpy()
    assert np.allclose(xm, (wg[:, None] * mg).sum(0), rtol=1e-9, atol=1e-12)
    # shard invariance: 3 ragged shards give the same combined statistics
    many = pmc_factory(); many.set_target(spec); many.set_proposal(w, m, chol=ch)
    cuts = [0, 3_333_333, 3_333_334, N]
    allb = torch.zeros((3, many.stat_block_len()), dtype=torch.float64, device="cuda")
    sb = many.alloc(cuts[1])
    for r in range(3):
        n = cuts[r + 1] - cuts[r]
        many.iteration_local(n, 1, 0, cuts[r], 1.0, allb[r], sb if n <= cuts[1] else None)
    s3 = many.update_prop_rb(3, allb, N)
    for k in ("nok", "nok_box", "ndead", "maxW"):
        assert s3[k] == st[k]
    for k in ("logSum", "perplexity", "ess", "enc"):
        assert abs(s3[k] - st[k]) <= 1e-11 * abs(st[k])
    for a, b in zip(many.get_proposal(), (wg, mg, chg, covg)):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-13)


def test_edge_sizes_and_limits(oracle, pmc_factory):
    pmc = pmc_factory()
    # d = 32, K = 4 (dimension limit) and d = 2, K = 64 (component limit) against the oracle
    for K, d in ((4, 32), (64, 2), (1, 3)):
        rng = np.random.default_rng(K * 7 + d)
        lo, hi = -5.0 * np.ones(d), 5.0 * np.ones(d)
        spec = T.TargetSpec(["dummy%d" % j for j in range(d)], lo, hi).add_mix(
            [1.0], [np.zeros(d)], [np.eye(d) * 0.8])
        mean = rng.normal(size=(K, d)) * 0.3
        cov = np.repeat((np.eye(d) * 1.5)[None], K, 0)
        w = np.full(K, 1.0 / K)
        pmc.set_target(spec); pmc.set_proposal(w, mean, cov=cov)
        N = 20000
        hX = np.empty((N, d)); hi_ = np.empty(N, np.int32); hf = np.empty(N, np.int16); hw = np.empty(N)
        st = pmc.iteration_host(N, 3, 0, 1.0, hX, hi_, hf, hw)
        o = oracle.iteration(spec, N, 3, 0, 1.0, w, mean, chol(cov), nthreads=8)
        assert st["nok"] == o["stats"]["nok"] and np.array_equal(hi_, o["idx"])
        assert abs(st["perplexity"] - o["stats"]["perplexity"]) < 1e-8 * st["perplexity"]
        wg, mg, _, covg = pmc.get_proposal()
        assert np.allclose(wg, o["wght"], rtol=1e-8, atol=1e-300) and np.allclose(mg[wg > 0], o["mean"][wg > 0], rtol=1e-7, atol=1e-9)
    # beyond the limits: loud errors, not truncation
    with pytest.raises(PMCError) as e:
        pmc.set_proposal(np.full(65, 1 / 65), np.zeros((65, 2)), cov=np.repeat(np.eye(2)[None], 65, 0))
    assert e.value.code == A.ERR["DIM"]
    with pytest.raises(PMCError) as e:
        pmc.set_proposal([1.0], np.zeros((1, 2)), cov=np.array([[[1.0, 2.0], [2.0, 1.0]]]))
    assert e.value.code == A.ERR["CHOLESKY"]


def test_single_sample_and_empty_shard(pmc_factory):
    spec = T.target_gauss2d()
    pmc = pmc_factory(); pmc.set_target(spec)
    pmc.set_proposal([1.0], [[0.5, 0.5]], cov=[np.eye(2) * 0.01])
    with pytest.raises(PMCError) as e:                    # N = 1: alpha < MINCOUNT draws => every component dies
        pmc.iteration_host(1, 5, 0)
    assert e.value.code == A.ERR["NOSAMPLE"] and pmc.last_stats["nok"] in (0, 1)
    # a rank with an empty shard contributes nothing but must not break the combine
    pmc.set_proposal([0.5, 0.5], [[0.45, 0.5], [0.55, 0.5]], cov=[np.eye(2) * 0.01] * 2)
    allb = torch.zeros((2, pmc.stat_block_len()), dtype=torch.float64, device="cuda")
    pmc.iteration_local(5000, 5, 0, 0, 1.0, allb[0])
    pmc.iteration_local(0, 5, 0, 5000, 1.0, allb[1])
    assert allb[1][0].item() == -np.inf and allb[1][1:].abs().sum().item() == 0.0
    st = pmc.update_prop_rb(2, allb, 5000)
    assert st["nok"] > 4000 and 0 < st["perplexity"] <= 1


def test_no_sample_in_box_is_an_error(pmc_factory):
    """cosmo_pmc.c:321-324: 'No point simulated' (pmc_nosamplep) when every draw is rejected."""
    spec = T.target_gauss2d()
    pmc = pmc_factory(); pmc.set_target(spec)
    pmc.set_proposal([1.0], [[50.0, 50.0]], cov=[np.eye(2) * 1e-4])     # far outside the unit box
    with pytest.raises(PMCError) as e:
        pmc.iteration_host(1000, 1, 0)
    assert e.value.code == A.ERR["NOSAMPLE"]
    assert pmc.last_stats["nok_box"] == 0


def test_call_order_errors(pmc_factory):
    pmc = pmc_factory()
    X = torch.zeros((4, 5), dtype=torch.float64, device="cuda")
    with pytest.raises(PMCError) as e:
        pmc.mix_mvdens_log_pdf(X)
    assert e.value.code == A.ERR["STATE"]
    pmc.set_target(T.target_sn_demo())
    w, m, cov = T.proposal_banana(3, 20)
    pmc.set_proposal(w, m, cov=cov)                         # 20-D proposal on a 5-D target
    with pytest.raises(PMCError) as e:
        pmc.iteration_host(100, 1, 0)
    assert e.value.code == A.ERR["DIM"]
    bad = T.target_sn_demo()
    bad.t.like[0].sn_chi2mode = 4                           # chi2_dust: no device path
    with pytest.raises(PMCError) as e:
        pmc.set_target(bad)
    assert e.value.code == A.ERR["UNSUP"]
