"""End-to-end: the reference's OWN, UNCHANGED PMC driver (exec/cosmo_pmc.c +
wrappers/ + tools/, compiled where they lie by tools/build_ref_cosmo_pmc.py and
linked against libpmc_b200.so) runs the SN Ia demo (Demo/MC_Demo/SN) on the GPU.

The binary and the demo inputs are built in the container (they derive from the
reference tree, so they are git-ignored) and travel to the GPU box; the test is
skipped where they are absent.  Checks are the reference's own documented
diagnostics: perplexity >= 0.8 and ENC >= 1.5 mean "explored sufficiently"
(README.md:166-169), and the posterior mean must agree with the SN posterior
printed in the manual (Manual/manual.tex:3226-3234)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T

pytestmark = pytest.mark.gpu
EXE = os.path.join(A.ROOT, "build_ref", "cosmo_pmc")
DEMO = os.path.join(A.ROOT, "build_ref", "demo_SN")


def write_fisher(path):
    """stands for max_post + go_fishing (exec/go_fishing.c): inverse of the manual's SN covariance"""
    F = np.linalg.inv(T.SN_POST_COV)
    with open(path, "w") as f:
        f.write("5 -1 5 0\n" + " ".join("%.10g" % v for v in T.SN_POST_MEAN) + "\n")
        for r in F:
            f.write(" ".join("%.10g" % v for v in r) + "\n")


@pytest.mark.skipif(not (os.path.exists(EXE) and os.path.isdir(DEMO)),
                    reason="build_ref/cosmo_pmc not built (tools/build_ref_cosmo_pmc.py, container only)")
def test_unchanged_reference_driver_runs_sn_demo(tmp_path, oracle):
    run = tmp_path / "run"
    shutil.copytree(DEMO, run, ignore=shutil.ignore_patterns("iter_*", "perplexity", "enc", "evidence*", "log_pmc",
                                                             "temperature", "proposal_fin", "run.log"))
    write_fisher(run / "fisher")
    out = subprocess.run([EXE, "-c", "config_pmc", "-s", "1", "-q"], cwd=run, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    perp = np.loadtxt(run / "perplexity")
    enc = np.loadtxt(run / "enc")
    assert perp.shape == (18, 5) and enc.shape == (19, 2)          # niter 18 (config_pmc:29)
    assert perp[-1, 1] == 190000                                     # 17 x 10000 + 2 x 10000 (fsfinal 2)
    assert perp[-1, 2] >= 0.8 and perp[0, 2] < perp[-1, 2]           # README.md:166-169
    assert enc[-1, 1] >= 1.5
    assert np.all(np.isfinite(np.loadtxt(run / "evidence")))
    # weighted mean of the final sample (exec/exec_helper.c:63-119) vs the manual's posterior
    rows = [l.split() for l in open(run / "iter_17" / "mean") if not l.startswith("#")]
    mean = np.array([float(r[2]) for r in rows])
    sig = np.sqrt(np.diag(T.SN_POST_COV))
    assert np.all(np.abs(mean - T.SN_POST_MEAN) < 1.0 * sig), (mean, T.SN_POST_MEAN, sig)
    # pmcsim format: 3 header lines, then log w, -component, 5 parameters (exec_helper.c:351-424)
    lines = open(run / "iter_17" / "pmcsim").read().split("\n")
    assert lines[0].startswith("# npar = 5, n_ded = 0") and len(lines[3].split()) == 7
    # resume path (cosmo_pmc.c:680-704): a second run re-reads iter_*/pmcsim + proposal instead of re-running
    out2 = subprocess.run([EXE, "-c", "config_pmc", "-s", "1", "-q"], cwd=run, capture_output=True, text=True, timeout=600)
    assert out2.returncode == 0, out2.stdout[-2000:] + out2.stderr[-2000:]
    perp2 = np.loadtxt(run / "perplexity")
    assert np.allclose(perp2[:, 2], perp[:, 2], rtol=2e-4)           # recomputed from the 9-digit text files
    # SURVEY 8f-1, importance_sample: the reference's unchanged tool (one scalar callback = one N=1 launch per
    # point) against this repo's batched driver (one launch per shard) on the first 1500 points of the final
    # sample, re-weighted under the same config: identical output files to the printed precision
    imp_ref, imp_b200 = os.path.join(A.ROOT, "build_ref", "importance_sample"), os.path.join(A.ROOT, "build_ref", "importance_sample_b200")
    if os.path.exists(imp_ref) and os.path.exists(imp_b200):
        with open(run / "sub.pmcsim", "w") as f:
            f.write("\n".join(lines[:3 + 1500]) + "\n")
        cfg = open(run / "config_pmc").read().replace("nsamples        10000", "nsamples        750")   # x fsfinal 2 = 1500 slots
        assert "nsamples        750" in cfg
        open(run / "config_is", "w").write(cfg)
        for exe, name, env in ((imp_ref, "ref.out", {}), (imp_b200, "b200.out", {}),
                               (imp_b200, "b200_sharded.out", {"PMCB200_NGPU": "2", "PMCB200_DEVICES": "0"})):
            o = subprocess.run([exe, "-c", "config_is", "-o", name, "-q", "sub.pmcsim"], cwd=run, capture_output=True,
                               text=True, timeout=900, env=dict(os.environ, **env))
            assert o.returncode == 0, o.stdout[-2000:] + o.stderr[-2000:]
        ref, b2, b3 = (np.loadtxt(run / n) for n in ("ref.out", "b200.out", "b200_sharded.out"))
        assert ref.shape == (1500, 7) and b2.shape == ref.shape and b3.shape == ref.shape
        assert np.allclose(b2, ref, rtol=1e-8, atol=0) and np.allclose(b3, ref, rtol=1e-8, atol=0)
        # log w_new - log w_old = log posterior: spot-check one row against the stored sample
        old = np.loadtxt(run / "sub.pmcsim")
        assert np.array_equal(old[:, 2:], ref[:, 2:]) and np.all(np.isfinite(ref[:, 0]))
        # ... and against the CPU oracle, so that this row is not only the device path against itself: the tool replaces
        # the weights by the log-posterior at the stored (9-digit) parameters, adds the previous log-weights and
        # normalises (importance_sample.c:75-77,273-290, exec_helper.c:408-420), i.e. new - old - log posterior is one
        # constant over the sample
        lp, elp = oracle.posterior_log_pdf(T.target_sn_demo(), np.ascontiguousarray(ref[:, 2:]))
        assert not elp.any()
        dlt = ref[:, 0] - old[:, 0] - lp
        assert np.max(np.abs(dlt - np.median(dlt))) < 5e-6, np.max(np.abs(dlt - np.median(dlt)))      # two %16.9g of ~ -180


@pytest.mark.skipif(not (os.path.exists(EXE) and os.path.isdir(DEMO)),
                    reason="build_ref/cosmo_pmc not built (tools/build_ref_cosmo_pmc.py, container only)")
def test_unchanged_reference_driver_sharded_over_contexts(tmp_path):
    """(e) under the UNCHANGED driver: PMCB200_NGPU shards every pmclib-named call over several
    contexts of the one process (here 3 contexts on device 0).  The Philox counter is the global
    sample index, so the run must reproduce the one-context run: same sample counts, perplexity
    and ENC histories equal to the files' printed precision."""
    res = {}
    for ns in (1, 3):
        run = tmp_path / ("run%d" % ns)
        shutil.copytree(DEMO, run, ignore=shutil.ignore_patterns("iter_*", "perplexity", "enc", "evidence*", "log_pmc",
                                                                 "temperature", "proposal_fin", "run.log"))
        write_fisher(run / "fisher")
        out = subprocess.run([EXE, "-c", "config_pmc", "-s", "7", "-q"], cwd=run, capture_output=True, text=True,
                             timeout=600, env=dict(os.environ, PMCB200_NGPU=str(ns), PMCB200_DEVICES="0"))
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
        res[ns] = (np.loadtxt(run / "perplexity"), np.loadtxt(run / "enc"), np.loadtxt(run / "evidence"))
    for a, b in zip(res[1], res[3]):
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-5, atol=0)


def _run_tempering_demo(name, fisher_mean, fisher_invcov, tmp_path, seed):
    demo = os.path.join(A.ROOT, "build_ref", "demo_" + name)
    if not (os.path.exists(EXE) and os.path.isdir(demo)):
        pytest.skip("build_ref not built (container only)")
    run = tmp_path / name
    shutil.copytree(demo, run, ignore=shutil.ignore_patterns("iter_*", "perplexity", "enc", "evidence*", "log_pmc",
                                                             "temperature", "proposal_fin"))
    with open(run / "fisher", "w") as f:        # stands for max_post + go_fishing at the peak
        d = len(fisher_mean)
        f.write("%d -1 %d 0\n" % (d, d) + " ".join("%g" % v for v in fisher_mean) + "\n")
        for r in fisher_invcov:
            f.write(" ".join("%g" % v for v in r) + "\n")
    out = subprocess.run([EXE, "-c", "config_pmc", "-s", str(seed), "-q"], cwd=run, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return np.loadtxt(run / "evidence"), np.loadtxt(run / "perplexity")


def test_reference_driver_tempering_demo_gaussian_evidence(tmp_path):
    """Demo/tempering/README.md:11-36: normalised 2-D Gaussian on the unit square =>
    the file `evidence` 'should be consistent with 1'."""
    evi, perp = _run_tempering_demo("1_mvnorm_2D_temp_none", [0.5, 0.5], [[100.0, 0.0], [0.0, 50.0]], tmp_path, 3)
    assert evi.shape[0] == 5
    assert abs(evi[-1, 3] - 1.0) < 0.05                # 5000 samples in the final iteration
    assert perp[-1, 2] > 0.8


def test_reference_driver_tempering_demo_mixture_evidence(tmp_path):
    """Demo/tempering/README.md:64-106: two separated modes => evidence approaches 1 (both
    modes found) or 0.5 (one mode); quoted runs 0.998802, 0.50055, 0.499927."""
    evi, perp = _run_tempering_demo("2_mixmvnorm_2D_temp_none", [0.2, 0.2], [[500.0, 0.0], [0.0, 500.0]], tmp_path, 5)
    e = evi[-1, 3]
    assert abs(e - 1.0) < 0.05 or abs(e - 0.5) < 0.03, e


def test_full_reference_pipeline_max_fisher_pmc(tmp_path):
    """bin/cosmo_pmc.pl:103-137 end to end with the reference's unchanged executables:
    max_post (amoeba) -> config_pmc_to_max_and_fish.pl -> go_fishing -> cosmo_pmc.  Every scalar
    likelihood call of max_post / go_fishing runs the batched CUDA kernels with N = 1."""
    ref = os.path.join(A.ROOT, "build_ref")
    pipe = os.path.join(ref, "demo_SN_pipeline")
    need = [os.path.join(ref, f) for f in ("cosmo_pmc", "max_post", "go_fishing", "config_pmc_to_max_and_fish.pl")]
    if not (all(os.path.exists(f) for f in need) and os.path.isdir(pipe) and shutil.which("perl")):
        pytest.skip("build_ref pipeline not built (container only)")
    run = tmp_path / "pipe"
    shutil.copytree(pipe, run, ignore=shutil.ignore_patterns("iter_*", "perplexity", "enc", "evidence*", "log_*",
                                                             "temperature", "proposal_fin", "maxlogP", "fisher",
                                                             "config_fish"))
    r = subprocess.run([need[1], "-t", "-m", "a", "-s", "1"], cwd=run, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and (run / "maxlogP").exists(), r.stdout[-2000:] + r.stderr[-2000:]
    tok = open(run / "maxlogP").read().split()
    maxlogp = float(tok[tok.index("=") + 1])
    best = np.array([float(v) for v in tok[-5:]])
    assert -182.0 < maxlogp < -176.0
    assert np.all(np.abs(best - T.SN_POST_MEAN) < 2.0 * np.sqrt(np.diag(T.SN_POST_COV)))
    with open(run / "config_fish", "w") as fo:
        subprocess.check_call(["perl", need[3], "-F", "-p", "maxlogP", "-c", "config_pmc"], cwd=run, stdout=fo)
    r = subprocess.run([need[2], "-q"], cwd=run, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and (run / "fisher").exists(), r.stdout[-2000:] + r.stderr[-2000:]
    tok = open(run / "fisher").read().split()
    F = np.array([float(v) for v in tok[9:34]]).reshape(5, 5)
    assert np.allclose(F, F.T, rtol=1e-3) and np.all(np.linalg.eigvalsh(F) > 0)
    # the Fisher matrix is the inverse posterior covariance to within the posterior's non-Gaussianity
    sig_f = np.sqrt(np.diag(np.linalg.inv(F)))
    assert np.all(sig_f < 1.5 * np.sqrt(np.diag(T.SN_POST_COV))) and np.all(sig_f > 0.3 * np.sqrt(np.diag(T.SN_POST_COV)))
    r = subprocess.run([need[0], "-c", "config_pmc", "-s", "1", "-q"], cwd=run, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    perp = np.loadtxt(run / "perplexity")
    assert perp[-1, 2] >= 0.8 and np.loadtxt(run / "enc")[-1, 1] >= 1.5


def test_reference_pipeline_wmap_distance_priors(tmp_path):
    """Demo/MC_Demo/WMAP_Distance_Priors (CMBDistPrior, 4 parameters, sdead_comp revive, nclipw 5) through
    max_post -> go_fishing -f -> cosmo_pmc, all from unchanged reference sources.  Exercises the CMB
    distance-prior kernel, clip_weights and revive_comp (which copies a component's std as a covariance)."""
    ref = os.path.join(A.ROOT, "build_ref")
    demo = os.path.join(ref, "demo_WMAP_DP")
    need = [os.path.join(ref, f) for f in ("cosmo_pmc", "max_post", "go_fishing", "config_pmc_to_max_and_fish.pl")]
    if not (all(os.path.exists(f) for f in need) and os.path.isdir(demo) and shutil.which("perl")):
        pytest.skip("build_ref pipeline not built (container only)")
    run = tmp_path / "dp"
    shutil.copytree(demo, run, ignore=shutil.ignore_patterns("iter_*", "perplexity", "enc", "evidence*", "log_*",
                                                             "temperature", "proposal_fin", "maxlogP", "fisher",
                                                             "config_fish", "run.log"))
    def sh(cmd, **kw):
        r = subprocess.run(cmd, cwd=run, capture_output=True, text=True, timeout=600, **kw)
        assert r.returncode == 0, " ".join(cmd) + "\n" + r.stdout[-2000:] + r.stderr[-2000:]
        return r
    sh([need[1], "-t", "-m", "a", "-s", "1", "-q"])
    best = np.array([float(v) for v in open(run / "maxlogP").read().split()[-4:]])
    assert np.all(np.abs(best - [0.045, 0.27, 0.73, 0.71]) < [0.01, 0.05, 0.05, 0.05])      # WMAP7 best fit region
    with open(run / "config_fish", "w") as fo:
        subprocess.check_call(["perl", need[3], "-F", "-p", "maxlogP", "-c", "config_pmc"], cwd=run, stdout=fo)
    sh([need[2], "-q", "-f"])                     # 3 data, 4 parameters: force a positive (diagonal) Fisher matrix
    r = sh([need[0], "-c", "config_pmc", "-s", "1", "-q"])
    perp = np.loadtxt(run / "perplexity")
    assert perp.shape[0] == 10 and perp[-1, 2] > 0.3 and perp[-1, 2] > perp[0, 2]
    rows = [l.split() for l in open(run / "iter_9" / "mean") if not l.startswith("#")]
    mean = np.array([float(r[2]) for r in rows])
    assert 0.03 < mean[0] < 0.07 and 0.2 < mean[1] < 0.45 and 0.55 < mean[2] < 0.85 and 0.5 < mean[3] < 0.85
    # clip_weights (cosmo_pmc.c:396-399, nclipw 5, reports on stderr): five points per iteration, and the mean
    # of the final iteration is the one of the CLIPPED sample -- recomputed here from the unclipped pmcsim file
    # (written at cosmo_pmc.c:392, before the clipping) with its five largest weights removed
    assert r.stderr.count("Clipping point") == 5 * 10, r.stderr[-2000:]
    sim = np.loadtxt(run / "iter_9" / "pmcsim")
    lw, X = sim[:, 0], sim[:, 2:6]
    w = np.exp(lw - lw.max())
    keep = np.ones(len(w), bool)
    keep[np.argsort(w)[-5:]] = False
    m_clip = (w[keep, None] * X[keep]).sum(0) / w[keep].sum()
    m_all = (w[:, None] * X).sum(0) / w.sum()
    assert np.allclose(mean, m_clip, rtol=2e-5), (mean, m_clip, m_all)       # text files carry 9 / 5 digits
    assert np.abs(m_all - m_clip).max() > 0.0


SUITE_DIR = os.path.join(A.ROOT, "build_ref", "test_suite")


@pytest.mark.parametrize("name", ["SN", "BAO_distance_A", "BAO_distance_d_z", "WMAP_Distance_Priors", "SN+BAO"])
def test_reference_test_suite_log_posterior_at_fiducial(tmp_path, name):
    """The reference's own regression recipe (bin/test_suite_cosmo_pmc.pl:51-58,321-367,431-452): `max_post -m n`
    evaluates the log-posterior at the suite's fiducial point.  The UNCHANGED max_post binary (reference
    exec/max_post.c + wrappers/ + tools/, linked against this library: the likelihood is one N = 1 launch of the
    batched CUDA kernels) must print the golden value (tests/golden/test_suite_logpost.json, oracle-made: the
    reference stores none) to the 6 digits of its `%g` (max_post.c:253)."""
    import json
    exe = os.path.join(A.ROOT, "build_ref", "max_post")
    src = os.path.join(SUITE_DIR, name)
    if not (os.path.exists(exe) and os.path.isdir(src)):
        pytest.skip("build_ref/max_post or build_ref/test_suite not built (container only)")
    gold = json.load(open(os.path.join(A.ROOT, "tests", "golden", "test_suite_logpost.json")))["cases"][name]
    run = tmp_path / "ts"
    shutil.copytree(src, run)
    r = subprocess.run([exe, "-m", "n", "-c", "config_max_test_suite"], cwd=run, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and (run / "maxlogP").exists(), r.stdout[-2000:] + r.stderr[-2000:]
    tok = open(run / "maxlogP").read().split()
    val = float(tok[tok.index("=") + 1])
    p = np.array([float(v) for v in tok[-len(gold["fid"]):]])
    assert np.allclose(p, gold["fid"], rtol=1e-6)                    # -m n: the point is not moved
    assert abs(val - gold["logpost"]) <= 6e-6 * abs(gold["logpost"]), (val, gold["logpost"])
