"""The DEVICE against the reference's own compiled code for the in-tree half of the posterior
(oracle/_ref/libref_param.so: unchanged wrappers/src/{param,sn,bao,wmap}.c + recording nicaea stand-in; built in
the container by oracle/build_ref_param.py, shipped prebuilt to the GPU box):

  * apply_params of the likelihood kernels (pmcb200_map_params) vs the reference's switch + set_base_parameters,
    BIT FOR BIT on random vectors, physical-density branch included, and the same error flags;
  * pmcb200_posterior_log_pdf vs the reference's posterior_log_pdf_common assembly (the per-probe log-likelihoods
    come from single-probe device calls and are fed through the recording chi2_* stand-ins): 1e-10 relative."""
import numpy as np
import pytest
import torch

from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T
from oracle import ref_param_lib as R

import ref_param_cases as RC

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_param.so not built (container only)")]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", RC.CASES, ids=[c[0] for c in RC.CASES])
def test_device_parameter_mapping_bitwise(pmc_factory, tmp_path, case):
    ref, spec = RC.build(case, tmp_path)
    pmc = pmc_factory()
    pmc.set_target(spec)
    X = RC.box_samples(case[2], case[3], 400, 7)
    X[0] = case[2]; X[1] = case[3]
    if "h_100" in case[1]:
        X[2:40, case[1].index("h_100")] = np.random.default_rng(1).choice([0.7, 0.73, 0.71, 0.6999999999999999, 1 / 3, 0.9], 38)
    R.set_returns(0.0, 0.0, 0.0, 0, 0)
    for i, (M, E) in enumerate(RC.ref_models(ref, spec, X)):
        md, ed = pmc.map_params(i, dev(X))
        md, ed = md.cpu().numpy(), ed.cpu().numpy()
        assert np.array_equal(ed != 0, E != 0), (case[0], i)
        ok = E == 0
        nc = 15 if spec.t.like[i].kind == A.LIKE["SNIa"] else 9
        cols = np.arange(nc)
        if "log_beta" in case[1]:      # Theta2[2] = -exp(x): libdevice's exp against glibc's, one ulp apart at most
            assert np.max(np.abs(md[ok][:, 11] - M[ok][:, 11]) / np.abs(M[ok][:, 11])) < 3e-16
            cols = cols[cols != 11]
        assert np.array_equal(bits(md[ok][:, cols]), bits(M[ok][:, cols])), (case[0], i)
    ref.close()


@pytest.mark.parametrize("special", ["none", "de_conservative"])
def test_device_posterior_vs_reference_assembly(pmc_factory, tmp_path, special):
    """C5 (CMB distance priors + BAO + SN): sum of the probes + logpr_default + special prior, assembled by the
    reference's posterior_log_pdf_common from the device's own per-probe values."""
    case = RC.CASES[3]
    ref, spec = RC.build(case, tmp_path, special=special)
    lo, hi = np.array(case[2]), np.array(case[3])
    if special == "de_conservative":
        lo[4], hi[4] = -0.95, -0.4
    X = RC.box_samples(lo, hi, 64, 5)
    pmc = pmc_factory()
    pmc.set_target(spec)
    got, err = pmc.posterior_log_pdf(dev(X))
    got, err = got.cpu().numpy(), err.cpu().numpy()
    # single-probe device values: a target holding only probe i, with `unity` so that no prior term is added
    single = []
    for i, probe in enumerate(case[4]):
        s1 = T.TargetSpec(case[1], case[2], case[3])
        {"SNIa": lambda: s1.add_snia(Theta2=RC.THETA2, cosmo=T.COSMO_DP, special="unity"),
         "BAO": lambda: s1.add_bao(T.BAO_BOSS12_DZ, cosmo=T.COSMO_DP, special="unity"),
         "CMBDistPrior": lambda: s1.add_cmbdp(cosmo=T.COSMO_DP, special="unity")}[probe]()
        p1 = pmc_factory()
        p1.set_target(s1)
        v, e = p1.posterior_log_pdf(dev(X))
        single.append((v.cpu().numpy(), e.cpu().numpy()))
    n_ok = 0
    for n, x in enumerate(X):
        if any(e[n] for _, e in single):
            assert err[n] != 0
            continue
        vals = dict(zip(case[4], [v[n] for v, _ in single]))
        R.set_returns(vals["SNIa"], vals["BAO"], vals["CMBDistPrior"], 0, 0)
        r, e, _ = ref.posterior(x)
        assert e == 0 and err[n] == 0
        assert abs(got[n] - r) <= 1e-10 * abs(r), (n, got[n], r)
        n_ok += 1
    assert n_ok > 40
    ref.close()
