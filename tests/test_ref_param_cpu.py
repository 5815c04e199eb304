"""The oracle against the REFERENCE'S OWN compiled code for the in-tree half of the posterior
(oracle/_ref/libref_param.so: unchanged wrappers/src/{param,sn,bao,wmap}.c, recording nicaea stand-in):

  * parameter mapping  sn.c:167-224, bao.c:100-147, wmap.c:966-1019 + set_base_parameters param.c:1544-1661
    -- BIT FOR BIT on random vectors, including the physical-density branch (x/h/h, param.c:1567-1572)
  * error conditions   tls_cosmo_par (mixing, overdetermined), ce_infnan (wmap.c:965), wmap_de_prior
  * posterior assembly posterior_log_pdf_common param.c:958-1041 with logpr_default (param.c:124-129) and
    prior_log_pdf_special (param.c:1055-1101) -- BIT FOR BIT, feeding the oracle's per-probe log-likelihoods
    through the recording chi2_* stand-ins
  * the order of SetDl / test_range_de_conservative / chi2_* in every probe

CPU only.  Skipped where the prebuilt library is absent."""
import ctypes as C

import numpy as np
import pytest

from cosmopmc_b200 import _abi as A
from oracle import ref_param_lib as R

import ref_param_cases as RC

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_param.so not built (container only)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("case", RC.CASES, ids=[c[0] for c in RC.CASES])
def test_parameter_mapping_bitwise(oracle, tmp_path, case):
    ref, spec = RC.build(case, tmp_path)
    X = RC.box_samples(case[2], case[3], 400, 7)
    X[0] = case[2]; X[1] = case[3]                      # the box corners
    if "h_100" in case[1]:
        X[2:40, case[1].index("h_100")] = np.random.default_rng(1).choice([0.7, 0.73, 0.71, 0.6999999999999999, 1 / 3, 0.9], 38)
    R.set_returns(0.0, 0.0, 0.0, 0, 0)
    for i, (M, E) in enumerate(RC.ref_models(ref, spec, X)):
        mo, eo = oracle.map_params(spec, i, X)
        assert np.array_equal(eo != 0, E != 0), (case[0], i)
        ok = E == 0
        if case[0].startswith("err_"):
            assert (~ok).sum() > 10
        else:
            assert ok.sum() > 0 or case[0] == "physical_no_h"
        nc = 15 if spec.t.like[i].kind == A.LIKE["SNIa"] else 9      # the cosmo_SN part exists for the SN probe only
        assert np.array_equal(bits(mo[ok][:, :nc]), bits(M[ok][:, :nc])), (case[0], i, np.abs(mo[ok][:, :nc] - M[ok][:, :nc]).max())
    ref.close()


def test_h_division_is_two_divisions(oracle, tmp_path):
    """param.c:1567-1572 divides omega_x / h100 / h100; the single division by h^2 differs in the last bit for
    a fraction of inputs -- the comparison must be sensitive to exactly that."""
    case = RC.CASES[4]
    ref, spec = RC.build(case, tmp_path)
    X = RC.box_samples([0.05, 0.01, 0.4, -1.0], [0.3, 0.05, 1.0, -0.9], 2000, 3)
    M, E = RC.ref_models(ref, spec, X)[1]
    one_div = X[:, 0] / (X[:, 2] * X[:, 2])
    assert (bits(one_div) != bits(M[:, 0])).sum() > 50           # the two forms do differ ...
    mo, _ = oracle.map_params(spec, 1, X)
    assert np.array_equal(bits(mo[:, 0]), bits(M[:, 0]))         # ... and the oracle follows the reference
    ref.close()


@pytest.mark.parametrize("special", ["none", "unity", "de_conservative"])
@pytest.mark.parametrize("ci", [0, 2, 3], ids=["sn_demo", "c4", "c5"])
def test_posterior_assembly_bitwise(oracle, tmp_path, ci, special):
    case = RC.CASES[ci]
    ref, spec = RC.build(case, tmp_path, special=special)
    assert bits([ref.logpr_default])[0] == bits([-np.sum(np.log(spec.box[1] - spec.box[0]))])[0] or \
        abs(ref.logpr_default + np.sum(np.log(spec.box[1] - spec.box[0]))) < 1e-14
    lo, hi = np.array(case[2]), np.array(case[3])
    if special == "de_conservative" and "w_0_de" in case[1]:     # inside the de_conservative range: no probe cuts
        j = case[1].index("w_0_de")
        lo[j], hi[j] = -0.95, -0.4
        if "w_1_de" in case[1]:
            lo[case[1].index("w_1_de")], hi[case[1].index("w_1_de")] = -0.1, 0.1
    X = RC.box_samples(lo, hi, 60, 11)
    lp, err = oracle.posterior_log_pdf(spec, X)
    L = oracle.lib()
    n_ok = 0
    for n, x in enumerate(X):
        ll = {}
        e_any = 0
        for i in range(spec.t.ndata):
            e = C.c_int(0)
            v = L.orc_loglike(C.byref(spec.t.like[i]), x.ctypes.data_as(C.c_void_p), C.byref(e))
            e_any |= e.value
            ll[spec.t.like[i].kind] = v
        if e_any:
            assert err[n] != 0
            continue
        R.set_returns(ll.get(A.LIKE["SNIa"], 0.0), ll.get(A.LIKE["BAO"], 0.0), ll.get(A.LIKE["CMBDistPrior"], 0.0), 0, 0)
        r, e, rec = ref.posterior(x)
        assert e == 0 and err[n] == 0
        assert bits([r])[0] == bits([lp[n]])[0], (case[0], special, r, lp[n])
        n_ok += 1
    assert n_ok > 30
    ref.close()


def test_special_prior_errors_and_values():
    """prior_log_pdf_special (param.c:1055-1101): the too-narrow w0 range is refused; values for w0 alone and w0 + w1"""
    P = A.P
    r, e = R.prior_special(A.SPECIAL["de_conservative"], [P["Omegam"], P["w0de"]], [0.0, -0.8], [1.0, 0.0])
    assert e != 0
    r, e = R.prior_special(A.SPECIAL["de_conservative"], [P["Omegam"], P["w0de"]], [0.0, -2.0], [1.0, 0.0])
    assert e == 0 and r == np.log(0.0 - -2.0) - np.log(2.0 / 3.0)
    r, e = R.prior_special(A.SPECIAL["de_conservative"], [P["w1de"], P["w0de"]], [-1.0, -2.0], [1.0, 0.0])
    a_acc = 2.0 / 3.0
    assert e == 0 and abs(r - (np.log(2.0) + np.log(2.0) - np.log(0.5 * 2 / 3 * 2 / 3 / (1 - a_acc)) - np.log(0.5 * 2 / 3 * 2 / 3))) < 1e-15
    r, e = R.prior_special(A.SPECIAL["unity"], [P["Omegam"], P["M"]], [0.0, 19.0], [1.0, 20.5])
    assert e == 0 and r == np.log(1.0) + np.log(1.5)
    r, e = R.prior_special(A.SPECIAL["de_conservative"], [P["Omegam"]], [0.0], [1.0])
    assert e == 0 and r == 0.0


def test_probe_call_order_and_de_conservative(oracle, tmp_path):
    """What each probe does with a model outside the de_conservative range (test_range_de_conservative = 1):
    likeli_SNIa runs SetDl first, then returns 0 without chi2_SN (sn.c:260-274) -- a SetDl error still is one;
    likeli_BAO returns 0 without any distance (bao.c:154-176); likeli_CMBDistPrior raises wmap_de_prior
    (wmap.c:1027-1044).  The oracle must agree on values and error flags."""
    case = RC.CASES[3]
    ref, spec = RC.build(case, tmp_path, special="de_conservative")
    names = {v: k for k, v in R.FN.items()}
    x_out = np.array([0.045, 0.27, 0.73, 0.71, -1.2, 19.31, 1.4, -2.4])        # w0 < -1: outside
    x_in = np.array([0.045, 0.27, 0.73, 0.71, -0.8, 19.31, 1.4, -2.4])
    R.set_returns(-100.0, -10.0, -1.0, 1, 0)
    r, e, rec = ref.likeli(2, x_out)          # SNIa
    assert e == 0 and r == 0.0 and [int(q[0]) for q in rec] == [names["SetDl"], names["test_range_de_conservative"]]
    r, e, rec = ref.likeli(1, x_out)          # BAO
    assert e == 0 and r == 0.0 and [int(q[0]) for q in rec] == [names["test_range_de_conservative"]]
    r, e, rec = ref.likeli(0, x_out)          # CMBDistPrior
    assert e != 0 and [int(q[0]) for q in rec] == [names["test_range_de_conservative"]]
    R.set_returns(-100.0, -10.0, -1.0, 1, 1)
    r, e, rec = ref.likeli(2, x_out)
    assert e != 0                              # SetDl's error precedes the cut
    R.set_returns(-100.0, -10.0, -1.0, 0, 0)
    r, e, rec = ref.likeli(2, x_in)
    assert e == 0 and r == -100.0 and [int(q[0]) for q in rec] == [names["SetDl"], names["test_range_de_conservative"], names["chi2_SN"]]
    r, e, rec = ref.likeli(0, x_in)
    assert e == 0 and r == -1.0 and [int(q[0]) for q in rec][-1] == names["chi2_cmbDP"]
    # the oracle on the same two points, probe by probe
    L = oracle.lib()
    def orc(i, x):
        ee = C.c_int(0)
        v = L.orc_loglike(C.byref(spec.t.like[i]), x.ctypes.data_as(C.c_void_p), C.byref(ee))
        return v, ee.value
    assert orc(2, x_out) == (0.0, 0) and orc(1, x_out) == (0.0, 0) and orc(0, x_out)[1] != 0
    assert all(orc(i, x_in)[1] == 0 and orc(i, x_in)[0] != 0.0 for i in range(3))
    # wmap.c:965: a non-finite parameter is ce_infnan for the CMB probe
    xb = x_in.copy(); xb[1] = np.nan
    R.set_returns(0.0, 0.0, 0.0, 0, 0)
    assert ref.likeli(0, xb)[1] != 0 and orc(0, xb)[1] != 0
    ref.close()
