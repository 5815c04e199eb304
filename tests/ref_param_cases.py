"""Shared case table for the parity tests against the reference's compiled parameter mapping
(oracle/_ref/libref_param.so = unchanged wrappers/src/{param,sn,bao,wmap}.c + recording nicaea stand-in).

Each case builds BOTH descriptions of the same posterior: a CosmoPMC config file for the reference's
read_config_base and a TargetSpec for the oracle / the device."""
import os

import numpy as np

from cosmopmc_b200 import targets as T
from oracle import ref_param_lib as R

THETA2 = (19.31, 1.6, -1.8, 0.0)
# (name, spar, min, max, probes) -- probes in data-set order; ranges deliberately cross zero where the
# reference's `> 0` tests decide between a sampled value and the default model
CASES = [
    ("sn_demo", ["Omega_m", "w_0_de", "M", "alpha", "beta"], [0.0, -3.5, 19.1, 0.5, -3.5], [1.2, 0.5, 19.8, 2.6, -0.8], ["SNIa"]),
    ("sn_curved", ["Omega_m", "Omega_de", "M", "alpha", "beta"], [-0.2, -0.2, 19.1, 0.5, -3.5], [1.2, 1.6, 19.8, 2.6, -0.8], ["SNIa"]),
    ("c4_sn_bao", ["Omega_m", "Omega_de", "w_0_de", "w_1_de", "M", "alpha", "beta"], [0.05, 0.0, -3.0, -3.0, 19.1, 0.5, -3.5],
     [1.0, 1.5, 0.0, 2.0, 19.8, 2.6, -0.8], ["SNIa", "BAO"]),
    ("c5_cmb_bao_sn", ["Omega_b", "Omega_m", "Omega_de", "h_100", "w_0_de", "M", "alpha", "beta"],
     [0.02, 0.1, 0.3, 0.5, -2.5, 19.1, 0.5, -3.5], [0.08, 0.6, 1.1, 0.9, -0.3, 19.8, 2.6, -0.8], ["CMBDistPrior", "BAO", "SNIa"]),
    ("physical", ["omega_m", "omega_b", "h_100", "w_0_de"], [-0.02, -0.005, 0.4, -2.0], [0.3, 0.05, 1.0, -0.3], ["CMBDistPrior", "BAO", "SNIa"]),
    ("physical_100ob_K", ["omega_m", "100_omega_b", "omega_K", "h_100"], [0.05, 1.0, -0.1, 0.4],
     [0.3, 4.0, 0.1, 1.0], ["CMBDistPrior", "BAO", "SNIa"]),
    ("physical_de_K_b_c", ["omega_de", "omega_K", "omega_b", "omega_c", "h_100"], [-0.1, -0.1, -0.005, -0.02, 0.4],
     [0.6, 0.1, 0.05, 0.3, 1.0], ["CMBDistPrior", "BAO", "SNIa"]),
    ("physical_c_b_nu", ["omega_c", "omega_b", "omega_nu_mass", "N_eff_nu_mass", "h_100"], [-0.02, -0.005, -0.001, 0.0, 0.4],
     [0.3, 0.05, 0.01, 4.0, 1.0], ["CMBDistPrior", "BAO", "SNIa"]),
    ("physical_no_h", ["omega_m", "omega_b"], [0.05, 0.01], [0.3, 0.05], ["BAO", "CMBDistPrior"]),
    ("Omega_c_b", ["Omega_c", "Omega_b", "Omega_K"], [-0.05, -0.01, -0.2], [0.6, 0.1, 0.2], ["BAO", "CMBDistPrior", "SNIa"]),
    ("Omega_de_K_nu", ["Omega_de", "Omega_K", "Omega_nu_mass"], [-0.1, -0.2, -0.01], [1.0, 0.2, 0.05], ["BAO", "CMBDistPrior", "SNIa"]),
    ("Omega_m_c", ["Omega_m", "Omega_c", "w_1_de"], [-0.1, -0.1, -1.0], [0.8, 0.6, 1.0], ["BAO", "CMBDistPrior", "SNIa"]),
    ("err_mixed", ["Omega_m", "omega_b"], [-0.1, -0.01], [0.8, 0.05], ["BAO", "CMBDistPrior", "SNIa"]),
    ("err_overdetermined", ["Omega_m", "Omega_de", "Omega_K"], [-0.1, 0.1, -0.2], [0.8, 1.0, 0.2], ["BAO", "CMBDistPrior", "SNIa"]),
    ("err_matter_overdetermined", ["Omega_m", "Omega_b", "Omega_c"], [-0.1, -0.01, -0.1], [0.8, 0.1, 0.6], ["BAO", "CMBDistPrior", "SNIa"]),
    ("sn_nuisance", ["Omega_m", "M", "alpha", "log_beta", "beta_z", "stretch", "color"], [0.0, 19.1, 0.5, -0.5, -1.0, 0.8, -0.2],
     [1.2, 19.8, 2.6, 1.5, 1.0, 1.2, 0.2], ["SNIa"]),
]


def box_samples(lo, hi, N, seed):
    lo, hi = np.array(lo), np.array(hi)
    return lo + np.random.default_rng(seed).random((N, len(lo))) * (hi - lo)


def build(case, tmp, special="none", sn_mode="chi2_simple", bao="dz"):
    """-> (RefConfig, TargetSpec).  The default models are T.COSMO_SN (SN) / T.COSMO_DP (BAO, CMB) on the oracle
    side; the reference's *_to_default* constructors are one function, so both probes get COSMO_DP there and on
    the TargetSpec when a case mixes probes."""
    name, spar, lo, hi, probes = case
    tmp = str(tmp)
    cosmo = T.COSMO_DP
    R.set_default_model(cosmo, THETA2)
    bao_data = {"dz": T.BAO_BOSS12_DZ, "A": T.BAO_REID10_A,
                "ratio": dict(method="distance_D_V_ratio", mean=[1.736, 1.52], covinv=[[260.0, -40.0], [-40.0, 400.0]],
                              z=[0.35, 0.2, 0.57, 0.35])}[bao]
    R.gauss_file(os.path.join(tmp, "bao.dat"), bao_data["mean"], bao_data["covinv"], bao_data["z"])
    R.gauss_file(os.path.join(tmp, "cmb.dat"), T.WMAP7_DP["mean"], T.WMAP7_DP["covinv"])
    spec = T.TargetSpec(spar, lo, hi)
    data = []
    for p in probes:
        if p == "SNIa":
            data.append(R.sn_section(chi2mode=sn_mode, special=special, Theta2_denom=(1.5, -2.0)))
            spec.add_snia(chi2mode=sn_mode, Theta2=THETA2, Theta2_denom=(0.0, 1.5, -2.0), cosmo=cosmo, special=special)
        elif p == "BAO":
            data.append(R.bao_section(os.path.join(tmp, "bao.dat"), bao_data["method"], special))
            spec.add_bao(bao_data, cosmo=cosmo, special=special)
        else:
            data.append(R.cmbdp_section(os.path.join(tmp, "cmb.dat"), special))
            spec.add_cmbdp(cosmo=cosmo, special=special)
    cfg = os.path.join(tmp, "config_%s" % name)
    R.write_config(cfg, spar, lo, hi, data)
    return R.RefConfig(cfg), spec


def ref_models(ref, spec, X):
    """per data set: (N, 15) mapped models recorded by the reference + error flags (record of the FIRST nicaea
    entry point the probe reaches)"""
    out = []
    for i in range(spec.t.ndata):
        M = np.full((len(X), 15), np.nan)
        E = np.zeros(len(X), dtype=np.int32)
        for n, x in enumerate(X):
            _, e, rec = ref.likeli(i, x)
            E[n] = e
            if len(rec):
                M[n] = R.record_model(rec[0])
        out.append((M, E))
    return out
