#!/usr/bin/env python
"""Generate tests/golden/test_suite_logpost.json: the log-posterior at the fiducial points of the reference's own
regression recipe (bin/test_suite_cosmo_pmc.pl:51-58 `@fid`, :431-452 `max_post -m n`) for the in-scope demo
directories, computed by the CPU oracle on the SAME configurations (Demo/MC_Demo/{SN,BAO/distance_A,BAO/distance_d_z,
WMAP_Distance_Priors}/config_pmc + the joint SN+BAO set of COSMOS-S10+SN+BAO without its out-of-scope lensing probe).

The reference stores no expected values (the suite prints them for a human, :321-367), so these are oracle-made
goldens: tests/test_gpu_reference_driver.py runs the reference's UNCHANGED max_post binary (linked against this
repo's library) on the same config files and compares its printed maxlogP with them; a run against real nicaea can
be diffed against the same file.  Run in the container: python tests/golden/make_test_suite_fixture.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cosmopmc_b200 import targets as T          # noqa: E402
from oracle import oracle_lib as O              # noqa: E402


def suite():
    """name -> (TargetSpec, fiducial) with the boxes of the demo config files"""
    return {
        "SN": (T.target_sn_demo(), [0.27, -1.0, 19.31, 1.6, -1.8]),
        "BAO_distance_A": (T.TargetSpec(["Omega_m", "Omega_de"], [0.0, 0.0], [1.2, 2.0]).add_bao(T.BAO_REID10_A), [0.27, 0.73]),
        "BAO_distance_d_z": (T.TargetSpec(["Omega_m", "Omega_de"], [0.0, 0.0], [1.2, 2.0]).add_bao(T.BAO_BOSS12_DZ), [0.27, 0.73]),
        "WMAP_Distance_Priors": (T.TargetSpec(["Omega_b", "Omega_m", "Omega_de", "h_100"], [0.01, 0.15, 0.4, 0.5],
                                              [0.08, 0.45, 1.0, 0.9]).add_cmbdp(), [0.045, 0.27, 0.73, 0.71]),
        "SN+BAO": (T.TargetSpec(["Omega_m", "w_0_de", "h_100", "M", "alpha", "beta"], [0.0, -3.5, 0.4, 19.1, 0.5, -3.5],
                                [1.2, 0.5, 1.0, 19.8, 2.6, -0.8]).add_snia().add_bao(T.BAO_REID10_A),
                   [0.27, -1.0, 0.7, 19.31, 1.6, -1.8]),
    }


if __name__ == "__main__":
    out = {"source": "oracle/pmc_oracle.c orc_posterior_log_pdf at the fiducials of bin/test_suite_cosmo_pmc.pl:51-58 "
                     "(max_post -m n); generator tests/golden/make_test_suite_fixture.py", "cases": {}}
    for name, (spec, fid) in suite().items():
        lp, err = O.posterior_log_pdf(spec, np.array([fid]))
        assert err[0] == 0
        out["cases"][name] = {"spar": spec.spar, "fid": fid, "logpost": float(lp[0])}
        print(name, fid, repr(float(lp[0])))
    with open(os.path.join(HERE, "test_suite_logpost.json"), "w") as f:
        json.dump(out, f, indent=1)
