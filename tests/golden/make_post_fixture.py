#!/usr/bin/env python
"""Generates tests/golden/post_ref.json from the REFERENCE'S OWN compiled post-processing code
(oracle/_ref/libref_post.so = exec/exec_helper.c + tools/src/nhist.c, recipe
oracle/build_ref_post.py).  Container-only (needs /root/reference); the fixture it writes is
committed so the pin travels to machines without the reference tree.

The sample mimics a pmcsim file read back by pmc_simu_from_file: the flagged points first,
an unflagged tail, normalised weights."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref_post, post_oracle as P   # noqa: E402


def sample(seed=20240229, N=400, nflag=360, d=3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, d)) * [0.08, 0.25, 0.05] + [0.29, -1.0, 19.3]
    w = rng.random(N) ** 4
    flg = np.zeros(N, np.int16)
    flg[:nflag] = 1
    w[nflag:] = 0.0
    w /= w.sum()
    return np.round(X, 12), w, flg


def main():
    assert build_ref_post.build() and P.ref() is not None, "needs /root/reference"
    X, w, flg = sample()
    n = int(flg.sum())
    mean, _ = P.moments(X, w, flg)
    out = dict(source="oracle/_ref/libref_post.so: sigma_from_psim, median_from_psim (exec/exec_helper.c:164-275), "
                      "acc_histogram (tools/src/nhist.c:87-162)",
               X=X.tolist(), w=[float.hex(v) for v in w], flg=flg.tolist(), center=[float.hex(v) for v in mean],
               sigma=[], median=[], hist=[])
    for a in range(X.shape[1]):
        out["sigma"].append([float.hex(v) for v in P.ref_sigma(X, w, flg, a, mean[a])])
        out["median"].append(float.hex(P.ref_median(X, w, flg, a)))
    for pidx, nbins, limits in (([0], [16], [0.05, 0.55]), ([1, 2], [8, 6], [-1.7, -0.3, 19.15, 19.45])):
        data, var, total = P.ref_histogram(X[:n], w[:n], pidx, nbins, limits)
        out["hist"].append(dict(pidx=pidx, nbins=nbins, limits=limits, data=[float.hex(v) for v in data],
                                var=[float.hex(v) for v in var], total=float.hex(total)))
    with open(os.path.join(ROOT, "tests", "golden", "post_ref.json"), "w") as f:
        json.dump(out, f)
    print("wrote post_ref.json")


if __name__ == "__main__":
    main()
