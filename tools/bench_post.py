#!/usr/bin/env python
"""Device post-processing of a weighted sample against its HBM roofline, with the
host code it replaces timed beside it.

  python tools/bench_post.py [--n 10000000] [--reps 5] [--cpu-n 2000000]

Sample = one SN-demo iteration left on the device.  Per stage: CUDA-event time, algorithmic bytes
(8d + 10 per sample and pass for the moments; 8 + 10 for one histogram axis / the gather), GB/s and
the fraction of the measured HBM peak (MEASURED_PEAKS.json).  CPU side: the reference's own
compiled sigma_from_psim / acc_histogram (oracle/_ref, exec_helper.c / nhist.c) where present,
else the numpy restatement, on --cpu-n samples of the same arrays (1 core, as the reference runs
post-processing on rank 0)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_config, SEED
from cosmopmc_b200.pmc import PMC
from oracle import post_oracle as P

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--cpu-n", type=int, default=2_000_000)
a = ap.parse_args()
peak = 6535.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
spec, w, m, ch, label = make_config("sn")
pmc = PMC(0); pmc.set_target(spec); pmc.set_proposal(w, m, chol=ch)
N, d = a.n, 5
b = pmc.alloc(N)
blk = torch.empty(pmc.stat_block_len(), dtype=torch.float64, device="cuda")
pmc.iteration_local(N, SEED, 0, 0, 1.0, blk, b)
pmc.update_prop_rb(1, blk, N)
pmc.normalize_importance_weight(b["flg"], b["logw"])
X, flg, wb = b["X"], b["flg"], b["logw"]
mean, cov = pmc.post_moments(X, flg, wb)
lim = [float(X[:, 0].min()) - 1e-9, float(X[:, 0].max()) + 1e-9, float(X[:, 1].min()) - 1e-9, float(X[:, 1].max()) + 1e-9]


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps

rows = []
def row(name, ms, nbytes, note=""):
    gbs = nbytes / (ms * 1e-3) * 1e-9
    rows.append(dict(stage=name, ms=ms, algorithmic_bytes=nbytes, GBs=gbs, frac_of_hbm_peak=gbs / peak, note=note))

row("post_moments (mean + covariance, 2 passes)", timed(lambda: pmc.post_moments(X, flg, wb)), 2 * N * (8 * d + 10),
    "includes 2 small D2H + sync")
row("post_moments (mean only, 1 pass)", timed(lambda: pmc.post_moments(X, flg, wb, with_cov=False)), N * (8 * d + 10))
row("post_histogram 1-D, 64 bins", timed(lambda: pmc.post_histogram(X, flg, wb, [0], [64], lim[:2])), N * (8 + 10),
    "algorithmic bytes: one 8-byte column of each 40-byte row (sector traffic is the whole row)")
row("post_histogram 2-D, 64x64 bins", timed(lambda: pmc.post_histogram(X, flg, wb, [0, 1], [64, 64], lim)), N * (16 + 10))
row("post_sigma (gather + CUB radix sort + scan + search), one parameter",
    timed(lambda: pmc.post_sigma(X, flg, wb, 0, mean[0], P.CONF_123_HALF)), N * (8 + 10 + 16),
    "algorithmic bytes = gather only; the 64-bit radix sort moves ~8 x 16 N bytes more")
# CPU side on a bounded sample
n = min(a.cpu_n, N)
hX = X[:n].cpu().numpy(); hf = flg[:n].cpu().numpy(); hw = wb[:n].cpu().numpy()
keep = hf != 0
hX, hw = np.ascontiguousarray(hX[keep]), np.ascontiguousarray(hw[keep]); hw /= hw.sum()
ones = np.ones(len(hX), np.int16)
cpu = {}
kind = "reference (oracle/_ref: exec_helper.c, nhist.c compiled unchanged)" if P.ref() is not None else "port (numpy)"
t = time.perf_counter(); (P.ref_sigma if P.ref() is not None else P.sigma)(hX, hw, ones, 0, mean[0]); cpu["sigma_from_psim"] = time.perf_counter() - t
if P.ref() is not None:
    t = time.perf_counter(); P.ref_histogram(hX, hw, [0, 1], [64, 64], lim); cpu["acc_histogram 2-D"] = time.perf_counter() - t
t = time.perf_counter(); P.moments(hX, hw, None); cpu["moments (numpy)"] = time.perf_counter() - t
print(json.dumps({"tool": "bench_post", "samples": N, "ndim": d, "hbm_peak_GBs": peak, "device": rows,
                  "cpu": {"kind": kind, "cores": 1, "samples": int(len(hX)), "seconds": cpu}}, indent=1))
