#!/usr/bin/env python
"""Time the SN likelihood kernel alone (CUDA events), for A/B library builds.
usage: PMCB200_LIB=variants/x.so python tools/time_sn.py [--n 4000000]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_config, SEED
from cosmopmc_b200.pmc import PMC
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=4_000_000); ap.add_argument("--config", default="sn"); ap.add_argument("--save", default="")
a = ap.parse_args()
spec, w, m, ch, label = make_config(a.config)
pmc = PMC(0); pmc.set_target(spec); pmc.set_proposal(w, m, chol=ch)
b = pmc.alloc(a.n)
pmc.simulate_mix_mvdens(a.n, SEED, 0, 0, b["X"], b["idx"], b["flg"])
for _ in range(2): pmc.posterior_log_pdf(b["X"])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
pmc.counters()
for _ in range(5): lp, err = pmc.posterior_log_pdf(b["X"])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
c = pmc.counters()
print("%s: %.3f ms for %d samples -> %.3f ns/sample, checksum %.10f, SN spectral %d node-by-node %d per call" % (
    os.environ.get("PMCB200_LIB", "default"), ms, a.n, ms * 1e6 / a.n, lp.sum().item(), c["sn_spec"] // 5, c["sn_exact"] // 5))
if a.save:
    torch.save({"lp": lp.cpu(), "err": err.cpu()}, a.save)
