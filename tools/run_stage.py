#!/usr/bin/env python
"""Run individual stages of the PMC iteration a few times (for ncu captures).
usage: python tools/run_stage.py [--config sn] [--n 2000000] [--reps 3] [--stage posterior|iteration]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_config, SEED
from cosmopmc_b200.pmc import PMC

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="sn")
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--stage", default="posterior")
a = ap.parse_args()
spec, w, m, ch, label = make_config(a.config)
pmc = PMC(0); pmc.set_target(spec); pmc.set_proposal(w, m, chol=ch)
b = pmc.alloc(a.n)
blk = torch.zeros(pmc.stat_block_len(), dtype=torch.float64, device="cuda")
pmc.simulate_mix_mvdens(a.n, SEED, 0, 0, b["X"], b["idx"], b["flg"])
for r in range(a.reps):
    if a.stage == "posterior":
        pmc.posterior_log_pdf(b["X"])
    else:
        pmc.set_proposal(w, m, chol=ch)
        pmc.iteration_local(a.n, SEED, r, 0, 1.0, blk, b)
        pmc.update_prop_rb(1, blk, a.n)
torch.cuda.synchronize()
print("done", label)
