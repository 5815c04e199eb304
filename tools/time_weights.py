#!/usr/bin/env python
"""Time the importance-weight kernel (log q fused) and the EM statistics kernel alone, for A/B runs.
usage: [PMCB200_ESTEP=k] python tools/time_weights.py [--config banana] [--n 4000000]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_config, SEED
from cosmopmc_b200.pmc import PMC
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=4_000_000); ap.add_argument("--config", default="banana")
a = ap.parse_args()
spec, w, m, ch, label = make_config(a.config)
pmc = PMC(0); pmc.set_target(spec); pmc.set_proposal(w, m, chol=ch)
b = pmc.alloc(a.n)
blk = torch.zeros(pmc.stat_block_len(), dtype=torch.float64, device="cuda")
pmc.simulate_mix_mvdens(a.n, SEED, 0, 0, b["X"], b["idx"], b["flg"])
pmc.iteration_local(a.n, SEED, 0, 0, 1.0, blk, b)
flg0 = b["flg"].clone()
def timed(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def wfn():
    b["flg"].copy_(flg0)
    pmc.get_importance_weight(b["X"], b["flg"], b["logw"], 1.0)
tw = timed(wfn)
tc = timed(lambda: b["flg"].copy_(flg0))
te = timed(lambda: pmc.em_local(b["X"], b["idx"], b["flg"], b["logw"], blk, a.n))
print("ESTEP=%s %s: weights %.3f ms, em_local %.3f ms for %d samples; checksum %.12e" % (
    os.environ.get("PMCB200_ESTEP", "-"), a.config, tw - tc, te, a.n, b["logw"].sum().item()))
