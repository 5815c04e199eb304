#!/usr/bin/env python
"""Single-process multi-GPU scaling of the C-ABI call pmcb200_iteration_host_multi:
one context per GPU of this process, page-locked host buffers, statistics blocks exchanged by
peer copies.  Complements bench.py (one PROCESS per GPU over torch.distributed / NCCL).

  python tools/bench_multi_ctx.py [--n 10000000] [--steps 5] [--warmup 2] [--gpus 1,2,4,8]

Prints one JSON line per GPU count and scaling mode; timing = wall clock around the blocking
call (the call returns after every device finished and the host arrays are complete)."""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_config, SEED
from cosmopmc_b200.pmc import PMC, iteration_host_multi

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--gpus", default="")
ap.add_argument("--config", default="sn")
a = ap.parse_args()
ndev = torch.cuda.device_count()
counts = [int(g) for g in a.gpus.split(",")] if a.gpus else [g for g in (1, 2, 4, 8) if g <= ndev]
spec, w, m, ch, label = make_config(a.config)
d = m.shape[1]
ref = None
for G in counts:
    pmcs = [PMC(g % ndev, use_torch_stream=False) for g in range(G)]
    for p in pmcs:
        p.set_target(spec)
    for mode in ("strong", "weak"):
        N = a.n if mode == "strong" else a.n * G
        hX = torch.empty((N, d), dtype=torch.float64).pin_memory()
        hidx = torch.empty(N, dtype=torch.int32).pin_memory()
        hflg = torch.empty(N, dtype=torch.int16).pin_memory()
        hw = torch.empty(N, dtype=torch.float64).pin_memory()
        ts = []
        for it in range(a.warmup + a.steps):
            for p in pmcs:
                p.set_proposal(w, m, chol=ch)
            for g in range(G):
                torch.cuda.synchronize(g % ndev)
            t = time.perf_counter()
            st = iteration_host_multi(pmcs, N, SEED, it, 1.0, hX, hidx, hflg, hw)
            ts.append(time.perf_counter() - t)
        ms = 1e3 * float(np.mean(ts[a.warmup:]))
        if G == counts[0] and mode == "strong":
            ref = (hX[:1000].clone(), hw[:1000].clone(), st["perplexity"])
        elif mode == "strong":   # same seed, same N: identical samples whatever the GPU count
            assert torch.equal(hX[:1000], ref[0]) and abs(st["perplexity"] - ref[2]) < 1e-10 * ref[2]
        print(json.dumps({"tool": "bench_multi_ctx", "config": label, "gpus": G, "devices": ndev, "scaling": mode,
                          "samples": N, "ms_per_iteration": ms, "samples_per_s": N / (ms * 1e-3),
                          "perplexity": st["perplexity"], "nok": st["nok"],
                          "d2h_bytes": N * (8 * d + 4 + 2 + 8)}), flush=True)
        del hX, hidx, hflg, hw
    for p in pmcs:
        p.close()
