#!/usr/bin/env python
"""Wall-clock phases of the multi-rank end-to-end step (bench.py step_e2e), per rank: where do the milliseconds go that the
8-GPU e2e line loses against the device-timed one?  Run under torchrun (or alone: world 1).
usage: torchrun --nproc-per-node 8 tools/e2e_phases.py [--mode lazy|noweights|noflags|nothing]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bench import make_config, SEED
from cosmopmc_b200.pmc import PMC

ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="lazy"); ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec, w, m, ch, label = make_config("sn")
d, K = len(m[0]), len(w)
nl, ng, of = a.n, a.n * world, rank * a.n
pmc = PMC(local); pmc.set_target(spec); pmc.set_proposal(w, m, chol=ch)
blen = pmc.stat_block_len()
block = torch.zeros(blen, dtype=torch.float64, device="cuda")
allb = torch.zeros((world, blen), dtype=torch.float64, device="cuda")
hs = [(torch.empty(nl, dtype=torch.int16).pin_memory(), torch.empty(nl, dtype=torch.float64).pin_memory()) for _ in range(2)]
names = ["set_proposal", "shard_host(enqueue)", "all_gather(enqueue)", "update_prop_rb(sync)", "weights_begin", "host_wait(1)"]
acc = np.zeros(len(names)); tot = 0.0; gacc = np.zeros(3)


def step(it, rec):
    global tot
    hflg, hw = hs[it % 2]
    if a.mode in ("noflags", "nothing"):
        hflg = None
    t = [time.perf_counter()]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    pmc.set_proposal(w, m, chol=ch); t.append(time.perf_counter())
    pmc.iteration_shard_host(nl, SEED, it, of, 1.0, block, None, None, hflg); t.append(time.perf_counter())
    ev[1].record()
    if world > 1:
        dist.all_gather_into_tensor(allb, block)
    ev[2].record()
    t.append(time.perf_counter())
    pmc.update_prop_rb(world, allb if world > 1 else block, ng); t.append(time.perf_counter())
    ev[3].record()
    if a.mode in ("lazy", "noflags"):
        pmc.shard_weights_host_begin(nl, hw)
    t.append(time.perf_counter())
    pmc.host_wait(1); t.append(time.perf_counter())
    if rec:
        acc[:] += np.diff(t); tot += t[-1] - t[0]
        torch.cuda.synchronize()
        gacc[:] += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]


for i in range(3):
    step(i, False)
pmc.host_wait(0)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(a.steps):
    step(3 + i, True)
pmc.host_wait(0)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / a.steps
line = "rank %d mode %s: %.3f ms/step wall | " % (rank, a.mode, 1e3 * wall) + "  ".join("%s %.3f" % (n, 1e3 * v / a.steps) for n, v in zip(names, acc)) + \
    " | device: shard kernels %.3f  all-gather %.3f  em_finish+result %.3f" % tuple(gacc / a.steps)
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        print("\n".join(out), flush=True)
    dist.destroy_process_group()
else:
    print(line, flush=True)
