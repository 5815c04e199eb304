#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the sharded
iteration (NCCL all-gather of the EM statistics blocks) must reproduce the
single-rank iteration over the same global samples, and every rank must hold
the bit-identical updated proposal."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bench import make_config, SEED
from cosmopmc_b200.pmc import PMC, run_iteration_distributed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(os.environ.get("CHECK_N", "400000"))
for cfg in ("sn", "banana"):
    spec, w, m, ch, label = make_config(cfg)
    pmc = PMC(local); pmc.set_target(spec); pmc.set_proposal(w, m, chol=ch)
    blen = pmc.stat_block_len()
    block = torch.zeros(blen, dtype=torch.float64, device="cuda")
    allb = torch.zeros(world * blen, dtype=torch.float64, device="cuda")
    stats = []
    for it in range(3):
        stats.append(run_iteration_distributed(pmc, N, SEED, it, 1.0, block, allb, None, rank, world))
    prop3 = pmc.get_proposal()
    # host-buffer shard path: overlapped D2H must deliver the same arrays as the device path
    per = (N + world - 1) // world
    off = rank * per
    n_loc = max(0, min(per, N - off))
    hX = torch.empty((max(n_loc, 1), len(m[0])), dtype=torch.float64).pin_memory()
    hidx = torch.empty(max(n_loc, 1), dtype=torch.int32).pin_memory()
    hflg = torch.empty(max(n_loc, 1), dtype=torch.int16).pin_memory()
    hw = torch.empty(max(n_loc, 1), dtype=torch.float64).pin_memory()
    ref = PMC(local); ref.set_target(spec); ref.set_proposal(*pmc.get_proposal()[:2], chol=pmc.get_proposal()[2])
    bufs = ref.alloc(n_loc)
    bref = torch.zeros(blen, dtype=torch.float64, device="cuda")
    ref.iteration_local(n_loc, SEED, 7, off, 1.0, bref, bufs)
    pmc.iteration_shard_host(n_loc, SEED, 7, off, 1.0, block, hX, hidx, hflg)
    dist.all_gather_into_tensor(allb, block)
    st_h = pmc.update_prop_rb(world, allb, N)
    pmc.shard_weights_host(n_loc, hw)
    assert torch.equal(hX[:n_loc], bufs["X"][:n_loc].cpu()) and torch.equal(hidx[:n_loc], bufs["idx"][:n_loc].cpu())
    assert torch.equal(hflg[:n_loc], bufs["flg"][:n_loc].cpu())
    assert torch.equal(block, bref)
    tot = torch.tensor([hw[:n_loc].sum().item()], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    assert abs(tot.item() - 1.0) < 1e-10, tot.item()
    prop = pmc.get_proposal()
    flat = torch.from_numpy(np.concatenate([p.ravel() for p in prop])).cuda()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for g in gathered:
        assert torch.equal(g, gathered[0]), "proposal differs between ranks"
    if rank == 0:
        one = PMC(local); one.set_target(spec); one.set_proposal(w, m, chol=ch)
        b1 = torch.zeros(blen, dtype=torch.float64, device="cuda")
        for it in range(3):
            one.iteration_local(N, SEED, it, 0, 1.0, b1)
            s1 = one.update_prop_rb(1, b1, N)
            for k in ("nok", "nok_box", "ndead"):
                assert s1[k] == stats[it][k], (k, s1[k], stats[it][k])
            for k in ("maxW", "logSum", "perplexity", "ess", "enc"):
                assert abs(s1[k] - stats[it][k]) <= 1e-11 * abs(s1[k]), (k, s1[k], stats[it][k])
        for a, b in zip(one.get_proposal(), prop3):
            assert np.allclose(a, b, rtol=1e-9, atol=1e-13)
        print("multirank ok: %s world=%d N=%d perplexity=%.6f" % (cfg, world, N, stats[-1]["perplexity"]), flush=True)
dist.barrier()
dist.destroy_process_group()
