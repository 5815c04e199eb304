"""world_size-2 gloo test of the N>1 host logic (no GPU): shard layout, the
all-gather of per-rank statistics blocks and the fixed-order combine.  The
per-rank blocks are produced by the CPU oracle's EM statistics restated in
numpy here (test infrastructure); the product's combine kernel itself is
covered on the GPU by test_sharded_blocks_combine_like_one_rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(N, world, rank):
    per = (N + world - 1) // world
    off = rank * per
    return off, max(0, min(per, N - off))


def local_block(logw, flg):
    ok = flg != 0
    M = logw[ok].max() if ok.any() else -np.inf
    w = np.where(ok, np.exp(logw - M), 0.0) if ok.any() else np.zeros_like(logw)
    return np.array([M, w.sum(), (w ** 2).sum(), (w * np.where(ok, logw - M, 0)).sum(), ok.sum()])


def combine(blocks):
    M = max(b[0] for b in blocks)
    S = S2 = T = nok = 0.0
    for b in blocks:                                  # fixed rank order
        sc = 0.0 if b[0] == -np.inf else np.exp(b[0] - M)
        S += b[1] * sc; S2 += b[2] * sc * sc; nok += b[4]
        if sc > 0:
            T += sc * (b[3] + (b[0] - M) * b[1])
    return M, S, S2, T, nok


def worker(rank, world, port, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                 # same global arrays on every rank
    logw = rng.normal(size=N) * 4 - 300
    flg = (rng.random(N) > 0.05).astype(np.int16)
    off, n = shard(N, world, rank)
    blk = torch.from_numpy(local_block(logw[off:off + n], flg[off:off + n]))
    allb = torch.empty(world * blk.numel(), dtype=torch.float64)
    dist.all_gather_into_tensor(allb, blk)
    allb = allb.view(world, -1)
    res = combine([allb[g].numpy() for g in range(world)])
    q.put((rank, res, (off, n)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allgather_combine():
    world, N = 2, 10001                               # ragged: the last shard is shorter
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, r0, s0), (_, r1, s1) = out
    assert s0 == (0, 5001) and s1 == (5001, 5000)
    assert r0 == r1                                   # bit-identical on both ranks
    rng = np.random.default_rng(123)
    logw = rng.normal(size=N) * 4 - 300
    flg = (rng.random(N) > 0.05).astype(np.int16)
    M, S, S2, T, nok = combine([local_block(logw, flg)])
    assert r0[0] == M and r0[4] == nok
    assert abs(r0[1] - S) < 1e-12 * S and abs(r0[2] - S2) < 1e-12 * S2 and abs(r0[3] - T) < 1e-10 * abs(T)
    # perplexity from the combined block equals the direct definition
    wbar = np.where(flg != 0, np.exp(logw - M), 0) / S
    perp_direct = np.exp(-np.sum(wbar[wbar > 0] * np.log(wbar[wbar > 0]))) / N
    assert abs(np.exp(np.log(r0[1]) - r0[3] / r0[1]) / N - perp_direct) < 1e-12


def test_shard_layout_covers_all_samples():
    for N in (0, 1, 7, 10 ** 7, 10 ** 7 + 3):
        for world in (1, 2, 4, 8):
            segs = [shard(N, world, r) for r in range(world)]
            assert sum(n for _, n in segs) == N
            pos = 0
            for off, n in segs:
                if n:
                    assert off == pos
                    pos += n
