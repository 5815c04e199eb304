#!/usr/bin/env python
"""bench.py -- PMC samples/sec for one full iteration (sample + likelihood +
importance weights + EM update) of the SN Ia configuration (BASELINE.json
configs[1]: SNLS/Union SN Ia likelihood, 5 parameters, 10-component proposal,
10^7 samples per iteration), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pmc_samples_per_sec_full_iteration"
UNIT = "samples/s"
SEED = 20090903

# FLOPs of the SN likelihood (convention SURVEY.md 8d: + - * = 1, FMA = 2, / sqrt exp log pow = 1 each); derivation in
# DESIGN.md section 6.  Node-by-node path (the reference algorithm as restated in oracle/pmc_oracle.c):
FLOP_PER_EVAL = 18.0       # int_for_w: a^4 E^2(a) incl. 1 pow + 1 exp (14), sqrt, 1/x, sum += (4)
FLOP_PER_ZSTEP = 86.0      # per (sample, redshift): 5 trapzd combines (22), NR polint K=5 (54), test (4), D_L + modulus (6)
FLOP_PER_SN = 29.0         # per (sample, supernova): mu_obs 7, sigma^2 18, chi^2 term 4
# Spectral path (k_like_sn_spec): M integrand evaluations, the folded DCT (M^2/2 FMA + M adds), the two certificates
# (~3 M), then per redshift an M-term dot product (M FMA) + D_L + modulus (8)
SPEC_M = 28
FLOP_SPEC_SAMPLE = SPEC_M * FLOP_PER_EVAL + SPEC_M * SPEC_M + 4 * SPEC_M
FLOP_SPEC_ZSTEP = 2.0 * SPEC_M + 8.0
# ncu evidence for the dominant kernel (one --set full capture, N = 4e6): FP64 pipe = DMMA sub-pipe 53.8 % + vector 23.1 %;
# DRAM bytes (read + write) per sample
NCU_SN = {"fp64_pipe_active_pct": 76.8, "dram_bytes_per_sample": 48.4, "source": "profiles/r02/sn_spec_mma_v6_summary.txt"}
NCU_SN_EXACT = {"fp64_pipe_active_pct": 64.6, "dram_bytes_per_sample": 44.1, "source": "profiles/sn_r01_v8_summary.txt"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nsamples", type=int, default=10_000_000, help="samples per GPU per iteration")
    ap.add_argument("--config", default="sn", choices=["sn", "sn_curved", "banana", "sn_bao", "cmb_bao_sn"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    return ap.parse_args()


def make_config(name):
    from cosmopmc_b200 import targets as T
    if name == "sn":
        spec = T.target_sn_demo()
        w, m, cov = T.proposal_sn(10)
        label = "SN Ia (Union 307 SNe, flat wCDM): Omega_m w0 M alpha beta, K=10 Gaussian proposal"
    elif name == "sn_curved":       # C2's curved variant (BASELINE.json text: Omega_m Omega_de M alpha beta); tools/ only
        spec = T.target_sn_curved()
        w, m, cov = T.proposal_generic(spec, 10, 4, [0.3, 0.75, 19.33, 1.4, -2.4], [0.06, 0.12, 0.03, 0.1, 0.1])
        label = "SN Ia (Union 307 SNe, curved LCDM): Omega_m Omega_de M alpha beta, K=10"
    elif name == "banana":
        spec = T.target_banana(20)
        w, m, cov = T.proposal_banana(10, 20)
        label = "20-D banana (Wraith et al. 2009), K=10"
    elif name == "sn_bao":
        spec = T.target_sn_bao_w0wa()
        # w0 + w1 stays well below 1/3: beyond it dark energy dominates the early universe and the
        # sound-horizon integral stops converging (up to 2^19 evaluations per sample, as in the reference)
        w, m, cov = T.proposal_generic(spec, 10, 4, [0.28, 0.72, -1.0, 0.0, 19.31, 1.4, -2.4],
                                       [0.04, 0.06, 0.15, 0.2, 0.03, 0.1, 0.1])
        label = "SN Ia + BAO d_z, w0-wa, 7 parameters, K=10"
    else:
        spec = T.target_cmb_bao_sn()
        w, m, cov = T.proposal_generic(spec, 30, 4, [0.045, 0.27, 0.73, 0.71, -1.0, 19.31, 1.4, -2.4],
                                       [0.003, 0.02, 0.02, 0.02, 0.1, 0.03, 0.1, 0.1])
        label = "WMAP7 distance priors + BAO + SN Ia, 8 parameters, K=30"
    ch = np.stack([np.linalg.cholesky(c) for c in cov])
    return spec, w, m, ch, label


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region: NVML (nvidia_ml_py, ~1 ms per sample)
    when it loads, else the nvidia-smi query of the profiling recipe (~100 ms per sample)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.bits = [getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                         getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                         getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                         getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        return [str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in self.bits]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self.rows.append(self._sample_nvml())
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows for i in range(4)
                          if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


def cpu_iteration_rate(spec, w, m, ch, seconds, nthreads, mode_b=False):
    """Times the CPU oracle (kind 'port': restatement of pmclib+nicaea, the
    reference binary cannot be built here) on a bounded sample of the workload."""
    from oracle import oracle_lib as O
    O.build()
    n0 = 2000
    t = time.perf_counter()
    O.iteration(spec, n0, SEED, 0, 1.0, w, m, ch, nthreads=nthreads, mode_b=mode_b)
    dt = max(time.perf_counter() - t, 1e-3)
    n = int(min(max(n0, n0 * seconds / dt), 2_000_000))
    t = time.perf_counter()
    O.iteration(spec, n, SEED, 0, 1.0, w, m, ch, nthreads=nthreads, mode_b=mode_b)
    dt = time.perf_counter() - t
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, w, m, ch, label = make_config(args.config)
    cores = os.cpu_count() or 1
    from oracle import oracle_lib as O
    O.build()
    # bounded sample per step, sized so the whole run ends within a few minutes
    rate, n_cal, _ = cpu_iteration_rate(spec, w, m, ch, 2.0, cores)
    per_step = int(min(max(2000, rate * min(20.0, 150.0 / max(1, args.steps + args.warmup))), 2_000_000))
    for i in range(args.warmup):
        O.iteration(spec, per_step, SEED, i, 1.0, w, m, ch, nthreads=cores)
    t = time.perf_counter()
    for i in range(args.steps):
        O.iteration(spec, per_step, SEED, args.warmup + i, 1.0, w, m, ch, nthreads=cores)
    dt = time.perf_counter() - t
    val = per_step * args.steps / dt
    tb = time.perf_counter()
    O.iteration(spec, per_step, SEED, 99, 1.0, w, m, ch, nthreads=cores, mode_b=True)
    val_b = per_step / (time.perf_counter() - tb)
    sample = "%d samples/step of the %s workload (full iteration: sample+likelihood+weights+EM)" % (per_step, args.config)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": label, "samples_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "mode": "A: reference-shaped (serial sampling / normalisation / EM as on the reference's rank 0, "
                                     "weights over all cores like its MPI scatter)",
                             "mode_b": {"value": val_b, "unit": UNIT, "cores": cores,
                                        "mode": "B: every stage parallel over samples (OpenMP); one step of the same sample"}},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU oracle (C restatement of pmclib+nicaea, OpenMP over samples in the weight stage); "
                    "the reference binary cannot be built in this image (pmclib/nicaea/GSL/MPI absent)"}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cosmopmc_b200.pmc import PMC

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    spec, w, m, ch, label = make_config(args.config)
    d, K = len(m[0]), len(w)
    n_loc = args.nsamples if args.scaling == "weak" else (args.nsamples + world - 1) // world
    n_glob = n_loc * world
    off = rank * n_loc

    pmc = PMC(local)
    pmc.set_target(spec)
    pmc.set_proposal(w, m, chol=ch)
    bufs = pmc.alloc(n_loc)
    blen = pmc.stat_block_len()
    block = torch.zeros(blen, dtype=torch.float64, device="cuda")
    allb = torch.zeros((world, blen), dtype=torch.float64, device="cuda")
    # pinned host buffers for the end-to-end path: two sets, the copies of one iteration drain while the next computes
    hsets = [(torch.empty((n_loc, d), dtype=torch.float64).pin_memory(), torch.empty(n_loc, dtype=torch.int32).pin_memory(),
              torch.empty(n_loc, dtype=torch.int16).pin_memory(), torch.empty(n_loc, dtype=torch.float64).pin_memory())
             for _ in range(2)]
    w_pin = torch.from_numpy(w.copy()).pin_memory()
    m_pin = torch.from_numpy(m.copy()).pin_memory()
    ch_pin = torch.from_numpy(ch.copy()).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_steps(nl):
        """step functions for nl samples on this rank (global nl * world, this rank's offset rank * nl)"""
        ng, of = nl * world, rank * nl

        def step_device(it):
            """hot path, inputs resident in HBM: the proposal is re-installed every
            step so that every step does identical work"""
            pmc.set_proposal(w, m, chol=ch)
            pmc.iteration_local(nl, SEED, it, of, 1.0, block, bufs)
            if world > 1:
                dist.all_gather_into_tensor(allb, block)
                return pmc.update_prop_rb(world, allb, ng)
            return pmc.update_prop_rb(1, block, ng)

        def step_e2e(it):
            """the reference-facing call: host proposal in; statistics, updated proposal, flags and normalised weights
            out to pinned host memory EVERY iteration.  The sample array X and the component indices stay in HBM (NULL
            pointers; pmcb200_samples_host_begin fetches it on request): the reference touches psim->X on the host only to
            dump it (cosmo_pmc.c:392) and to post-process the final sample, which the device post-processing does in place.
            Pipelined delivery (pmcb200_iteration_host_begin / pmcb200_host_wait): the call returns when the update is
            done; this step then waits for the PREVIOUS iteration's host arrays, whose copies drained while this
            iteration's kernels ran.  Every array of every timed iteration is complete on the host before the clock stops."""
            _, _, hflg, hw = hsets[it % 2]
            pmc.set_proposal(w_pin.numpy(), m_pin.numpy(), chol=ch_pin.numpy())
            if world == 1:
                st = pmc.iteration_host_begin(nl, SEED, it, 1.0, None, None, hflg[:nl], hw[:nl])
            else:
                pmc.iteration_shard_host(nl, SEED, it, of, 1.0, block, None, None, hflg[:nl])
                dist.all_gather_into_tensor(allb, block)
                st = pmc.update_prop_rb(world, allb, ng)
                pmc.shard_weights_host_begin(nl, hw[:nl])
            pmc.host_wait(1)
            return st

        def step_e2e_allx(it):
            """the same with the sample array delivered to the host every iteration as well (what a driver that dumps
            every iteration's pmcsim needs): 8 d more bytes per sample"""
            hX, hidx, hflg, hw = hsets[it % 2]
            pmc.set_proposal(w_pin.numpy(), m_pin.numpy(), chol=ch_pin.numpy())
            if world == 1:
                st = pmc.iteration_host_begin(nl, SEED, it, 1.0, hX[:nl], hidx[:nl], hflg[:nl], hw[:nl])
            else:
                pmc.iteration_shard_host(nl, SEED, it, of, 1.0, block, hX[:nl], hidx[:nl], hflg[:nl])
                dist.all_gather_into_tensor(allb, block)
                st = pmc.update_prop_rb(world, allb, ng)
                pmc.shard_weights_host_begin(nl, hw[:nl])
            pmc.host_wait(1)
            return st
        return step_device, step_e2e, step_e2e_allx

    def timed(fn, steps, warmup, sampler=None, drain=False):
        for i in range(warmup):
            fn(i)
        if drain:
            pmc.host_wait(0)
        barrier()
        pmc.counters()
        l0 = pmc.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        e0.record()
        st = None
        for i in range(steps):
            st = fn(warmup + i)
        if drain:
            pmc.host_wait(0)          # the last iteration's host arrays, inside the timed region
        e1.record()
        barrier()
        if sampler:
            sampler.stop_flag = True
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), st, pmc.launch_count() - l0, pmc.counters()

    sampler = ClockSampler(local) if rank == 0 else None
    step_device, step_e2e, step_e2e_allx = make_steps(n_loc)
    ms, st, launches, cnt = timed(step_device, args.steps, args.warmup, sampler)
    ms_e2e, st_e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup - 1), drain=True)
    ms_allx, _, _, _ = timed(step_e2e_allx, args.steps, max(1, args.warmup - 1), drain=True)
    # one on-request fetch of the sample array (the once-per-run cost of a final pmcsim dump), timed alone
    barrier()
    t0 = time.perf_counter()
    pmc.samples_host_begin(n_loc, hsets[0][0][:n_loc], hsets[0][1][:n_loc])
    pmc.host_wait(0)
    barrier()
    fetch_ms = 1e3 * (time.perf_counter() - t0)
    value = n_glob * args.steps / (ms * 1e-3)
    e2e = n_glob * args.steps / (ms_e2e * 1e-3)

    # strong scaling beside the weak run: the SAME global sample count as one GPU's (BASELINE.json's target is
    # "10^7 samples in < 100 ms on 8 GPUs"), sharded over the ranks, same steps
    strong = None
    if args.scaling == "weak" and world > 1:
        nl_s = (args.nsamples + world - 1) // world
        sd, se, _ = make_steps(nl_s)
        ms_s, _, _, _ = timed(sd, args.steps, args.warmup)
        ms_se, _, _, _ = timed(se, args.steps, max(1, args.warmup - 1), drain=True)
        strong = {"n_gpus": world, "samples_global": nl_s * world, "samples_per_gpu": nl_s, "ms_per_step": ms_s / args.steps,
                  "value": nl_s * world * args.steps / (ms_s * 1e-3), "unit": UNIT,
                  "e2e": {"value": nl_s * world * args.steps / (ms_se * 1e-3), "ms_per_step": ms_se / args.steps}}

    elif args.scaling == "weak":
        strong = {"n_gpus": 1, "samples_global": n_glob, "samples_per_gpu": n_loc, "ms_per_step": ms / args.steps, "value": value,
                  "unit": UNIT, "e2e": {"value": e2e, "ms_per_step": ms_e2e / args.steps}}

    # dominant kernel alone (SN likelihood), CUDA events on the launching stream
    roof = None
    if args.config in ("sn", "sn_curved", "sn_bao", "cmb_bao_sn"):
        pmc.set_proposal(w, m, chol=ch)
        pmc.simulate_mix_mvdens(n_loc, SEED, 0, off, bufs["X"], bufs["idx"], bufs["flg"])
        for _ in range(2):
            pmc.posterior_log_pdf(bufs["X"])
        torch.cuda.synchronize()
        pmc.counters()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            pmc.posterior_log_pdf(bufs["X"])
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        c = pmc.counters()
        sn_like = [spec.t.like[i] for i in range(spec.t.ndata) if spec.t.like[i].kind == 3][0]      # PMCB200_LIKE_SNIa
        n_sn, n_z = sn_like.sn_n, c["sn_zsteps"] // max(1, c["sn_spec"] + c["sn_exact"])
        spectral = c["sn_spec"] > 0
        # the flops of the algorithm the kernels run: spectral samples by the spectral count, the rest node by node
        ev_exact = c["sn_evals"] - SPEC_M * c["sn_spec"]
        zs_exact = c["sn_zsteps"] - n_z * c["sn_spec"]
        flops = (c["sn_spec"] * (FLOP_SPEC_SAMPLE + n_z * FLOP_SPEC_ZSTEP) + ev_exact * FLOP_PER_EVAL + zs_exact * FLOP_PER_ZSTEP
                 + (c["sn_spec"] + c["sn_exact"]) * n_sn * FLOP_PER_SN) / reps
        # BAO / CMB distance-prior kernels of the joint configurations: their Romberg integrals by the in-kernel counters
        flops_gen = (c["gen_evals"] * FLOP_PER_EVAL + c["gen_integrals"] * FLOP_PER_ZSTEP) / reps
        flops += flops_gen
        # what the reference's node-by-node algorithm would have spent on the same samples (17 evaluations per redshift
        # when stage 5 converges, which the spectral kernel certifies for every sample it keeps)
        flops_ref = (c["sn_spec"] * n_z * (17 * FLOP_PER_EVAL + FLOP_PER_ZSTEP) + ev_exact * FLOP_PER_EVAL + zs_exact * FLOP_PER_ZSTEP
                     + (c["sn_spec"] + c["sn_exact"]) * n_sn * FLOP_PER_SN) / reps + flops_gen
        peak = pmc.fp64_peak_tflops()
        ach = flops / (k_ms * 1e-3) * 1e-12
        ncu = NCU_SN if spectral else NCU_SN_EXACT
        kern = "k_like_sn_spec_mma (+ k_like_sn_warp_list for the samples it hands over)" if spectral else "k_like_sn"
        if args.config == "sn_bao":
            kern = "likelihood stage: " + kern + " + k_like_bao"
        elif args.config == "cmb_bao_sn":
            kern = "likelihood stage: k_like_cmbdp + " + kern + " + k_like_bao"
        roof = {"kernel": kern,
                "bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak if peak else None,
                "traffic": ncu["dram_bytes_per_sample"] * n_loc if (ncu["dram_bytes_per_sample"] and args.config == "sn") else None,
                "traffic_source": "ncu dram__bytes_read+write per sample (%s) x samples per launch; algorithmic = %d B/sample" % (ncu["source"], 8 * d + 12),
                "fp64_pipe_active_pct_ncu": ncu["fp64_pipe_active_pct"] if args.config == "sn" else None,
                "kernel_ms": k_ms, "flop_per_launch": flops,
                "flop_convention": "flops of the algorithm the kernel runs: spectral samples %g + %d x %g + %d x %g per sample; "
                                   "node-by-node samples 18 per evaluation + 86 per redshift + 29 per SN" % (FLOP_SPEC_SAMPLE, n_z, FLOP_SPEC_ZSTEP, n_sn, FLOP_PER_SN),
                "samples_spectral": c["sn_spec"] / reps, "samples_node_by_node": c["sn_exact"] / reps,
                "evals_per_sample": c["sn_evals"] / reps / n_loc,
                "generic_integrals": {"evals_per_sample": c["gen_evals"] / reps / n_loc, "integrals_per_sample": c["gen_integrals"] / reps / n_loc,
                                      "flop_per_launch": flops_gen,
                                      "convention": "BAO / CMB distance-prior Romberg integrals: 18 per evaluation + 86 per integral"},
                "reference_algorithm": {"flop_per_launch": flops_ref, "equivalent_TFLOPs": flops_ref / (k_ms * 1e-3) * 1e-12,
                                        "note": "work the reference's 17-node Romberg per redshift would need for the same result; "
                                                "a speed-up measure, not a roofline fraction"},
                "peak_source": "measured live: DFMA-only kernel (pmcb200_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "hbm_algorithmic_GBs": n_loc * (8 * d + 8 + 4) / (k_ms * 1e-3) * 1e-9}

    if rank == 0:
        clocks = sampler.summary() if sampler else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": label, "samples_per_gpu": n_loc, "samples_global": n_glob,
                           "ndim": d, "ncomp": K, "seed": SEED,
                           "l2": "inputs larger than L2 (sample array %.0f MB per GPU)" % (n_loc * d * 8 / 1e6),
                           "parallelism": "samples sharded over %d GPU(s); one NCCL all-gather of %d doubles per iteration" % (world, blen)},
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "api": ("pmcb200_iteration_host_begin + pmcb200_host_wait" if world == 1 else
                                "pmcb200_iteration_shard_host + NCCL all-gather + pmcb200_em_finish + pmcb200_shard_weights_host_begin "
                                "+ pmcb200_host_wait") +
                               " (pipelined host delivery; proposal in, statistics + updated proposal + flags + normalised weights out "
                               "every iteration, all complete inside the timed region; the sample array and the component indices stay in HBM, "
                               "pmcb200_samples_host_begin fetches it on request)",
                        "h2d_bytes_per_step": int(w_pin.numel() + m_pin.numel() + ch_pin.numel()) * 8,
                        "d2h_bytes_per_step": int(n_loc * (2 + 8) + 8 * (16 + K * (1 + d + d * d))),
                        "with_X_every_step": {"value": n_glob * args.steps / (ms_allx * 1e-3), "ms_per_step": ms_allx / args.steps,
                                              "d2h_bytes_per_step": int(n_loc * (8 * d + 4 + 2 + 8) + 8 * (16 + K * (1 + d + d * d))),
                                              "note": "the sample array delivered every iteration as well (a driver that dumps "
                                                      "every iteration's pmcsim); bound by the host's aggregate D2H bandwidth on 8 GPUs"},
                        "sample_fetch_ms": fetch_ms,
                        "sample_fetch_note": "one pmcb200_samples_host_begin + wait of this rank's %d x %d doubles + indices, all ranks "
                                             "at once, wall clock (the once-per-run cost of a final dump)" % (n_loc, d)},
                "strong": strong,
                "gpu_launches": launches, "clocks": clocks,
                "counters_timed_region": cnt,
                "stats": {k: st[k] for k in ("perplexity", "ess", "nok", "enc", "ndead")} if st else None}
        if roof:
            line["roofline"] = roof
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            rate, n, dt = cpu_iteration_rate(spec, w, m, ch, args.cpu_seconds, cores)
            rate_b, n_b, dt_b = cpu_iteration_rate(spec, w, m, ch, args.cpu_seconds, cores, mode_b=True)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "mode": "A: reference-shaped (sampling, normalisation and EM serial as on the reference's rank 0; "
                                            "weights over all cores like its MPI scatter)",
                                    "sample": "%d samples, one full iteration of the same workload, %.1f s" % (n, dt),
                                    "mode_b": {"value": rate_b, "unit": UNIT, "cores": cores,
                                               "mode": "B: every stage parallel over samples (OpenMP)",
                                               "sample": "%d samples, %.1f s" % (n_b, dt_b)}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
