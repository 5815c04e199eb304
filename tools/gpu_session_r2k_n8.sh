#!/bin/bash
# 8-GPU session K of round 2: bench with the lazy-sample e2e contract (C2 weak + strong, C5 at 1e8 samples)
cd "$(dirname "$0")/.."
O=gpurun_out/r2k; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8_c2.json 2> $O/bench_n8_c2.err
tail -1 $O/bench_n8_c2.json | cut -c1-300
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --config cmb_bao_sn --nsamples 12500000 > $O/bench_n8_c5.json 2> $O/bench_n8_c5.err
tail -1 $O/bench_n8_c5.json | cut -c1-300
