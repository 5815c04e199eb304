#!/bin/bash
# 8-GPU session of round 2 (one box, NCCL over NVLink): two-device test, rank-count independence, weak + strong bench
# (C2), C5 at 1e8 samples over 8 GPUs, single-process multi-context scaling
cd "$(dirname "$0")/.."
O=gpurun_out/r2i; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_devices" > $O/pytest_two.log 2>&1; tail -2 $O/pytest_two.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multirank_check.py > $O/multirank_check.txt 2>&1; tail -4 $O/multirank_check.txt
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8_c2.json 2> $O/bench_n8_c2.err
tail -1 $O/bench_n8_c2.json | cut -c1-1200
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --config cmb_bao_sn --nsamples 12500000 > $O/bench_n8_c5.json 2> $O/bench_n8_c5.err
tail -1 $O/bench_n8_c5.json | cut -c1-600
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR4 bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_n4_c2.json 2> $O/bench_n4_c2.err
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR2 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2_c2.json 2> $O/bench_n2_c2.err
timeout 300 python tools/bench_multi_ctx.py --n 10000000 --gpus 1,2,4,8 > $O/multi_ctx.jsonl 2> $O/multi_ctx.err
cat $O/multi_ctx.jsonl | cut -c1-300
