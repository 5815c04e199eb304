#!/bin/bash
# 8-GPU session R of round 2 (HEAD kernels): two-device test, multi-rank parity check, C2 weak + strong, C5 at 1e8 samples
cd "$(dirname "$0")/.."
O=gpurun_out/r2r; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_devices or multi_contexts or sharded_blocks" > $O/pytest_multi.log 2>&1; tail -2 $O/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 300 $TR tools/multirank_check.py > $O/multirank_check.txt 2>&1; tail -4 $O/multirank_check.txt
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8_c2.json 2> $O/bench_n8_c2.err
tail -1 $O/bench_n8_c2.json | cut -c1-300
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --config cmb_bao_sn --nsamples 12500000 > $O/bench_n8_c5.json 2> $O/bench_n8_c5.err
tail -1 $O/bench_n8_c5.json | cut -c1-300
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522"
timeout 300 $TR2 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2_c2.json 2> $O/bench_n2_c2.err
tail -1 $O/bench_n2_c2.json | cut -c1-200
