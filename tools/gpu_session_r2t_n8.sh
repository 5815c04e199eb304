#!/bin/bash
# 8-GPU session T of round 2: the update's result by stores into pinned memory instead of the copy engine -- phases of the
# end-to-end step with and without, then the bench line
cd "$(dirname "$0")/.."
O=gpurun_out/r2t; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 150 $TR tools/e2e_phases.py --mode lazy > $O/phases_n8_lazy_stores.txt 2> $O/p1.err; grep "^rank [01]" $O/phases_n8_lazy_stores.txt
PMCB200_RESULT_MEMCPY=1 timeout 150 $TR tools/e2e_phases.py --mode lazy > $O/phases_n8_lazy_memcpy.txt 2> $O/p2.err; grep "^rank [01]" $O/phases_n8_lazy_memcpy.txt
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8_c2.json 2> $O/bench_n8_c2.err
tail -1 $O/bench_n8_c2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n8 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'allx',d['e2e']['with_X_every_step']['ms_per_step'],'strong',d['strong']['ms_per_step'],d['strong']['e2e']['ms_per_step'])"
