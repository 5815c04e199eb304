#!/bin/bash
# 8-GPU session S of round 2: where the 8-GPU end-to-end step loses 2.5 ms against the device-timed one (wall-clock phases per rank)
cd "$(dirname "$0")/.."
O=gpurun_out/r2s; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
for mode in lazy noweights nothing; do
  timeout 200 $TR tools/e2e_phases.py --mode $mode > $O/phases_n8_$mode.txt 2> $O/phases_n8_$mode.err; grep "^rank" $O/phases_n8_$mode.txt | head -3
done
timeout 200 python tools/e2e_phases.py --mode lazy > $O/phases_n1_lazy.txt 2> $O/phases_n1.err; grep "^rank" $O/phases_n1_lazy.txt
