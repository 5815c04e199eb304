#!/bin/bash
# GPU session B of round 2: full parity suite, SN kernel occupancy variants (flat and curved), C4/C5 launch lists with the
# warp-cooperative deep Romberg stages, benches
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
for cfg in sn sn_curved; do
for r in 1 2; do
  for v in default chain mb3 chain_mb3 b128mb5 b128mb6 b192mb4 chain_b192mb4; do
    if [ $v = default ]; then timeout 120 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1
    else PMCB200_LIB=$PWD/variants/$v.so timeout 120 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1; fi
  done
done > $O/ab_$cfg.txt 2>&1
done
for cfg in cmb_bao_sn sn_bao; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$cfg.csv \
    python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_$cfg.log 2>&1
done
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
tail -5 $O/pytest.log; cat $O/ab_sn.txt $O/ab_sn_curved.txt
