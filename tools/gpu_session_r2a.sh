#!/bin/bash
# GPU session A of round 2: parity suite, SN kernel occupancy variants, C4/C5 launch lists (lean vs round-1 BAO/CMB kernels)
cd "$(dirname "$0")/.."
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
for r in 1 2; do
  for v in default chain mb3 chain_mb3 b128mb5 b192mb4 chain_b192mb4; do
    if [ $v = default ]; then timeout 120 python tools/time_sn.py --n 10000000 2>&1 | tail -1
    else PMCB200_LIB=$PWD/variants/$v.so timeout 120 python tools/time_sn.py --n 10000000 2>&1 | tail -1; fi
  done
done > $O/ab_sn.txt 2>&1
for v in default chain mb3 chain_mb3 b192mb4; do
  if [ $v = default ]; then timeout 120 python tools/time_sn.py --n 4000000 --config cmb_bao_sn 2>&1 | tail -1
  else PMCB200_LIB=$PWD/variants/$v.so timeout 120 python tools/time_sn.py --n 4000000 --config cmb_bao_sn 2>&1 | tail -1; fi
done > $O/ab_c5_posterior.txt 2>&1
for cfg in cmb_bao_sn sn_bao; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$cfg.csv \
    python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_$cfg.log 2>&1
  PMCB200_LIKE_V1=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_${cfg}_v1.csv \
    python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_${cfg}_v1.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_banana.csv \
  python bench.py --config banana --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_banana.log 2>&1
timeout 300 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python bench.py --config banana --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_cmbdp -s 1 -c 1 -o $O/cmbdp_lean \
  python tools/time_sn.py --n 2000000 --config cmb_bao_sn > $O/ncu_cmbdp.log 2>&1
tail -3 $O/pytest.log; cat $O/ab_sn.txt $O/ab_c5_posterior.txt; cat $O/bench_sn.json | head -c 1500
