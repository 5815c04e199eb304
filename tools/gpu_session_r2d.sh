#!/bin/bash
# GPU session D of round 2: tensor-core spectral SN kernel -- parity, timing against v1 / exact, benches, ncu
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sn or posterior" > $O/pytest_sn.log 2>&1; echo "pytest sn rc=$?" | tee -a $O/pytest_sn.log
tail -30 $O/pytest_sn.log
for cfg in sn sn_curved sn_bao cmb_bao_sn; do
  timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_mma_$cfg.pt 2>&1 | tail -1
  PMCB200_SN_SPEC_V1=1 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1
  PMCB200_SN_EXACT=1 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_exact_$cfg.pt 2>&1 | tail -1
  python tools/cmp_lp.py $O/lp_mma_$cfg.pt $O/lp_exact_$cfg.pt
  rm -f $O/lp_mma_$cfg.pt $O/lp_exact_$cfg.pt
done > $O/ab_spec.txt 2>&1
cat $O/ab_spec.txt
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
cat $O/bench_sn.json $O/bench_c5.json $O/bench_c4.json | cut -c1-700
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_sn.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_sn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma \
  python tools/time_sn.py --n 4000000 > $O/ncu_sn_spec_mma.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -5 $O/pytest.log
