#!/bin/bash
# GPU session G of round 2: full parity suite at HEAD (branch-free f_K in the tensor-core SN kernel, table-based
# Box-Muller in the sampler), Chebyshev order variants, benches of all configurations
cd "$(dirname "$0")/.."
O=gpurun_out/r2g; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -15 $O/pytest.log
for cfg in sn sn_curved cmb_bao_sn; do
  timeout 200 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1
  for v in m28 m24; do PMCB200_LIB=$PWD/variants/$v.so timeout 200 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1; done
done > $O/ab_m.txt 2>&1
cat $O/ab_m.txt
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config banana --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
cat $O/bench_sn.json $O/bench_c3.json $O/bench_c5.json $O/bench_c4.json | cut -c1-420
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c3.csv \
  python bench.py --config banana --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_c3.log 2>&1
grep -E "k_simulate|k_weights|k_em_stats" $O/launches_c3.csv | awk -F'","' '{print $5, $NF}' | head -12
