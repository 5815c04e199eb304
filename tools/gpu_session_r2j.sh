#!/bin/bash
# GPU session J of round 2: lazy sample delivery (new e2e contract), E-step cache for d < 10 (A/B), small batches
cd "$(dirname "$0")/.."
O=gpurun_out/r2j; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "iteration or lazy or pipelined" > $O/pytest_it.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_it.log
tail -5 $O/pytest_it.log
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
for c in sn sn_bao cmb_bao_sn; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > $O/bench_${c}_rho10.json 2> $O/bench_${c}_rho10.err
  PMCB200_RHO_MIN_DIM=2 timeout 300 python bench.py --config $c --no-cpu-baseline > $O/bench_${c}_rho2.json 2> $O/bench_${c}_rho2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/bench_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']
            print(f.split('/')[-1], 'ms %.3f'%d['ms_per_step'], 'e2e %.3f'%e['ms_per_step'], 'allX %.3f'%e['with_X_every_step']['ms_per_step'], 'fetch %.2f'%e['sample_fetch_ms'])
PY
for n in 10000 50000 200000; do
  echo "N=$n"
  timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1
  PMCB200_SN_WARP_MAX=0 timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1
  PMCB200_SN_WARP_MAX=1000000000 timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1
  PMCB200_SN_WARP_MAX=0 PMCB200_SN_EXACT=1 timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1
done > $O/small_n.txt 2>&1
cat $O/small_n.txt
timeout 200 python bench.py --nsamples 10000 --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_c1_1e4.json 2> $O/bench_c1_1e4.err
cut -c1-300 $O/bench_c1_1e4.json
PMCB200_RHO_MIN_DIM=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "iteration or em_" > $O/pytest_rho2.log 2>&1; echo "pytest rho2 rc=$?" | tee -a $O/pytest_rho2.log
tail -3 $O/pytest_rho2.log
