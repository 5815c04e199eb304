#!/bin/bash
# GPU session L of round 2: TF32 tail of the coefficient contraction -- parity, error against the node-by-node kernel,
# timing against the all-FP64 tensor-core kernel, bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2l; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sn or posterior" > $O/pytest_sn.log 2>&1; echo "pytest sn rc=$?" | tee -a $O/pytest_sn.log
tail -25 $O/pytest_sn.log
for cfg in sn sn_curved sn_bao; do
  timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_t32_$cfg.pt 2>&1 | tail -1
  PMCB200_SN_TAIL32=0 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_f64_$cfg.pt 2>&1 | tail -1
  PMCB200_SN_EXACT=1 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_exact_$cfg.pt 2>&1 | tail -1
  python tools/cmp_lp.py $O/lp_t32_$cfg.pt $O/lp_exact_$cfg.pt
  python tools/cmp_lp.py $O/lp_f64_$cfg.pt $O/lp_exact_$cfg.pt
  rm -f $O/lp_*_$cfg.pt
done > $O/ab_t32.txt 2>&1
cat $O/ab_t32.txt
timeout 400 python bench.py --no-cpu-baseline > $O/bench_sn.json 2> $O/bench_sn.err
cut -c1-300 $O/bench_sn.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma_t32 \
  python tools/time_sn.py --n 4000000 > $O/ncu_t32.log 2>&1
python tools/ncu_summary.py $O/sn_spec_mma_t32.ncu-rep "k_like_sn_spec_mma<0,1,T32> (TF32 tail), M=28, N=4e6" > $O/sn_spec_mma_t32_summary.txt
cat $O/sn_spec_mma_t32_summary.txt | grep -E "gpu__time|pipe_fp64|dmma|issue_active|registers|pipe_tensor"
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -4 $O/pytest.log
