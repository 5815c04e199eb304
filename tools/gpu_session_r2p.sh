#!/bin/bash
# GPU session P of round 2: SN tensor-core path as two kernels (coefficients at 12 warps per SM -> A fragments in HBM ->
# persistent tile kernel) -- parity suite, timing against the node-by-node kernel, benches, launch lists, ncu
cd "$(dirname "$0")/.."
O=gpurun_out/r2p; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -8 $O/pytest.log
for cfg in sn sn_curved sn_bao; do
  timeout 200 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1
  timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_s_$cfg.pt 2>&1 | tail -1
  PMCB200_SN_EXACT=1 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_exact_$cfg.pt 2>&1 | tail -1
  python tools/cmp_lp.py $O/lp_s_$cfg.pt $O/lp_exact_$cfg.pt
  rm -f $O/lp_*_$cfg.pt
done > $O/ab_sn_split.txt 2>&1
cat $O/ab_sn_split.txt
for n in 10000 100000 1000000; do timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1; done > $O/time_sn_small.txt 2>&1
cat $O/time_sn_small.txt
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
for f in sn c5 c4; do echo "$f: $(cut -c1-200 $O/bench_$f.json)"; done
for c in sn cmb_bao_sn; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$c.csv \
  python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_$c.log 2>&1
done
grep -E "k_sn_spec" $O/launches_sn.csv | tail -4 | cut -c1-100,280-
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sn_spec_coef -s 2 -c 1 -o $O/sn_spec_coef \
  python tools/time_sn.py --n 2000000 > $O/ncu_sn_spec_coef.log 2>&1
python tools/ncu_summary.py $O/sn_spec_coef.ncu-rep "k_sn_spec_coef<0,1> (flat), M=28, N=2e6 (one chunk)" > $O/sn_spec_coef_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sn_spec_tiles -s 2 -c 1 -o $O/sn_spec_tiles \
  python tools/time_sn.py --n 2000000 > $O/ncu_sn_spec_tiles.log 2>&1
python tools/ncu_summary.py $O/sn_spec_tiles.ncu-rep "k_sn_spec_tiles<1> (flat; 31 primary + 12 secondary tiles), M=28, N=2e6 (one chunk)" > $O/sn_spec_tiles_summary.txt
cat $O/*_summary.txt | grep -E "kernel:|gpu__time|pipe_fp64|dmma|lsu_wavefronts.avg|issue_active|dram__|registers|warps_active|long_scoreboard"
du -sh $O
