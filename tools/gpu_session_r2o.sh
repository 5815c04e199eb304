#!/bin/bash
# GPU session O of round 2: persistent SN tensor-core kernel with the second warp of each scheduler half a task late,
# warp-sliced EM kernel with the feature matrix in shared memory -- parity suite, A/B over the stagger, benches, ncu
cd "$(dirname "$0")/.."
O=gpurun_out/r2o; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -8 $O/pytest.log
for cfg in sn sn_curved sn_bao; do
  for st in 20000 0 10000 40000; do
    echo -n "stagger $st: "; PMCB200_SN_STAGGER_NS=$st timeout 200 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1
  done
  timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_s_$cfg.pt 2>&1 | tail -1
  PMCB200_SN_EXACT=1 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_exact_$cfg.pt 2>&1 | tail -1
  python tools/cmp_lp.py $O/lp_s_$cfg.pt $O/lp_exact_$cfg.pt
  rm -f $O/lp_*_$cfg.pt
done > $O/ab_sn_stagger.txt 2>&1
cat $O/ab_sn_stagger.txt
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
for f in sn c5 c4; do echo "$f: $(cut -c1-200 $O/bench_$f.json)"; done
for c in sn cmb_bao_sn; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$c.csv \
  python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_$c.log 2>&1
done
grep -E "k_em_stats|k_like_sn_spec" $O/launches_sn.csv $O/launches_cmb_bao_sn.csv | tail -4 | cut -c1-120,300-
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_em_stats_mma_ws -s 1 -c 1 -o $O/c2_k_em_stats_mma_ws \
  python bench.py --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c2_em_ws.log 2>&1
python tools/ncu_summary.py $O/c2_k_em_stats_mma_ws.ncu-rep "k_em_stats_mma_ws (feature matrix in shared memory, next-step prefetch), C2 (SN d=5 K=10), N=4e6, round 2 session O" > $O/c2_k_em_stats_mma_ws_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma_v5 \
  python tools/time_sn.py --n 4000000 > $O/ncu_sn_spec_mma_v5.log 2>&1
python tools/ncu_summary.py $O/sn_spec_mma_v5.ncu-rep "k_like_sn_spec_mma<0,1,0> (flat; persistent, staggered warps; 31 primary + 12 secondary tiles), M=28, N=4e6" > $O/sn_spec_mma_v5_summary.txt
cat $O/*_summary.txt | grep -E "kernel:|gpu__time|pipe_fp64|dmma|lsu_wavefronts.avg|issue_active|dram__|registers|warps_active|long_scoreboard"
du -sh $O
