#!/bin/bash
# GPU session M of round 2: warp-sliced EM statistics kernel (d <= 10) and the 12-warp SN kernel with the A operands in
# shared memory -- full parity suite, A/B against the previous kernels, benches, launch list, ncu
cd "$(dirname "$0")/.."
O=gpurun_out/r2m; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -8 $O/pytest.log
for cfg in sn sn_curved sn_bao; do
  timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_s_$cfg.pt 2>&1 | tail -1
  PMCB200_SN_A_SMEM=0 PMCB200_SN_TAIL32=0 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_r_$cfg.pt 2>&1 | tail -1
  PMCB200_SN_EXACT=1 timeout 200 python tools/time_sn.py --n 10000000 --config $cfg --save $O/lp_exact_$cfg.pt 2>&1 | tail -1
  python tools/cmp_lp.py $O/lp_s_$cfg.pt $O/lp_exact_$cfg.pt
  python tools/cmp_lp.py $O/lp_s_$cfg.pt $O/lp_r_$cfg.pt
  rm -f $O/lp_*_$cfg.pt
done > $O/ab_sn_a_smem.txt 2>&1
cat $O/ab_sn_a_smem.txt
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
PMCB200_EM_NO_WS=1 timeout 300 python bench.py --no-cpu-baseline > $O/bench_sn_nows.json 2> $O/bench_sn_nows.err
timeout 300 python bench.py --config banana --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
PMCB200_EM_NO_WS=1 timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5_nows.json 2> $O/bench_c5_nows.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
for f in sn sn_nows c3 c5 c5_nows c4; do echo "$f: $(cut -c1-200 $O/bench_$f.json)"; done
for c in sn cmb_bao_sn; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$c.csv \
  python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_$c.log 2>&1
done
grep -E "k_em_stats|k_like_sn_spec|k_weights|k_simulate" $O/launches_sn.csv | tail -8 | cut -c1-220
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_em_stats_mma_ws -s 1 -c 1 -o $O/c2_k_em_stats_mma_ws \
  python bench.py --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c2_em_ws.log 2>&1
python tools/ncu_summary.py $O/c2_k_em_stats_mma_ws.ncu-rep "k_em_stats_mma_ws, C2 (SN d=5 K=10), N=4e6, round 2 session M" > $O/c2_k_em_stats_mma_ws_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma_s -s 2 -c 1 -o $O/sn_spec_mma_s \
  python tools/time_sn.py --n 4000000 > $O/ncu_sn_spec_mma_s.log 2>&1
python tools/ncu_summary.py $O/sn_spec_mma_s.ncu-rep "k_like_sn_spec_mma_s<0,1> (flat, A operands in shared memory, 12 warps), M=28, N=4e6" > $O/sn_spec_mma_s_summary.txt
cat $O/*_summary.txt | grep -E "kernel:|gpu__time|pipe_fp64|dmma|lsu_wavefronts.avg|issue_active|dram__|registers|warps_active"
du -sh $O
