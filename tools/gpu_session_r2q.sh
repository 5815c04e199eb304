#!/bin/bash
# GPU session Q of round 2: HEAD (persistent SN tensor-core kernel, warp-sliced EM kernel) -- parity suite, every bench,
# launch lists, ncu of the BAO kernel (C4) and of the final SN / EM kernels, build variants (cooperative Romberg stages of
# the sound-horizon integral from 128 nodes; split coefficient chain), compute-sanitizer
cd "$(dirname "$0")/.."
O=gpurun_out/r2q; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -4 $O/pytest.log
# variants first (short)
for v in default split2; do
  lib=""; [ $v != default ] && lib=$PWD/variants/$v.so
  for r in 1 2; do echo -n "$v sn: "; PMCB200_LIB=$lib timeout 200 python tools/time_sn.py --n 10000000 --config sn 2>&1 | tail -1; done
done > $O/ab_variants.txt 2>&1
for v in default coop128; do
  lib=""; [ $v != default ] && lib=$PWD/variants/$v.so
  for c in sn_bao cmb_bao_sn; do
    echo -n "$v $c: "; PMCB200_LIB=$lib timeout 300 python bench.py --config $c --no-cpu-baseline 2>/dev/null | cut -c1-260
  done
done >> $O/ab_variants.txt 2>&1
cat $O/ab_variants.txt | cut -c1-200
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config banana --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python bench.py --nsamples 10000 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c1_1e4.json 2> $O/bench_c1.err
for f in sn c3 c5 c4 c1_1e4; do echo "$f: $(cut -c1-200 $O/bench_$f.json)"; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
cut -c1-300 $O/bench_reference.json
for c in sn cmb_bao_sn sn_bao banana; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$c.csv \
  python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_bao -s 1 -c 1 -o $O/c4_k_like_bao \
  python bench.py --config sn_bao --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c4_bao.log 2>&1
python tools/ncu_summary.py $O/c4_k_like_bao.ncu-rep "k_like_bao<1> (d_z, w0-wa), C4, N=4e6" > $O/c4_k_like_bao_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_em_stats_mma_ws -s 1 -c 1 -o $O/c5_k_em_stats_mma_ws \
  python bench.py --config cmb_bao_sn --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c5_em_ws.log 2>&1
python tools/ncu_summary.py $O/c5_k_em_stats_mma_ws.ncu-rep "k_em_stats_mma_ws<8,4,0,1>, C5 (d=8 K=30), N=4e6" > $O/c5_k_em_stats_mma_ws_summary.txt
rm -f $O/c5_k_em_stats_mma_ws.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma_v6 \
  python tools/time_sn.py --n 4000000 > $O/ncu_sn_spec_mma_v6.log 2>&1
python tools/ncu_summary.py $O/sn_spec_mma_v6.ncu-rep "k_like_sn_spec_mma<0,1,0> (flat; persistent; 31 primary + 12 secondary tiles), M=28, N=4e6" > $O/sn_spec_mma_v6_summary.txt
rm -f $O/sn_spec_mma_v6.ncu-rep
cat $O/*_summary.txt | grep -E "kernel:|gpu__time|pipe_fp64|dmma|lsu_wavefronts.avg|issue_active|dram__|registers|warps_active|long_scoreboard"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_iteration_sn_demo or test_iteration_cmb_bao_sn or test_sn_spectral_large_batch or test_iteration_student_t" > $O/sanitizer.log 2>&1
tail -4 $O/sanitizer.log
du -sh $O
