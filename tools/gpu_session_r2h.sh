#!/bin/bash
# GPU session H of round 2: HEAD (Chebyshev order 28, f_K without a select) -- full parity suite, all benches, launch lists,
# ncu of C2's small kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r2h; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -8 $O/pytest.log
for cfg in sn sn_curved sn_bao; do timeout 200 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1; done > $O/time_sn.txt 2>&1
cat $O/time_sn.txt
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err
timeout 300 python bench.py --config banana --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config sn_bao --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
cat $O/bench_sn.json $O/bench_c3.json $O/bench_c5.json $O/bench_c4.json | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
cut -c1-300 $O/bench_reference.json
for c in sn cmb_bao_sn sn_bao; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$c.csv \
  python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_$c.log 2>&1
done
for k in k_em_stats_mma k_weights_multi k_simulate; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/c2_$k \
    python bench.py --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c2_$k.log 2>&1
  python tools/ncu_summary.py $O/c2_$k.ncu-rep "$k, C2 (SN d=5 K=10), N=4e6, round 2" > $O/c2_${k}_summary.txt
done
rm -f $O/c2_k_weights_multi.ncu-rep $O/c2_k_simulate.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma_curved \
  python tools/time_sn.py --n 4000000 --config sn_curved > $O/ncu_sn_spec_mma_curved.log 2>&1
python tools/ncu_summary.py $O/sn_spec_mma_curved.ncu-rep "k_like_sn_spec_mma<0,0> (curved), M=28, N=4e6" > $O/sn_spec_mma_curved_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma_v3 \
  python tools/time_sn.py --n 4000000 > $O/ncu_sn_spec_mma_v3.log 2>&1
python tools/ncu_summary.py $O/sn_spec_mma_v3.ncu-rep "k_like_sn_spec_mma<0,1> (flat), M=28, N=4e6" > $O/sn_spec_mma_v3_summary.txt
rm -f $O/sn_spec_mma_curved.ncu-rep
cat $O/*_summary.txt | grep -E "kernel:|gpu__time|pipe_fp64|dmma|lsu_wavefronts.avg|issue_active|dram__"
du -sh $O
