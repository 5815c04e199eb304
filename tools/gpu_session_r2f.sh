#!/bin/bash
# GPU session F of round 2: ncu captures at HEAD (spectral SN kernel without spills, the C3 kernels, k_like_cmbdp);
# summaries are made on the box, only three reports travel back (gpurun_out is capped at 64 MiB)
cd "$(dirname "$0")/.."
O=gpurun_out/r2f; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma_v2 \
  python tools/time_sn.py --n 4000000 > $O/ncu_sn_spec_mma.log 2>&1
python tools/ncu_summary.py $O/sn_spec_mma_v2.ncu-rep "k_like_sn_spec_mma<0,1> v2 (no spills), C2, N=4e6" > $O/sn_spec_mma_v2_summary.txt
for k in k_simulate_staged k_weights_multi k_em_stats_mma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/c3_$k \
    python bench.py --config banana --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c3_$k.log 2>&1
  python tools/ncu_summary.py $O/c3_$k.ncu-rep "$k, C3 (banana d=20 K=10), N=4e6, round 2 HEAD" > $O/c3_${k}_summary.txt
done
rm -f $O/c3_k_weights_multi.ncu-rep $O/c3_k_em_stats_mma.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_cmbdp -s 1 -c 1 -o $O/c5_cmbdp \
  python bench.py --config cmb_bao_sn --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c5_cmbdp.log 2>&1
python tools/ncu_summary.py $O/c5_cmbdp.ncu-rep "k_like_cmbdp<0> (lean integrand + warp-cooperative deep stages), C5, N=4e6" > $O/c5_cmbdp_summary.txt
timeout 200 python tools/time_weights.py --config banana --n 10000000 2>&1 | tail -1 > $O/estep_default.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c5.csv \
  python bench.py --config cmb_bao_sn --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_c5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_c2.log 2>&1
cat $O/*_summary.txt | grep -E "kernel:|gpu__time|fp64|dmma|lsu_wavefronts.avg|issue_active|dram__" ; cat $O/estep_default.txt
du -sh $O
