#!/bin/bash
# GPU session E of round 2: spill-free spectral SN kernel; ncu of the C3 and C5 kernels at HEAD; E-step variant A/B;
# option A (unchanged driver) at 1e6 / 1e7 samples; compute-sanitizer record
cd "$(dirname "$0")/.."
O=gpurun_out/r2e; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
for cfg in sn sn_curved; do timeout 200 python tools/time_sn.py --n 10000000 --config $cfg 2>&1 | tail -1; done > $O/time_sn.txt 2>&1
cat $O/time_sn.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_sn_spec_mma -s 2 -c 1 -o $O/sn_spec_mma_v2 \
  python tools/time_sn.py --n 4000000 > $O/ncu_sn_spec_mma.log 2>&1
timeout 300 python bench.py --config banana --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
cut -c1-400 $O/bench_c3.json
for k in k_simulate_staged k_weights_multi k_em_stats_mma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/c3_$k \
    python bench.py --config banana --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c3_$k.log 2>&1
done
for e in 0 4; do PMCB200_ESTEP=$e timeout 200 python tools/time_weights.py --config banana --n 10000000 2>&1 | tail -1; done > $O/estep_ab.txt 2>&1
cat $O/estep_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_cmbdp -s 1 -c 1 -o $O/c5_cmbdp \
  python bench.py --config cmb_bao_sn --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c5_cmbdp.log 2>&1
timeout 400 python tools/run_option_a.py 1000000 3 > $O/option_a_1e6.txt 2>&1
timeout 900 python tools/run_option_a.py 10000000 2 > $O/option_a_1e7.txt 2>&1
cat $O/option_a_1e6.txt $O/option_a_1e7.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_iteration_sn_demo or test_posterior_sn_box or test_iteration_banana" > $O/sanitizer.log 2>&1
tail -5 $O/sanitizer.log
