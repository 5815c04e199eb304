#!/bin/bash
# GPU session AD of round 2: HEAD with the spectral distance to a* -- full parity suite, benches, launch list and ncu of k_like_cmbdp
cd "$(dirname "$0")/.."
O=gpurun_out/r2ad; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err; cut -c1-200 $O/bench_c5.json
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err; cut -c1-200 $O/bench_sn.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_cmb_bao_sn.csv \
  python bench.py --config cmb_bao_sn --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_like_cmbdp -s 1 -c 1 -o $O/c5_cmbdp_spec \
  python bench.py --config cmb_bao_sn --nsamples 4000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c5_cmbdp.log 2>&1
python tools/ncu_summary.py $O/c5_cmbdp_spec.ncu-rep "k_like_cmbdp<0> (spectral distance to a*), C5, N=4e6" > $O/c5_cmbdp_spec_summary.txt
rm -f $O/c5_cmbdp_spec.ncu-rep
grep -E "gpu__time|pipe_fp64|issue_active|registers|warps_active" $O/c5_cmbdp_spec_summary.txt
