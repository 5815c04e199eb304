timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python bench.py --steps 5 --warmup 3 --config banana --no-cpu-baseline > gpurun_out/bench_rho_banana.json 2> gpurun_out/bench_rho.err; cut -c1-260 gpurun_out/bench_rho_banana.json; tail -2 gpurun_out/bench_rho.err
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/l_banana4.csv python tools/run_stage.py --config banana --n 10000000 --reps 2 --stage iteration > /dev/null 2>&1
grep -E "k_em_stats|k_weights" gpurun_out/l_banana4.csv | cut -d, -f5,15 | cut -c1-120
