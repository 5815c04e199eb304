timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sampler or philox or iteration_sn_demo or iteration_banana or component_selection or empty" 2>&1 | tail -2
timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; cut -c1-200 gpurun_out/bench_last.json; tail -2 gpurun_out/bench_last.err
