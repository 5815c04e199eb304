#!/bin/bash
# GPU session AA of round 2: compute-sanitizer racecheck / memcheck over the kernels new in sessions M-Z
cd "$(dirname "$0")/.."
O=gpurun_out/r2aa; mkdir -p $O
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "em_tensor_core_shapes and (10-5 or 30-8 or 6-5 or 12-7 or 8-9)" > $O/racecheck_em.log 2>&1; tail -4 $O/racecheck_em.log
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sn_tile_layouts or iteration_sn_demo" > $O/racecheck_sn.log 2>&1; tail -4 $O/racecheck_sn.log
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sn_tile_layouts or build_variants_behind or em_tensor_core_shapes" > $O/memcheck.log 2>&1; tail -4 $O/memcheck.log
