#!/bin/bash
# GPU session U of round 2: HEAD -- full parity suite (new: synthetic supernova tables, EM shapes d = 9 / 6 / 4, warp-sliced
# against tile-sliced EM kernel), smoke, default bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2u; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -12 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err; cut -c1-260 $O/bench_sn.json
