#!/bin/bash
# GPU session AE of round 2 (final): HEAD -- full parity suite, smoke, C5 and default bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2ae; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err; cut -c1-200 $O/bench_c5.json
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err; cut -c1-200 $O/bench_sn.json
