#!/bin/bash
# GPU session Z of round 2: SN tensor-core kernel with fewer warps per block for small batches -- parity suite, small-batch timings, benches
cd "$(dirname "$0")/.."
O=gpurun_out/r2z; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -4 $O/pytest.log
for n in 4100 10000 30000 100000 1000000 10000000; do timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1; done > $O/time_sn_small.txt 2>&1
cat $O/time_sn_small.txt
timeout 300 python bench.py --nsamples 10000 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c1_1e4.json 2> $O/bench_c1.err; cut -c1-200 $O/bench_c1_1e4.json
timeout 400 python bench.py > $O/bench_sn.json 2> $O/bench_sn.err; cut -c1-200 $O/bench_sn.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
