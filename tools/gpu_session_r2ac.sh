#!/bin/bash
# GPU session AC of round 2: spectral form of the distance to a* in k_like_cmbdp -- parity, timing against the node-by-node path
cd "$(dirname "$0")/.."
O=gpurun_out/r2ac; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "cmb or wmap or de_conservative or reference_test_suite or error_policy or build_variants" > $O/pytest_cmb.log 2>&1; tail -12 $O/pytest_cmb.log
for v in 0 1; do echo -n "PMCB200_CMB_EXACT=$v: "; PMCB200_CMB_EXACT=$v timeout 300 python bench.py --config cmb_bao_sn --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'likelihood stage', round(d['roofline']['kernel_ms'],3), d['counters_timed_region'])"; done
