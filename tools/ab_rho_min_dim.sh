cd /root/repo
for c in sn sn_bao cmb_bao_sn; do
  for v in 5 10; do echo -n "$c rho_min_dim=$v: "; PMCB200_RHO_MIN_DIM=$v timeout 300 python bench.py --config $c --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"; done
done
