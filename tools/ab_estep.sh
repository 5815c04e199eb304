cd /root/repo
for c in sn cmb_bao_sn; do
  for v in 2 1 3; do echo -n "$c PMCB200_ESTEP=$v: "; PMCB200_ESTEP=$v timeout 300 python bench.py --config $c --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3))"; done
done
