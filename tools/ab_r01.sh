timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for m in 0 2; do PMCB200_ESTEP=$m timeout 100 python tools/time_weights.py --n 10000000 --config sn 2>&1 | tail -1; done
PMCB200_EM_NO_MMA=1 timeout 100 python tools/time_weights.py --n 10000000 --config sn 2>&1 | tail -1
for c in banana cmb_bao_sn; do timeout 100 python tools/time_weights.py --n 10000000 --config $c 2>&1 | tail -1; done
PMCB200_EM_NO_MMA=1 timeout 100 python tools/time_weights.py --n 10000000 --config cmb_bao_sn 2>&1 | tail -1
