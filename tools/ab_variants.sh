#!/bin/bash
# Build named variants of libpmc_b200.so (extra nvcc -D flags) into variants/ and time the SN likelihood kernel of
# each, alternating, in one GPU call:
#   tools/ab_variants.sh build  "pf:-DSN_PREFETCH=1" "pf288:-DSN_PREFETCH=1 -DSN_BLOCK=288"     (here, no GPU needed)
#   gpurun -- 'bash tools/ab_variants.sh run 10000000 3 pf pf288 > gpurun_out/ab.txt 2>&1'       (on the GPU box)
# "run" always includes the default build (cosmopmc_b200/libpmc_b200.so) as the baseline.
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p variants
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    PMCB200_OBJ_DIR=$PWD/variants/obj_$name PMCB200_LIB_OUT=$PWD/variants/$name.so PMCB200_NVCC_FLAGS="$flags" \
      python -c "from cosmopmc_b200 import build as b; b.build()"
    echo "built variants/$name.so ($flags)"
  done
else
  n=$1; reps=$2; shift 2
  for r in $(seq 1 $reps); do
    echo -n "rep$r "; timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1
    for name in "$@"; do
      echo -n "rep$r "; PMCB200_LIB=$PWD/variants/$name.so timeout 100 python tools/time_sn.py --n $n 2>&1 | tail -1
    done
  done
fi
