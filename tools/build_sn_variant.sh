#!/bin/bash
# Build a variant of libpmc_b200.so that differs only in k_cosmo.cu's compile flags (the likelihood kernels):
#   tools/build_sn_variant.sh NAME "-DFLAG=1 ..."    -> variants/NAME.so   (use with PMCB200_LIB=variants/NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
mkdir -p variants/obj_$name
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda \
  -Xcompiler -fPIC,-O2 $flags -c cosmopmc_b200/csrc/k_cosmo.cu -o variants/obj_$name/k_cosmo.o
objs=$(ls cosmopmc_b200/build/*.o | grep -v k_cosmo.o)
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/$name.so $objs variants/obj_$name/k_cosmo.o
echo "built variants/$name.so ($flags)"
