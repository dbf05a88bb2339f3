#!/bin/bash
# Build kernel variants for A/B timing under gpurun.
#   scripts/ab_variants.sh name1 "-DFLAG=1 ..." name2 "-D..." ...
# -> ebpmf.alpha_b200/lib/variants/libebpmf_<name>.so + ptxas_<name>.log; run each with
#   EBPMF_B200_LIB=ebpmf.alpha_b200/lib/variants/libebpmf_<name>.so python scripts/ab_run.py c4slab
set -e
cd "$(dirname "$0")/../ebpmf.alpha_b200/csrc"
mkdir -p ../lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
       -shared -o ../lib/variants/libebpmf_$name.so ebpmf_api.cu -lcudart -ldl -Xptxas -v 2> ../lib/variants/ptxas_$name.log &
done
wait
ls ../lib/variants
