#!/bin/bash
# Build kernel variants (E columns per unit, MINB resident CTAs) for A/B timing under gpurun.
set -e
cd "$(dirname "$0")/../ebpmf.alpha_b200/csrc"
mkdir -p ../lib/variants
for v in "2 2" "4 2" "2 3" "4 1"; do
  set -- $v
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DEBPMF_E=$1 -DEBPMF_MINB=$2 \
       -shared -o ../lib/variants/libebpmf_E$1_B$2.so ebpmf_api.cu -lcudart -ldl &
done
wait
ls -la ../lib/variants
