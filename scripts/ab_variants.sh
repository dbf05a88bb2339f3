#!/bin/bash
# Build kernel variants for A/B timing under gpurun: "E MINB STREAM_MINB"
set -e
cd "$(dirname "$0")/../ebpmf.alpha_b200/csrc"
mkdir -p ../lib/variants
rm -f ../lib/variants/*.so
for v in "4 2 2" "4 2 3" "2 4 4"; do
  set -- $v
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DEBPMF_E=$1 -DEBPMF_MINB=$2 -DEBPMF_STREAM_MINB=$3 \
       -shared -o ../lib/variants/libebpmf_E$1_B$2_S$3.so ebpmf_api.cu -lcudart -ldl -Xptxas -v 2> ../lib/variants/ptxas_E$1_B$2_S$3.log &
done
wait
ls ../lib/variants
