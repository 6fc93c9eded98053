#!/bin/bash
# tools/build_variants.sh name "-DFLAG ..." [name "-DFLAG ..."]...  -> mpc-ntm-control_b200/lib/variants/name.so
set -e
cd "$(dirname "$0")/../mpc-ntm-control_b200/csrc"
mkdir -p ../lib/variants
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $2 -o ../lib/variants/$1.so ntm_kernels.cu ntm_cabi.cu &
  shift 2
done
wait
ls -la ../lib/variants
