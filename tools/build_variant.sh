#!/bin/bash
# usage: tools/build_variant.sh NAME "DEFINES"   -> root_digger_b200/lib/variants/NAME/librdk_b200.so
set -e
cd "$(dirname "$0")/.."
mkdir -p root_digger_b200/lib/variants/$1
DEFS=""; for d in $2; do DEFS="$DEFS -D$d"; done
/usr/local/cuda/bin/nvcc $DEFS -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-ffp-contract=off,-Wall -shared -o root_digger_b200/lib/variants/$1/librdk_b200.so \
  root_digger_b200/csrc/rdk_abi.cu root_digger_b200/csrc/rdk_host_math.cpp -ldl
