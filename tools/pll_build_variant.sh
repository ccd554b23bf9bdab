#!/bin/bash
# Build a k_pll variant library here (no GPU needed) so that tools/pll_variants.sh on the box only has to time it:
#   tools/pll_build_variant.sh name -DFLAG ...   ->  build/libsdr_b200_<name>.so   (then: PREBUILT=1 tools/pll_variants.sh "name:")
set -eu
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
  -Xcompiler -fPIC,-ffp-contract=off -shared "$@" -o build/libsdr_b200_$name.so \
  real-time-sdr_b200/csrc/sdr_chain.cu real-time-sdr_b200/csrc/sdr_design.cpp
