#!/bin/bash
# On the GPU box: build the library with per-test careful-path counters (-DSDRB_PLL_DIAG) into build/ and run the drift probe with it.
cd "$(dirname "$0")/.."
mkdir -p build gpurun_out
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared \
  -DSDRB_PLL_DIAG -o build/libsdr_b200_diag.so real-time-sdr_b200/csrc/sdr_chain.cu real-time-sdr_b200/csrc/sdr_design.cpp || exit 1
SDRB_LIB=$PWD/build/libsdr_b200_diag.so python tools/pll_drift.py "$@"
