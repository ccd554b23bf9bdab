#!/bin/bash
# Build k_pll variants on the GPU box and time each with bench.py (device-resident leg only).
#   tools/pll_variants.sh "name1:-DFLAG ..." "name2:..."      results -> gpurun_out/variants.txt
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/variants.txt
mkdir -p gpurun_out
LIB=real-time-sdr_b200/libsdr_b200.so
cp $LIB /tmp/libsdr_default.so
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  if [ "$name" != "default" ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
      -Xcompiler -fPIC,-ffp-contract=off -shared $flags -o $LIB \
      real-time-sdr_b200/csrc/sdr_chain.cu real-time-sdr_b200/csrc/sdr_design.cpp || { echo "$name: build failed" >> $OUT; continue; }
  fi
  echo "== $name ($flags)" >> $OUT
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 48 --warmup 8 2>>gpurun_out/variants.err | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(json.dumps({'ms_per_step': d['ms_per_step'], 'value': d['value'], 'kernel_ms': r['kernel_ms'], 'serial': r.get('kernel_ms_serialised'), 'pll_cycles': d['fp32'].get('pll_cycles_per_step')}))
" >> $OUT
done
cp /tmp/libsdr_default.so $LIB
