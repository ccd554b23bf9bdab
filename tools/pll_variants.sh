#!/bin/bash
# Build k_pll variants on the GPU box and time each with bench.py (device-resident leg only) + a parity subset.
#   tools/pll_variants.sh "name1:-DFLAG ..." "name2:..."      results -> gpurun_out/variants.txt
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/variants.txt
mkdir -p gpurun_out build
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  LIB=$PWD/build/libsdr_b200_$name.so
  if [ "$name" = "default" ]; then LIB=$PWD/real-time-sdr_b200/libsdr_b200.so; elif [ -n "${PREBUILT:-}" ] && [ -f $LIB ]; then :; else
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
      -Xcompiler -fPIC,-ffp-contract=off -shared $flags -o $LIB \
      real-time-sdr_b200/csrc/sdr_chain.cu real-time-sdr_b200/csrc/sdr_design.cpp 2>>gpurun_out/variants.err || { echo "$name: build failed" >> $OUT; continue; }
  fi
  echo "== $name ($flags)" >> $OUT
  SDRB_LIB=$LIB timeout 600 python -m pytest tests/test_chain_gpu.py -x -q -k "single_stream_all_stages and 0-r or batch_equals or edge_inputs" 2>&1 | tail -1 >> $OUT
  for rep in 1 2; do
  SDRB_LIB=$LIB timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 128 --warmup 8 2>>gpurun_out/variants.err | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(json.dumps({'ms_per_step': d['ms_per_step'], 'value': d['value'], 'pll_cycles': r['pll_cycles_per_sample'], 'pll_ms_timed': r['kernel_ms'].get('pll'), 'serial': r['kernel_ms_serialised']}))
" >> $OUT
  done
done
cat $OUT
