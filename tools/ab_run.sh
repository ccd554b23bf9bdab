#!/bin/bash
# On the GPU box: parity subset + serialised per-kernel times for prebuilt variant libraries (build/libsdr_b200_<name>.so,
# made here with tools/pll_build_variant.sh; "default" = the in-tree library).   tools/ab_run.sh name1 name2 ...  -> gpurun_out/ab.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab.txt
: > $OUT
for name in "$@"; do
  LIB=$PWD/build/libsdr_b200_$name.so
  [ "$name" = "default" ] && LIB=$PWD/real-time-sdr_b200/libsdr_b200.so
  echo "== $name" >> $OUT
  SDRB_LIB=$LIB timeout 300 python -m pytest tests/test_chain_gpu.py -x -q -k "${AB_TESTS:-single_stream_all_stages and 0-r or batch_equals or edge_inputs or golden}" 2>&1 | tail -3 >> $OUT
  for rep in 1 2; do SDRB_LIB=$LIB timeout 120 python tools/quick_time.py --streams 1024 --blocks 10 2>>gpurun_out/ab.err | head -1 >> $OUT; done
  SDRB_LIB=$LIB timeout 120 python tools/quick_time.py --streams 4096 --blocks 6 2>>gpurun_out/ab.err | head -1 >> $OUT
done
cat $OUT
