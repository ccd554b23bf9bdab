#!/bin/bash
# On the GPU box: the in-tree library under different environment settings: tools/ab_env_run.sh "VAR=a" "VAR=b" ... -> gpurun_out/ab_env.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_env.txt
: > $OUT
for setting in "$@"; do
  echo "== $setting" >> $OUT
  for rep in 1 2; do env $setting timeout 120 python tools/quick_time.py --streams 1024 --blocks 10 2>>gpurun_out/ab_env.err | head -1 >> $OUT; done
  env $setting timeout 120 python tools/quick_time.py --streams 4096 --blocks 6 2>>gpurun_out/ab_env.err | head -1 >> $OUT
done
cat $OUT
