#!/bin/bash
# On an N-GPU box (gpurun --gpus N): the bench line at N ranks and the bare concurrent pinned-copy rate of the box.
#   tools/multi_gpu_run.sh <N> <tag>   -> gpurun_out/bench_<tag>_<N>gpu.json, gpurun_out/h2d_concurrent_<tag>.json (appended)
N=$1; tag=${2:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python tools/h2d_concurrent.py 2>> gpurun_out/h2d_${tag}.err | grep "^{" >> gpurun_out/h2d_concurrent_${tag}_${N}.json
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/h2d_concurrent.py 2>> gpurun_out/h2d_${tag}.err | grep "^{" >> gpurun_out/h2d_concurrent_${tag}_${N}.json
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err
  tail -c 600 gpurun_out/bench_${tag}_${N}gpu.err
  python - gpurun_out/bench_${tag}_${N}gpu.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "strong", {k: v for k, v in d["roofline"].items() if k.startswith("strong")})
PY
fi
tail -2 gpurun_out/h2d_concurrent_${tag}_${N}.json
