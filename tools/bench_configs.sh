#!/bin/bash
# On the GPU box: the bench line of every BASELINE.json configuration -> gpurun_out/bench_<tag>_<config>.json
#   configs[0] m0 (mode 0 mono), configs[1] m2 (mode 2 mono, 147/800 resampler), configs[2] s0 (mode 0 stereo),
#   configs[3] r0 with one station (block latency), configs[4] r0 with 1024 stations (the headline)
tag=${1:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in m0 m2 s0; do
  timeout 600 python bench.py --config $cfg --no-extras > gpurun_out/bench_${tag}_${cfg}.json 2> gpurun_out/bench_${tag}_${cfg}.err
done
timeout 600 python bench.py --config r0 --streams 1 --no-extras > gpurun_out/bench_${tag}_r0_single.json 2> gpurun_out/bench_${tag}_r0_single.err
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
for f in gpurun_out/bench_${tag}*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], d.get("metric"), "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"),
          "cpu", (d.get("cpu_baseline") or {}).get("value"), "fir_frac", r.get("fir_frac_no_fma"), "pll_cyc", r.get("pll_cycles_per_sample"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
