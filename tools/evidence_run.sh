#!/bin/bash
# On the GPU box: everything the committed evidence of a round is made from -> gpurun_out/
#   tools/evidence_run.sh <tag>
tag=${1:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
[ -n "${SKIP_TESTS:-}" ] || { python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${tag}.log 2>&1; tail -2 gpurun_out/pytest_gpu_${tag}.log; }
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; tail -1 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
# ncu only after the plain runs exited: launch list of the bench command, then one full capture of one step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 21 -c 7 -o gpurun_out/prof_${tag} \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_full_${tag}.log 2>&1
[ -z "${SOAK_BLOCKS:-}" ] || timeout ${SOAK_TIMEOUT:-90} python tools/soak.py --blocks $SOAK_BLOCKS --stations 2 > gpurun_out/soak_${tag}.txt 2> gpurun_out/soak_${tag}.err
python - gpurun_out/bench_${tag}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "pll", r["pll_cycles_per_sample"], "fir", r["fir_frac_no_fma"],
      "capacity", r.get("capacity_value"), r.get("capacity_ms_per_step"), "sustained", r.get("sustained_value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
ls -la gpurun_out | tail -12
