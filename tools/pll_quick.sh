#!/bin/bash
# On the GPU box: parity of the chain (fast subset) and the device-resident bench leg -> gpurun_out/quick_<tag>.txt
tag=${1:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_chain_gpu.py -x -q -k "single_stream_all_stages or long_run or large_nco or batch_equals" > gpurun_out/quick_${tag}_pytest.log 2>&1
tail -2 gpurun_out/quick_${tag}_pytest.log
python bench.py --no-cpu-baseline --no-e2e --steps 128 --warmup 8 2>gpurun_out/quick_${tag}.err | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(json.dumps({'ms_per_step': d['ms_per_step'], 'value': d['value'], 'kernel_ms': r['kernel_ms'], 'serial': r.get('kernel_ms_serialised'), 'pll_cycles': d['fp32'].get('pll_cycles_per_step'), 'frac': d['fp32']['per_kernel_frac']}))
" | tee gpurun_out/quick_${tag}.txt
