#!/bin/bash
# On the GPU box: step time at 1024 stations with k_pll on 64 / 32 / 16 SMs (one / two / four warps per SM), with and without the green-context SM partition -> gpurun_out/cta_sweep.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for part in ${PARTS:-1 0}; do
for cap in ${CAPS:-64 32 16}; do
  SDRB_SM_PARTITION=$part SDRB_PLL_MAX_CTAS=$cap timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 128 --warmup 8 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(json.dumps({'partition': d['config'].get('sm_partition'), 'cap': $cap, 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'pll_cycles': r['pll_cycles_per_sample'], 'timed': r['kernel_ms'], 'serial_pll': r['kernel_ms_serialised'].get('pll'), 'fir_serial': r['fir_ms_serialised']}))
"
done
done | tee gpurun_out/cta_sweep.txt
