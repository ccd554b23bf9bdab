#!/bin/bash
cd "$(dirname "$0")/.."
for S in 4096 2048; do
for cap in 16 24 32 48; do
  SDRB_PLL_MAX_CTAS=$cap timeout 300 python bench.py --streams $S --no-cpu-baseline --no-e2e --no-extras --steps 48 --warmup 6 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(json.dumps({'streams': $S, 'cap': $cap, 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'pll_timed': r['kernel_ms']['pll'], 'partition': d['config'].get('sm_partition')}))
"
done
done
