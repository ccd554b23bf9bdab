#!/bin/bash
# On the GPU box: device-resident bench for several station counts and SM budgets of k_pll -> gpurun_out/pll_sm_sweep.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/pll_sm_sweep.txt
: > $OUT
for S in ${STREAMS:-1024 2048 4096}; do
  for cap in ${CAPS:-64 32 16}; do
    SDRB_PLL_MAX_CTAS=$cap timeout 300 python bench.py --streams $S --no-cpu-baseline --no-e2e --steps 64 --warmup 8 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(json.dumps({'streams': $S, 'pll_sms': '$cap', 'ms_per_step': d['ms_per_step'], 'GS/s': round(d['value']/1e3,1), 'pll_ms': r['kernel_ms'].get('pll'), 'pll_serial': r['kernel_ms_serialised'].get('pll')}))
" >> $OUT
  done
done
cat $OUT
