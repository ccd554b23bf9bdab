#!/bin/bash
# On the GPU box, with a library built with -DSDRB_PLL_TIMESTAMPS (tools/pll_build_variant.sh ts -DSDRB_PLL_TIMESTAMPS):
# true duration of k_pll and the gap between two launches from %globaltimer, in the overlapped step.
cd "$(dirname "$0")/.."
SDRB_LIB=$PWD/build/libsdr_b200_ts.so python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 64 --warmup 8 2>gpurun_out/ts.err >gpurun_out/ts.json
grep PLLTS gpurun_out/ts.err | python -c "
import sys, collections, statistics as st
rows=[tuple(map(int,l.split()[1:6])) for l in sys.stdin if l.startswith('PLLTS')]
# group launches: CTAs of one launch start within a few us of each other
rows.sort()
launches=[]; cur=[rows[0]]
for r in rows[1:]:
    if r[0]-cur[0][0] > 200000: launches.append(cur); cur=[r]
    else: cur.append(r)
launches.append(cur)
launches=[l for l in launches if len(l)==len(launches[len(launches)//2])][8:40]   # overlapped, timed steps
durs=collections.defaultdict(list)
for l in launches:
    t0=min(r[0] for r in l)
    for r in l: durs[(r[2],r[3])].append(((r[0]-t0)/1e3,(r[1]-r[0])/1e3,r[4]))
print('launches',len(launches),'CTAs per launch',len(launches[0]))
print('kernel span us (first start to last end), median',st.median([(max(r[1] for r in l)-min(r[0] for r in l))/1e3 for l in launches]))
print('gap us between launches (last end to next first start), median',st.median([(min(r[0] for r in launches[i+1])-max(r[1] for r in launches[i]))/1e3 for i in range(len(launches)-1)]))
for k in sorted(durs):
    v=durs[k]
    print(k,'start+%.1f'%st.median([x[0] for x in v]),'dur %.1f'%st.median([x[1] for x in v]),'sm',v[0][2])
"
