#!/bin/bash
# usage (on the GPU box): tools/collector_sweep.sh "0 1 2 ..."  -- per-kernel ms for A-operand collector masks (tc.cuh)
for m in $1; do
  AASIST_COLLECTOR=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('mask', $m, round(d['value']), d['clocks']['sm_mhz'], [(k['kernel'].replace('_tc','').replace('fused_conv1_conv2','f'), round(k['ms_per_step'], 2)) for k in d['kernels']])
"
done
