#!/bin/bash
# usage (on the GPU box): tools/xslots_sweep.sh  -- fused 32->32 block with 3 or 4 x-ring slots (v ring gets the rest)
for x in 3 4; do
  AASIST_BF_XSLOTS=$x timeout 150 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | \
    python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('xslots', $x, round(d['value'], 1), [(k['kernel'], round(k['ms_per_step'], 2)) for k in d['kernels'] if 'enc1' in k['kernel']])
"
done
