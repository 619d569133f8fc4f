#!/bin/bash
# usage (on the GPU box): tools/chunk_sweep.sh  -- bench value for several encoder chunk sizes
for c in 128 256 512; do
  AASIST_TC_CHUNK=$c timeout 150 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | \
    python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('chunk', $c, round(d['value'], 1), d['clocks']['sm_mhz'])
"
done
