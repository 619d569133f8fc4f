#!/bin/bash
# usage (on the GPU box): tools/streams_sweep.sh [bench args]  -- step throughput with an encoder pass on one / two streams
for v in ${STREAMS:-1 2 1 2}; do AASIST_TC_STREAMS=$v python bench.py --no-cpu-baseline --no-eager-baseline "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('streams $v', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], round(d['ms_per_step'], 2))
"; done
