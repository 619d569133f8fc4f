#!/bin/bash
# usage (on the GPU box): tools/ab_bench.sh <variant.so>  -- per-kernel ms of the product build vs a variant build
for lib in "" "$1"; do
  AASIST_B200_LIB=${lib:+$PWD/$lib} timeout 150 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | \
    python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('${lib:-product}', round(d['value'], 1), d['clocks']['sm_mhz'], [(k['kernel'].split('.')[0] + '.' + k['kernel'].split('.')[-1][:5], round(k['ms_per_step'], 2)) for k in d['kernels']])
"
done
