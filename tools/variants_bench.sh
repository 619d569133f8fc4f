#!/bin/bash
# usage (on the GPU box): tools/variants_bench.sh "<variant.so|product> ..." [bench args]  -- per-kernel ms per library
LIBS=$1; shift
for lib in $LIBS; do
  L=""; [ "$lib" != product ] && L=$PWD/$lib
  AASIST_B200_LIB=$L timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-eager-baseline "$@" 2>gpurun_out/variant_err.log | \
    python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', round(d['value'], 1), d['clocks']['sm_mhz'], [(k['kernel'].split('.')[0] + '.' + k['kernel'].split('.')[-1][:5], round(k['ms_per_step'], 2)) for k in d['kernels']])
"
  grep "stats\]" gpurun_out/variant_err.log | tail -2
done
