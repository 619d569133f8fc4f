#!/bin/bash
# usage (on the GPU box): tools/profile_round.sh <outdir under gpurun_out>  -- the bench lines of every BASELINE configuration
# on one GPU (contract line with all baselines; the others without the CPU / eager legs) + the ncu capture
OUT=gpurun_out/$1; mkdir -p $OUT
Q="--no-cpu-baseline --no-eager-baseline"
python bench.py > $OUT/bench_c2_aasist_b512.json 2> $OUT/bench_c2.err
python bench.py --model AASIST-L > $OUT/bench_aasist_l_b512.json 2>> $OUT/err.log
python bench.py --model RawGAT-ST $Q > $OUT/bench_c4_rawgat_b512.json 2>> $OUT/err.log
python bench.py --model AASIST2 --batch 256 $Q > $OUT/bench_aasist2_res2net_b256.json 2>> $OUT/err.log
for L in 16000 32000 64000 96000 128000 192000 256000; do
  python bench.py --batch 256 --samples $L --steps 10 $Q > $OUT/bench_c5_b256_L$L.json 2>> $OUT/err.log
done
python bench.py --workload evalset $Q > $OUT/evalset_n1.json 2>> $OUT/err.log
tools/ncu_capture.sh $1/ncu_full 2>&1 | tail -2
for f in $OUT/*.json; do python - "$f" <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1], round(d['value']), 'e2e', round(d.get('e2e', {}).get('value', 0)), d['clocks']['sm_mhz'], d.get('roofline', {}).get('frac'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
P
done
