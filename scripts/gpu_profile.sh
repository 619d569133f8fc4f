#!/bin/bash
# Run ON THE GPU BOX (via gpurun): plain run, ncu launch list, and one full capture of chosen kernels.
# usage: scripts/gpu_profile.sh <tag> <kernel-regex> [batch] [skip] [count]
set -u
TAG=${1:-r01}; KRE=${2:-conv_tc_kernel}; BATCH=${3:-128}; SKIP=${4:-0}; COUNT=${5:-2}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --batch $BATCH --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $COUNT -o $OUT/prof_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
