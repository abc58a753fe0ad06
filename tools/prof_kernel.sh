#!/bin/bash
# ncu --set full capture of kernels matching $2 (regex) during a short bench run; tag $1; skip $3 launches
set -u
TAG=$1; PAT=$2; SKIP=${3:-2}; CNT=${4:-2}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-extras"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s $SKIP -c $CNT -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_$TAG.log 2>&1
tail -c 600 $OUT/plain_$TAG.log; tail -5 $OUT/ncu_$TAG.log
