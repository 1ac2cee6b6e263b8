#!/bin/bash
# usage: scratch/prof_bench.sh <tag>   (run under gpurun): full bench line + ncu launch list of the same command
TAG=$1
O=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
TRU_BENCH_DETAIL=1 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.detail || exit 1
$CMD > $O/plain_$TAG.log 2>&1 || exit 1
# 3 warm-up + 3 timed + 3 profiled steps of ~150 launches each: list the launches of one timed step
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 160 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
ls -la $O/
