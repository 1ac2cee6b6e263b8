#!/bin/bash
# usage (under gpurun): tools/prof_bench.sh <tag>  -> full bench line + ncu launch list of one step + DRAM traffic pass
TAG=$1
O=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-inference"
TRU_BENCH_DETAIL=1 timeout 400 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.detail || exit 1
timeout 200 $CMD > $O/plain_$TAG.log 2>&1 || exit 1
# 3 warm-up + 3 timed + 3 profiled steps of ~175 launches each (torch's included): capture about two steps, the summary keeps one
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_igemm -s 162 -c 54 --csv --log-file $O/traffic_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-inference > $O/traffic_$TAG.log 2>&1
ls -la $O/ | tail -5
