#!/bin/bash
# usage (under gpurun): tools/prof_infer.sh <tag>
#   1. per-launch-shape CUDA-event times of the two inference paths (4096-stream step, 37 x 10-s clips)  -> gpurun_out/infer_detail_<tag>.txt
#   2. ncu launch list of one pass of each (gpu__time_duration.sum, cold cache, serialised: compare shares) -> gpurun_out/infer_launches_<tag>.csv
TAG=$1
O=gpurun_out
timeout 200 python tools/prof_infer.py > $O/infer_detail_$TAG.txt 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/infer_launches_$TAG.csv \
    python tools/prof_infer.py --ncu > $O/infer_ncu_$TAG.log 2>&1
tail -3 $O/infer_ncu_$TAG.log
