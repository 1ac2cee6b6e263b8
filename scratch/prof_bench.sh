#!/bin/bash
# usage: scratch/prof_bench.sh <tag>   (run under gpurun)
#  launch list of the bench command at the real batch; full captures (batch 8: ncu's save/restore of a
#  24 GB workspace per replay is what makes batch 32 take 10+ minutes) exported to CSV on the box.
TAG=$1
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 624 -c 312 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
CMD8="$CMD --batch 8"
$CMD8 > $O/plain8_$TAG.log 2>&1 || exit 1
# all GEMM launches of one step, no source: raw metrics only
ncu --set full --clock-control none -k regex:tc_ -s 213 -c 71 -o /tmp/all_$TAG $CMD8 > $O/ncu_a_$TAG.log 2>&1
ncu -i /tmp/all_$TAG.ncu-rep --page raw --csv > $O/gemm_raw_$TAG.csv 2>/dev/null
# source-level capture of selected launches (index within a step given by SEL, after 3 warm-up steps)
for SEL in $2; do
  ncu --set full --clock-control none --import-source on -k regex:tc_ -s $((213 + SEL)) -c 1 -o $O/src_${TAG}_$SEL $CMD8 > $O/ncu_s_${TAG}_$SEL.log 2>&1
done
ls -la $O/
