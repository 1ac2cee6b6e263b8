#!/bin/bash
# usage (under gpurun): tools/prof_step.sh <tag>
#   bench line + per-shape detail, ncu launch list of one training step, ncu --set full of every kernel of one step (raw page as csv;
#   the .ncu-rep stays on the box: it is > 64 MiB), DRAM traffic per launch of the GEMM family.
TAG=$1
O=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-inference --no-stock-gpu"
TRU_BENCH_DETAIL=1 timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.detail || exit 1
timeout 300 $CMD > $O/plain_$TAG.log 2>&1 || exit 1
# one step = ~190 launches (torch's included); skip the 3 warm-up steps and the first timed one
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 760 -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"tc_|dw_|fgru|tgru|enc0|convt8|frontend|backend|loss_|bn_|wgrad_reduce|adamw|grad_sumsq" -s 668 -c 167 -f -o /tmp/full_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1
ncu -i /tmp/full_$TAG.ncu-rep --page raw --csv > $O/full_raw_$TAG.csv 2>/dev/null
for k in tc_wgrad_stream tc_igemm dw_bwd_stream tgru_bwd loss_fwd frontend_kernel backend_fwd; do
  ncu -i /tmp/full_$TAG.ncu-rep --page source --csv --kernel-name regex:$k > $O/src_${k}_$TAG.csv 2>/dev/null
done
ls -la $O | tail -12
