#!/bin/bash
# usage (under gpurun): tools/prof_step.sh <tag>
#   1. the full bench line (+ per-shape detail on stderr)                              -> gpurun_out/bench_<tag>.json / .detail
#   2. ncu launch list of one training step (gpu__time_duration.sum, cold cache)       -> gpurun_out/launches_<tag>.csv
#   3. ncu --set full of every launch of one step of the four kernel groups, raw page as csv (the .ncu-rep files stay on
#      the box: together they exceed what gpurun copies back)                          -> gpurun_out/full_<group>_<tag>.csv
TAG=$1
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-inference --no-stock-gpu"
TRU_BENCH_DETAIL=1 timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.detail || exit 1
timeout 300 $CMD > $O/plain_$TAG.log 2>&1 || exit 1
# warm-up = 3 steps of ~190 launches (torch's included); capture about two steps, the summary keeps one
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 570 -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
# kernels per training step of each group, counted between two front-end launches of the launch list
count() { python - "$O/launches_$TAG.csv" "$1" <<'PY'
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit() and r[12] == "gpu__time_duration.sum"]
st = [i for i, r in enumerate(rows) if "frontend_kernel" in r[4]]
one = rows[st[0]:st[1]] if len(st) >= 2 else rows
print(sum(1 for r in one if re.search(sys.argv[2], r[4])))
PY
}
grp() {  # name regex
  n=$(count "$2")
  echo "group $1: $n kernels per step" >> $O/ncu_groups_$TAG.log
  timeout 1500 ncu --set full --clock-control none -k regex:"$2" -s $((3 * n)) -c $n -f -o /tmp/full_$1_$TAG $CMD > $O/ncu_f_$1_$TAG.log 2>&1
  ncu -i /tmp/full_$1_$TAG.ncu-rep --page raw --csv > $O/full_$1_$TAG.csv 2>/dev/null
}
grp igemm "tc_igemm"
grp wgrad "tc_wgrad_stream"
grp dw "dw_fwd_stream|dw_bwd_stream"
grp misc "frontend_kernel|backend_|loss_fwd|loss_bwd|tgru_|fgru_|enc0_|convt8_|adamw_flat|grad_sumsq"
ls -la $O | tail -12
