O=gpurun_out
TAG=r02
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-inference --no-stock-gpu"
TRU_BENCH_DETAIL=1 timeout 170 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.detail; echo "bench rc=$?"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -s 570 -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 70 ncu --set full --clock-control none -k regex:"tgru_" -s 6 -c 2 -f -o /tmp/full_tgru_$TAG $CMD > $O/ncu_f_tgru_$TAG.log 2>&1; echo "tgru full rc=$?"
ncu -i /tmp/full_tgru_$TAG.ncu-rep --page raw --csv > $O/full_tgru_$TAG.csv 2>/dev/null
ls -la $O/bench_$TAG.json $O/launches_$TAG.csv $O/full_tgru_$TAG.csv
