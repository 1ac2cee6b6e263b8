#!/bin/bash
# DRAM bytes of every GEMM-kernel launch of one training step (4th step: 3 warm-up steps are skipped)
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-inference"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_igemm -s 156 -c 52 --csv --log-file $O/traffic_igemm.csv $CMD > $O/traffic_igemm.log 2>&1
