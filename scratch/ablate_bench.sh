#!/bin/bash
# per-kernel times of the training step with parts of the GEMM kernel disabled
for f in 0 1 2 8 16 10; do
  TRU_DBG_FLAGS=$f TRU_BENCH_DETAIL=1 timeout 120 python bench.py --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$f.json 2> gpurun_out/ab_$f.detail
done
