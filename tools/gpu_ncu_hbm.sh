#!/bin/bash
# ncu --set full captures of the HBM-bound kernels of the step (LayerNorm forward / backward, AdamW) on the final build,
# each one launch out of a 2-step bench run, after the plain run has exited 0
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$BENCH > gpurun_out/plain_bench.log 2>&1 || { tail -5 gpurun_out/plain_bench.log; exit 1; }
for k in ln_fwd_kernel ln_bwd_ring_kernel adamw_kernel; do
  skip=40; [ $k = adamw_kernel ] && skip=3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_$k $BENCH > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
