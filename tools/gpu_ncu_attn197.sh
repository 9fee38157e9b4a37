#!/bin/bash
# ncu captures (stall reasons per SASS instruction) of the N = 197 attention kernels, after the plain run has exited 0
set -u
mkdir -p gpurun_out
GB_NOSDPA=1 timeout 100 python tools/attn_bench.py 256 197 12 > gpurun_out/attn197_plain.txt 2>&1 || { cat gpurun_out/attn197_plain.txt; exit 1; }
cat gpurun_out/attn197_plain.txt
GB_NOSDPA=1 GB_ITERS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd4 -s 2 -c 1 -f -o gpurun_out/prof_fwd4_197 python tools/attn_bench.py 256 197 12 > gpurun_out/ncu_fwd4.log 2>&1
echo "ncu fwd4 rc=$?"
GB_NOSDPA=1 GB_ITERS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd4 -s 2 -c 1 -f -o gpurun_out/prof_bwd4_197 python tools/attn_bench.py 256 197 12 > gpurun_out/ncu_bwd4.log 2>&1
echo "ncu bwd4 rc=$?"
