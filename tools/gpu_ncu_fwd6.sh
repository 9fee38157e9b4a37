#!/bin/bash
# one ncu capture (stall reasons per SASS instruction) of attn_fwd6 at N = 577, after the plain run has exited 0
set -u
mkdir -p gpurun_out
VITK_ATTN_FWD=6 GB_NOSDPA=1 timeout 100 python tools/attn_bench.py 64 577 16 > gpurun_out/fwd6_plain.txt 2>&1 || { cat gpurun_out/fwd6_plain.txt; exit 1; }
cat gpurun_out/fwd6_plain.txt
VITK_ATTN_FWD=6 GB_NOSDPA=1 GB_ITERS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd6 -s 2 -c 1 -f -o gpurun_out/prof_fwd6_577 python tools/attn_bench.py 64 577 16 > gpurun_out/ncu_fwd6.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_fwd6.log
