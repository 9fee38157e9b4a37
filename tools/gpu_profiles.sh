#!/bin/bash
set -u
mkdir -p gpurun_out
T="timeout 1500 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
$T tests -m gpu -s > gpurun_out/t_all.log 2>&1; echo "all rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed|^FAILED" gpurun_out/t_all.log | tail -6
python tools/attn_bench.py 2>&1 | grep -v Warn > gpurun_out/attn_table.txt; cat gpurun_out/attn_table.txt
python tools/patch_embed_bench.py 256 224 768 2>&1 | grep -v Warn > gpurun_out/patch_embed.txt
python tools/patch_embed_bench.py 64 384 1024 2>&1 | grep -v Warn >> gpurun_out/patch_embed.txt; cat gpurun_out/patch_embed.txt
python tools/gemm_bench.py 384 768 1024 2>&1 | grep -v Warn > gpurun_out/gemm_shapes_after.txt
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$BENCH > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/rc.txt
export GB_ITERS=2 GB_NOLIB=1
GB_ONLY="fc1 fprop" python tools/gemm_bench.py 768 > gpurun_out/plain_fc1.log 2>&1 &&
GB_ONLY="fc1 fprop" ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_fc1_768 python tools/gemm_bench.py 768 > gpurun_out/ncu_fc1.log 2>&1
echo "ncu fc1 rc=$?" >> gpurun_out/rc.txt
PE_ONLY=tma python tools/patch_embed_bench.py 256 224 768 > gpurun_out/plain_pe.log 2>&1 &&
PE_ONLY=tma ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_patch_tma python tools/patch_embed_bench.py 256 224 768 > gpurun_out/ncu_pe.log 2>&1
echo "ncu patch rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
