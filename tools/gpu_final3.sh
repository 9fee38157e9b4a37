#!/bin/bash
# end-of-round ncu refresh on the final build: launch list of one training step + full capture of the dominant kernel
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$BENCH > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" > gpurun_out/rc.txt
export GB_ITERS=2 GB_NOLIB=1
GB_ONLY="fc1 fprop" python tools/gemm_bench.py 768 > gpurun_out/plain_fc1.log 2>&1 &&
GB_ONLY="fc1 fprop" ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_fc1_768_final python tools/gemm_bench.py 768 > gpurun_out/ncu_fc1.log 2>&1
echo "ncu fc1 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
