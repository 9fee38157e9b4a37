#!/bin/bash
set -u
mkdir -p gpurun_out
T="timeout 1500 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
$T tests -m gpu -s > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/t_all.log | tail -3
python tools/gemm_bench.py 384 768 1024 > gpurun_out/gemm_shapes.txt 2>&1
cat gpurun_out/gemm_shapes.txt
export GB_ITERS=2 GB_NOLIB=1
for what in "qkv fprop" "fc1 fprop" "proj fprop"; do
  tag=$(echo $what | tr ' ' '_')
  GB_ONLY="$what" python tools/gemm_bench.py 384 > gpurun_out/plain_$tag.log 2>&1 &&
  GB_ONLY="$what" ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_s_$tag python tools/gemm_bench.py 384 > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?" >> gpurun_out/rc.txt
done
cat gpurun_out/rc.txt
