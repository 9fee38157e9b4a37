#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/build/tma_probe > gpurun_out/tma_probe.txt 2>&1; echo "probe rc=$?" > gpurun_out/rc.txt
cat gpurun_out/tma_probe.txt
export GB_NOLIB=1 GB_ONLY="wgrad"
for s in 0 1 2 4 8; do GB_SPLITS=$s python tools/gemm_bench.py 1024 2>&1 | grep -v Warn > gpurun_out/splits_1024_$s.txt; done
paste -d'|' <(cut -c1-62 gpurun_out/splits_1024_0.txt) <(cut -c37-62 gpurun_out/splits_1024_1.txt) <(cut -c37-62 gpurun_out/splits_1024_2.txt) <(cut -c37-62 gpurun_out/splits_1024_4.txt) <(cut -c37-62 gpurun_out/splits_1024_8.txt)
for s in 0 1 2 4; do GB_SPLITS=$s python tools/gemm_bench.py 384 2>&1 | grep -v Warn > gpurun_out/splits_384_$s.txt; done
paste -d'|' <(cut -c1-62 gpurun_out/splits_384_0.txt) <(cut -c37-62 gpurun_out/splits_384_1.txt) <(cut -c37-62 gpurun_out/splits_384_2.txt) <(cut -c37-62 gpurun_out/splits_384_4.txt)
unset GB_ONLY
T="timeout 1500 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
$T tests -m gpu --deselect tests/test_gpu_gemm.py::test_patch_embed_image_operand > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed|^FAILED" gpurun_out/t_all.log | tail -8
cat gpurun_out/rc.txt
