#!/bin/bash
set -u
mkdir -p gpurun_out
T="timeout 1500 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
$T tests/test_gpu_gemm.py -m gpu -x > gpurun_out/t_gemm.log 2>&1; echo "gemm tests rc=$?" > gpurun_out/rc.txt
tail -3 gpurun_out/t_gemm.log
$T tests -m gpu --deselect tests/test_gpu_gemm.py > gpurun_out/t_rest.log 2>&1; echo "rest rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed|^FAILED" gpurun_out/t_rest.log | tail -8
export GB_NOLIB=1
for bn in 0 128 256; do
  GB_BN=$bn python tools/gemm_bench.py 384 > gpurun_out/bn_384_$bn.txt 2>&1
done
VITK_GEMM_192_2CTA=0 python tools/gemm_bench.py 384 > gpurun_out/bn_384_old192.txt 2>&1
GB_BN=128 python tools/gemm_bench.py 192 > gpurun_out/bn_192_128.txt 2>&1
GB_BN=0 python tools/gemm_bench.py 192 > gpurun_out/bn_192_0.txt 2>&1
paste -d'|' <(cut -c1-62 gpurun_out/bn_384_old192.txt) <(cut -c37-62 gpurun_out/bn_384_0.txt) <(cut -c37-62 gpurun_out/bn_384_128.txt) <(cut -c37-62 gpurun_out/bn_384_256.txt)
paste -d'|' <(cut -c1-62 gpurun_out/bn_192_0.txt) <(cut -c37-62 gpurun_out/bn_192_128.txt)
cat gpurun_out/rc.txt
timeout 600 python bench.py --config 2 --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/c2_pair192.json 2>/dev/null
VITK_GEMM_192_2CTA=0 timeout 600 python bench.py --config 2 --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/c2_old.json 2>/dev/null
python -c "
import json
for f in ('c2_pair192','c2_old'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['roofline']['by_shape_us'])
"
