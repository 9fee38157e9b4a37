#!/bin/bash
# A/B on one box: GELU epilogue outputs through st.shared + TMA store (default) against direct 256-bit global stores.
# Record of a REVERTED experiment (profiles/r02_ab_gelu_direct_stores.txt): VITK_GEMM_GELU_STG is no longer read by the library.
set -u
mkdir -p gpurun_out
{
VITK_GEMM_GELU_STG=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -q -x -k "gelu or dgelu" 2>&1 | tail -3
for v in 0 1; do
  echo "==== VITK_GEMM_GELU_STG=$v ===="
  export VITK_GEMM_GELU_STG=$v
  GB_ONLY="fc1 fprop" GB_NOLIB=1 timeout 200 python tools/gemm_bench.py 384 768 1024 2>&1 | grep -v Warn
  for c in 3 2; do
    timeout 300 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/gelu_stg${v}_c$c.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/gelu_stg${v}_c$c.json')); print('config $c', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms clk', d['clocks']['sm_mhz'], {k: v for k, v in d['roofline']['by_shape_us'].items() if 'epi1' in k or 'epi4' in k or 'epi2' in k})"
  done
done
} > gpurun_out/gelu_stg_ab.txt 2>&1
cat gpurun_out/gelu_stg_ab.txt
