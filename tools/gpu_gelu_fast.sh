#!/bin/bash
# A/B on one box: the exact-erf GELU epilogue (A&S 7.1.26: MUFU.RCP + MUFU.EX2, ~17 math instructions per element) against the
# fitted one-MUFU approximant (gelu_fwd_bwd_fast, |err| <= 3.4e-5 / 9.3e-5: ~14) built with -DVITK_GELU_FAST into tools/tmp/
set -u
mkdir -p gpurun_out
FAST="VITK_LIB=$PWD/tools/tmp/libvitk_fastgelu.so VITK_ALLOW_STALE_LIB=1"
{
for lib in exact fast; do
  echo "==== $lib ===="
  if [ $lib = fast ]; then export VITK_LIB=$PWD/tools/tmp/libvitk_fastgelu.so VITK_ALLOW_STALE_LIB=1; fi
  GB_ONLY="fc1 fprop" GB_NOLIB=1 timeout 200 python tools/gemm_bench.py 384 768 1024 2>&1 | grep -v Warn
  for c in 2 3; do
    timeout 300 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/gelu_${lib}_c$c.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/gelu_${lib}_c$c.json')); print('config $c', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms clk', d['clocks']['sm_mhz'], {k: v for k, v in d['roofline']['by_shape_us'].items() if 'epi1' in k or 'epi4' in k})"
  done
done
} > gpurun_out/gelu_fast_ab.txt 2>&1
cat gpurun_out/gelu_fast_ab.txt
