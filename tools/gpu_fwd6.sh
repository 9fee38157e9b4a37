#!/bin/bash
# attn_fwd6: parity under a short timeout (a protocol bug would hang), then timing next to the default kernels and SDPA
set -u
mkdir -p gpurun_out
VITK_ATTN_FWD=6 timeout 60 python tools/attn_trace6.py 2 197 3 > gpurun_out/fwd6_first.txt 2>&1; echo "first rc=$?" | tee -a gpurun_out/fwd6_first.txt
grep -q "first rc=0" gpurun_out/fwd6_first.txt || { cat gpurun_out/fwd6_first.txt; exit 1; }
VITK_ATTN_FWD=6 timeout 240 python -m pytest tests/test_gpu_attn.py -x -q -k "fwd or narrow" 2>&1 | tail -8 > gpurun_out/fwd6_tests.log; echo "tests rc=$?" >> gpurun_out/fwd6_tests.log
cat gpurun_out/fwd6_tests.log
if grep -q "passed" gpurun_out/fwd6_tests.log && ! grep -q "failed" gpurun_out/fwd6_tests.log; then
  echo "== default ==" > gpurun_out/fwd6_bench.txt
  timeout 200 python tools/attn_bench.py >> gpurun_out/fwd6_bench.txt 2>&1
  echo "== VITK_ATTN_FWD=6 ==" >> gpurun_out/fwd6_bench.txt
  VITK_ATTN_FWD=6 GB_NOSDPA=1 timeout 200 python tools/attn_bench.py >> gpurun_out/fwd6_bench.txt 2>&1
  echo "== VITK_ATTN_FWD=6 VITK_ATTN_FWD6_LAZY=8 ==" >> gpurun_out/fwd6_bench.txt
  VITK_ATTN_FWD=6 VITK_ATTN_FWD6_LAZY=8 GB_NOSDPA=1 timeout 200 python tools/attn_bench.py >> gpurun_out/fwd6_bench.txt 2>&1
  cat gpurun_out/fwd6_bench.txt
  VITK_ATTN_FWD=6 timeout 60 python tools/attn_trace6.py 64 577 16 > gpurun_out/trace6_577.txt 2>&1
  VITK_ATTN_FWD=6 timeout 60 python tools/attn_trace6.py 256 197 12 > gpurun_out/trace6_197.txt 2>&1
  cat gpurun_out/trace6_577.txt gpurun_out/trace6_197.txt
fi
