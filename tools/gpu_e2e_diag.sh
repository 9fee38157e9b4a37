#!/bin/bash
set -u
mkdir -p gpurun_out
B="timeout 600 python bench.py --no-cpu-baseline --steps 30 --warmup 5"
$B > gpurun_out/d_default.json 2>/dev/null
VITK_BENCH_NO_GEMM_TIMING=1 $B > gpurun_out/d_notiming.json 2>/dev/null
VITK_BENCH_NO_SAMPLER=1 $B > gpurun_out/d_nosampler.json 2>/dev/null
VITK_BENCH_NO_GEMM_TIMING=1 VITK_BENCH_NO_SAMPLER=1 $B > gpurun_out/d_neither.json 2>/dev/null
VITK_BENCH_E2E_NO_MIXUP=1 $B > gpurun_out/d_nomixup.json 2>/dev/null
$B > gpurun_out/d_default2.json 2>/dev/null
python - <<'PY'
import json
for f in ('d_default','d_notiming','d_nosampler','d_neither','d_nomixup','d_default2'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f"{f:14s} value {d['value']:8.1f} ({d['ms_per_step']:.2f} ms)  e2e {d['e2e']['value']:8.1f}  clk {d['clocks']['sm_mhz']} pw {d['clocks'].get('power_w_max')}")
    except Exception as e: print(f, 'ERR', e)
PY
