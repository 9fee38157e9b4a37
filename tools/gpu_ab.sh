#!/bin/bash
# A/B runs on ONE box (clocks differ from box to box): usage  bash tools/gpu_ab.sh
set -u
mkdir -p gpurun_out
T="timeout 1500 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
$T tests -m gpu -s > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/t_all.log | tail -3
python tools/diag_parity.py vit_large_patch16_384 2 > gpurun_out/diag_l384.log 2>&1
python tools/diag_parity.py vit_base_patch16_224 8 > gpurun_out/diag_b.log 2>&1
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 5"
for rep in 1 2; do
  VITK_GELU_AUX=bf16 $B --config 3 > gpurun_out/ab_c3_bf16_$rep.json 2>/dev/null
  VITK_GELU_AUX=q8   $B --config 3 > gpurun_out/ab_c3_q8_$rep.json 2>/dev/null
done
VITK_GELU_AUX=bf16 $B --config 2 > gpurun_out/ab_c2_bf16.json 2>/dev/null
VITK_GELU_AUX=q8   $B --config 2 > gpurun_out/ab_c2_q8.json 2>/dev/null
VITK_GELU_AUX=q8 VITK_GEMM_EW16_MAXK=512 $B --config 2 > gpurun_out/ab_c2_q8_ew16.json 2>/dev/null
VITK_GELU_AUX=q8 VITK_GEMM_EW16_MAXK=1024 $B --config 3 > gpurun_out/ab_c3_q8_ew16.json 2>/dev/null
VITK_GELU_AUX=q8   $B --config 5 --steps 10 > gpurun_out/ab_c5_q8.json 2>/dev/null
VITK_GELU_AUX=bf16 $B --config 5 --steps 10 > gpurun_out/ab_c5_bf16.json 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --config 4 --steps 10 --warmup 3 > gpurun_out/c4.json 2> gpurun_out/c4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_*.json') + ['gpurun_out/c4.json']):
    try:
        d = json.load(open(f))
        r = d['roofline']
        print(f"{f:40s} {d['value']:9.1f} img/s {d['ms_per_step']:7.2f} ms  clk {d['clocks']['sm_mhz']}  dom {r['kernel'][-34:]} {r['us_per_launch']:.1f} us  teacher {d['config'].get('teacher_fwd_ms_per_step')}")
        print("      ", {k.replace('gemm.', ''): v for k, v in list(r['by_shape_us'].items())[:6]})
    except Exception as e:
        print(f, 'ERR', e)
PY
