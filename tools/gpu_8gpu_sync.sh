#!/bin/bash
# 8-GPU A/B of the gradient-sync schedules on config 3 (same box, back to back), plus the 1-GPU rate of that box.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
: > gpurun_out/rc.txt
for mode in step tail:6 tail:9 tail:3 step; do
  tag=$(echo $mode | tr -d ':')
  f=gpurun_out/r02_sync8_${tag}.json; [ -s $f ] && f=gpurun_out/r02_sync8_${tag}_b.json
  VITK_DP_SYNC=$mode timeout 300 $TR bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $f 2> gpurun_out/sync8.err
  echo "$mode rc=$?" >> gpurun_out/rc.txt
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02_sync8_1gpu.json 2>/dev/null; echo "1gpu rc=$?" >> gpurun_out/rc.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_sync8_*.json')):
    try:
        d = json.load(open(f)); print(f, 'n', d['n_gpus'], round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms', d['config']['grad_allreduce'], 'clk', d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/rc.txt
