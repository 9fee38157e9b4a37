#!/bin/bash
# dress rehearsal of the driver's 8-GPU launches on the final build: reference arm, then ours (default flags)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $TR bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_final_8gpu_reference.json 2> gpurun_out/f8_ref.err; echo "reference rc=$?" > gpurun_out/rc.txt
timeout 400 $TR bench.py --gpus 8 > gpurun_out/r02_final_8gpu.json 2> gpurun_out/f8.err; echo "ours rc=$?" >> gpurun_out/rc.txt
python - <<'PY'
import json
for f in ('r02_final_8gpu_reference', 'r02_final_8gpu'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, d.get('impl'), 'n', d['n_gpus'], round(d['value'], 1), d['unit'], round(d['ms_per_step'], 2), 'ms', 'e2e', d['e2e'] and round(d['e2e']['value'], 1), d['config'].get('grad_allreduce'))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/rc.txt
