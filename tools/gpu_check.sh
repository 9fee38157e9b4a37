#!/bin/bash
# One GPU-box visit: the GPU test suite (risky new paths in their own process), then the benches.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
T="timeout 1500 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
$T tests -m gpu --deselect tests/test_gpu_attn.py::test_attn_narrow_heads --deselect tests/test_gpu_configs.py -x -s > gpurun_out/t_main.log 2>&1; echo "main rc=$?" >> gpurun_out/rc.txt
$T tests/test_gpu_attn.py -m gpu -k narrow_heads > gpurun_out/t_narrow.log 2>&1; echo "narrow rc=$?" >> gpurun_out/rc.txt
$T tests/test_gpu_configs.py -m gpu -s > gpurun_out/t_configs.log 2>&1; echo "configs rc=$?" >> gpurun_out/rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench3 rc=$?" >> gpurun_out/rc.txt
for c in 2 4 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --breakdown > gpurun_out/bench$c.json 2> gpurun_out/bench$c.err; echo "bench$c rc=$?" >> gpurun_out/rc.txt
done
cat gpurun_out/rc.txt
tail -5 gpurun_out/t_main.log gpurun_out/t_narrow.log gpurun_out/t_configs.log
