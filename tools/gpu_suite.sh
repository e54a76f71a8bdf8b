#!/bin/bash
# Runs the GPU test files as separate processes (a device-side trap in one must not hide the rest),
# then smoke() and a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for t in probes eer models; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $?" | tee -a gpurun_out/summary.txt
  tail -n 30 gpurun_out/test_$t.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -n 8 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-6} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
cat gpurun_out/bench.json
