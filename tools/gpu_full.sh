#!/bin/bash
# Full GPU validation: every -m gpu test file as its own process, smoke(), the default bench and the side workloads.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
: > gpurun_out/summary.txt
for t in probes eer models cae_layers dropin cli dlq bench_contract; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $? $(tail -n 1 gpurun_out/test_$t.log)" | tee -a gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-60} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/bench.err
cat gpurun_out/bench.json
for w in cae cnn1d hybrid; do
  timeout 600 python bench.py --workload $w --pool 9472 --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w exit $?" | tee -a gpurun_out/summary.txt
  cut -c1-330 gpurun_out/bench_$w.json
done
timeout 600 python bench.py --workload eer --steps 5 --warmup 3 > gpurun_out/bench_eer.json 2> gpurun_out/bench_eer.err
echo "bench eer exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_eer.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench reference exit $?" | tee -a gpurun_out/summary.txt
cut -c1-400 gpurun_out/bench_ref.json
