#!/bin/bash
# Round 2: conv1 / enc1 with half-accumulator hand-over: parity tests + rates
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for t in models cae_layers round2; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short -x > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $? $(tail -n 1 gpurun_out/test_$t.log)" | tee -a gpurun_out/summary.txt
done
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -20
timeout 300 python tools/split_rate.py > gpurun_out/split_rate.txt 2>&1
cat gpurun_out/split_rate.txt
timeout 300 python tools/model_rates.py 9472 2>&1 | head -4 | tee gpurun_out/model_rates.txt
