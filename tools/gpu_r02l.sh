#!/bin/bash
# Round 2: conv1 + conv2 fused kernel (default): parity tests + rates
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "fused_conv1" > gpurun_out/test_fused.log 2>&1
echo "test_fused exit $? $(tail -n 1 gpurun_out/test_fused.log)" | tee -a gpurun_out/summary.txt
tail -n 30 gpurun_out/test_fused.log
for t in models round2; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short -x > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $? $(tail -n 1 gpurun_out/test_$t.log)" | tee -a gpurun_out/summary.txt
done
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -20
timeout 300 python tools/split_rate.py > gpurun_out/split_rate.txt 2>&1
cat gpurun_out/split_rate.txt
grep -A4 conv12 gpurun_out/parity_round2.json
