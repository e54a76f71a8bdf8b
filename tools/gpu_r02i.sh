#!/bin/bash
# Round 2: split precision (2D-CNN on the tensor cores with value + residual operands): tests + rate
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "split" > gpurun_out/test_split.log 2>&1
echo "test_split exit $? $(tail -n 1 gpurun_out/test_split.log)" | tee -a gpurun_out/summary.txt
tail -n 40 gpurun_out/test_split.log
timeout 300 python tools/split_rate.py > gpurun_out/split_rate.txt 2>&1
echo "split_rate exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/split_rate.txt
cat gpurun_out/parity_round2.json 2>/dev/null | head -60
