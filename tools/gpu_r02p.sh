#!/bin/bash
# Round 2: fused conv1 + conv2 on CTA pairs: parity + sustained A/B
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "fused_conv1" > gpurun_out/test_fused.log 2>&1
echo "test_fused exit $? $(tail -n 1 gpurun_out/test_fused.log)" | tee -a gpurun_out/summary.txt
tail -n 25 gpurun_out/test_fused.log | cut -c1-300
timeout 300 python tools/ab_conv12.py 2>&1 | tee gpurun_out/ab_conv12.txt
