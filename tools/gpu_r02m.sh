#!/bin/bash
# Round 2: fused conv1 + conv2 kernel with wider TMEM loads: parity + rate + ncu source capture
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_models.py -m gpu -q --tb=short -x > gpurun_out/test_fused.log 2>&1
echo "tests exit $? $(tail -n 1 gpurun_out/test_fused.log)" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -20
timeout 300 python tools/split_rate.py 2>&1 | head -1 | tee gpurun_out/split_rate.txt
timeout 120 python tools/prof_cnn2d_small.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv12_fused_kernel" -s 1 -c 1 -f -o gpurun_out/prof_conv12 python tools/prof_cnn2d_small.py > gpurun_out/ncu_conv12.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
