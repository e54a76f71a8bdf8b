#!/bin/bash
# Round 2: full GPU suite after the split-precision template changes + ncu source captures of conv1 / enc1 (current kernels)
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/test_all.log 2>&1
echo "test_all exit $? $(tail -n 1 gpurun_out/test_all.log)" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_all.log | head -20
EER_N=0 timeout 120 python tools/prof_all_small.py > gpurun_out/prof_plain.log 2>&1 &&
EER_N=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_tc_kernel|cae_enc1_tc_kernel" -s 2 -c 2 -f -o gpurun_out/prof_conv1_enc1 python tools/prof_all_small.py > gpurun_out/ncu_conv1.log 2>&1
echo "ncu conv1/enc1 exit $?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/ncu_conv1.log
