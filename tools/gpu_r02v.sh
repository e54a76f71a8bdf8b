#!/bin/bash
# Round 2: census test by mode + ncu source capture of the radix scatter kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "census" > gpurun_out/test_census.log 2>&1
echo "census exit $? $(tail -n 1 gpurun_out/test_census.log)"
grep -h "FAILED\|Error" gpurun_out/test_census.log | head
EER_N=100000000 timeout 200 python tools/prof_eer_small.py > gpurun_out/prof_plain.log 2>&1 &&
EER_N=100000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"radix_downsweep_kernel" -s 5 -c 1 -f -o gpurun_out/prof_down python tools/prof_eer_small.py > gpurun_out/ncu_down.log 2>&1
echo "ncu exit $?"
tail -n 2 gpurun_out/ncu_down.log
