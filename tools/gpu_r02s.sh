#!/bin/bash
# Round 2: ncu source capture of the one-kernel 1D-CNN
mkdir -p gpurun_out
EER_N=0 timeout 120 python tools/prof_all_small.py > gpurun_out/prof_plain.log 2>&1 &&
EER_N=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cnn1d_fused_kernel" -s 1 -c 1 -f -o gpurun_out/prof_c1d python tools/prof_all_small.py > gpurun_out/ncu_c1d.log 2>&1
echo "ncu exit $?"
tail -n 2 gpurun_out/ncu_c1d.log
