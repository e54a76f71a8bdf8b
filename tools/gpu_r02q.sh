#!/bin/bash
# Round 2: ncu source capture of the CTA-pair fused kernel
mkdir -p gpurun_out
timeout 120 python tools/prof_cnn2d_small.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv12_fused_kernel" -s 1 -c 1 -f -o gpurun_out/prof_conv12p python tools/prof_cnn2d_small.py > gpurun_out/ncu_conv12p.log 2>&1
echo "ncu exit $?"
tail -n 3 gpurun_out/ncu_conv12p.log
