#!/bin/bash
# Round 2, session 4: ncu --set full of the roofline kernel (2D-CNN conv3) and of conv12_fused in the final state (one pass of 416 utterances)
mkdir -p gpurun_out
timeout 200 python tools/prof_cnn2d_small.py > gpurun_out/prof_cnn2d_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|conv12_fused_kernel" -s 2 -c 2 -f -o gpurun_out/prof_cnn2d_final python tools/prof_cnn2d_small.py > gpurun_out/ncu_cnn2d_final.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_cnn2d_final.ncu-rep --page raw --csv > gpurun_out/prof_cnn2d_final_raw.csv 2>/dev/null
ls -la gpurun_out | head
