#!/bin/bash
# Round 2, session 3: ncu --set full of the final one-sweep pass (form 1, a middle pass) and of the two private-counter histogram kernels
mkdir -p gpurun_out
EER_FORM=1 EER_N=100000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"radix_onesweep_kernel" -s 5 -c 1 -f -o gpurun_out/prof_onesweep_final python tools/prof_eer_small.py > gpurun_out/ncu_onesweep_final.log 2>&1
echo "ncu pass exit $?"
EER_FORM=5 EER_N=100000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sort_prep_hist_kernel" -s 1 -c 1 -f -o gpurun_out/prof_prep_hist0 python tools/prof_eer_small.py > gpurun_out/ncu_prep_hist0.log 2>&1
echo "ncu prep exit $?"
EER_FORM=3 EER_N=100000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"radix_hist_kernel" -s 3 -c 1 -f -o gpurun_out/prof_radix_hist python tools/prof_eer_small.py > gpurun_out/ncu_radix_hist.log 2>&1
echo "ncu hist exit $?"
