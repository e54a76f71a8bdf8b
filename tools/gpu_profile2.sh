#!/bin/bash
# one ncu --set full capture of each remaining distinct hot kernel (the -s/-c windows of gpu_profile.sh catch repeats of the first ones)
mkdir -p gpurun_out
E1="python bench.py --workload eer --eer-method select --steps 1 --warmup 3"
$E1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:select_hist_tma" -s 4 -c 4 -f -o gpurun_out/prof_sel $E1 > gpurun_out/ncu_sel.log 2>&1
echo "select exit $?"
H="python bench.py --workload hybrid --pool 4736 --steps 1 --warmup 3"
$H > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:cnn1d_l1_fused|cae_mse_finish|xt_prep" -s 6 -c 4 -f -o gpurun_out/prof_c1d $H > gpurun_out/ncu_c1d2.log 2>&1
echo "cnn1d fused exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ConvCfg<2, 64, 32|ConvCfg<0, 128, 64|ConvCfg<1, 32, 64, 128, 80, 2, 3, 4, 1, 2" -s 6 -c 3 -f -o gpurun_out/prof_cae $H > gpurun_out/ncu_cae2.log 2>&1
echo "cae layers exit $?"
ls -la gpurun_out/*.ncu-rep
