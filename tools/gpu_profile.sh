#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): launch lists (shares) for the headline, hybrid and EER workloads, and full-set
# captures of the dominant kernels.  Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --pool 2080 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --e2e-pool 416"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|conv1_tc" -s 30 -c 3 -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set (cnn2d conv kernels) exit $?"
# hybrid workload: CAE, 1D-CNN, blend, select kernels
HCMD="python bench.py --workload hybrid --pool 4736 --steps 1 --warmup 3"
$HCMD > gpurun_out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/launches_hybrid.csv $HCMD > gpurun_out/ncu_list_h.log 2>&1
echo "hybrid launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:cae_enc1_tc_kernel|cae_final_tc_kernel|cnn1d_l1_fused_kernel" -s 8 -c 3 -f -o gpurun_out/prof_hybrid $HCMD > gpurun_out/ncu_full_h.log 2>&1
echo "full set (hybrid kernels) exit $?"
# EER workload: sort + select
ECMD="python bench.py --workload eer --steps 1 --warmup 3"
$ECMD > gpurun_out/eer_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/eer_launches.csv $ECMD > gpurun_out/eer_ncu_list.log 2>&1
echo "eer launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:select_hist_tma|radix_downsweep|radix_upsweep" -s 10 -c 6 -f -o gpurun_out/prof_eer $ECMD > gpurun_out/eer_ncu_full.log 2>&1
echo "full set (eer kernels) exit $?"
ls -la gpurun_out | head -40
