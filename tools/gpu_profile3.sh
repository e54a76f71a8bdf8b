#!/bin/bash
# r01f evidence: launch lists of the headline and the CAE workload, full-set captures of the three 2D-CNN conv kernels and of one
# whole CAE pass (prep, enc2-4, dec1-3).  Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv gpurun_out/eer_launches.csv
CMD="python bench.py --pool 2080 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --e2e-pool 416"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|conv1_tc" -s 30 -c 3 -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set (cnn2d conv kernels) exit $?"
CCMD="python bench.py --workload cae --pool 2368 --steps 1 --warmup 3"
$CCMD > gpurun_out/plain_cae.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_cae.csv $CCMD > gpurun_out/ncu_list_cae.log 2>&1
echo "cae launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|xt_prep|cae_enc1" -s 16 -c 8 -f -o gpurun_out/prof_cae $CCMD > gpurun_out/ncu_full_cae.log 2>&1
echo "full set (cae pass) exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
