#!/bin/bash
# ncu evidence for one short bench command (B200_PROFILING.md recipe): launch list + full set on the tensor-core kernels.
mkdir -p gpurun_out
CMD="python bench.py --pool 2080 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --e2e-pool 416"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|conv1_tc" -s 30 -c 3 -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | head -30
# launch list of the hybrid workload (CAE, 1D-CNN, blend, sort kernels)
HCMD="python bench.py --workload hybrid --pool 2080 --steps 1 --warmup 3"
$HCMD > gpurun_out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/launches_hybrid.csv $HCMD > gpurun_out/ncu_list_h.log 2>&1
echo "hybrid launch list exit $?"
