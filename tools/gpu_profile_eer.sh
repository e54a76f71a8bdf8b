#!/bin/bash
# ncu evidence for the EER sort/sweep kernels (BASELINE config 5): launch list + one full-set capture of each kernel.
mkdir -p gpurun_out
N=${EER_N:-100000000}
CMD="python bench.py --workload eer --eer-n $N --steps 2 --warmup 3"
$CMD > gpurun_out/eer_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/eer_launches.csv $CMD > gpurun_out/eer_ncu_list.log 2>&1
echo "eer launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:radix_|sort_prep|sweep_min|onesweep|hist" -c 8 -f -o gpurun_out/prof_eer $CMD > gpurun_out/eer_ncu_full.log 2>&1
echo "eer full set exit $?"
cat gpurun_out/eer_plain.log
