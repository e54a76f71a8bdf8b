#!/bin/bash
# Round 2: full GPU suite + smoke with the fused default, sustained A/B fused vs separate, launch list of every kernel
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/test_all.log 2>&1
echo "test_all exit $? $(tail -n 1 gpurun_out/test_all.log)" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_all.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $? $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
timeout 300 python tools/ab_conv12.py 2>&1 | tee gpurun_out/ab_conv12.txt
EER_N=100000000 timeout 300 python tools/prof_all_small.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_all.csv python tools/prof_all_small.py > gpurun_out/ncu_all.log 2>&1
echo "ncu launch list exit $?" | tee -a gpurun_out/summary.txt
