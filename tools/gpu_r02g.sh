#!/bin/bash
# Round 2, seventh GPU call: conv1 / enc1 bias through the tensor core; tests + bench + launch list.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for t in round2 models cae_layers dropin cli bench_contract; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $? $(tail -n 1 gpurun_out/test_$t.log)" | tee -a gpurun_out/summary.txt
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
EER_N=0 timeout 300 python tools/prof_all_small.py > gpurun_out/prof_plain.log 2>&1 &&
EER_N=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_all.csv python tools/prof_all_small.py > gpurun_out/ncu_all.log 2>&1
echo "ncu launch list exit $?" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -40
