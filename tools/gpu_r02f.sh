#!/bin/bash
# Round 2, sixth GPU call: NaN-propagating ReLUs, coalesced producer mapping of the one-kernel 1D-CNN, metric mutex; full suite + bench.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for t in round2 probes models cae_layers dropin cli dlq eer bench_contract; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $? $(tail -n 1 gpurun_out/test_$t.log)" | tee -a gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python tools/c1d_rate.py > gpurun_out/c1d_rate.txt 2>&1
echo "c1d_rate exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/c1d_rate.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -40
