#!/bin/bash
# Round 2, third GPU call: everything after the elect_one_sync() issue-loop fix (MMA issue no longer serialised by the compiler's
# ELECT / BRA.U.ANY loops), parallel radix scan; per-kernel launch lists by ncu.
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
timeout 300 python tools/umma_bench.py > gpurun_out/umma_bench.txt 2>&1
echo "umma_bench exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
timeout 300 python tools/prof_all_small.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_all.csv python tools/prof_all_small.py > gpurun_out/ncu_all.log 2>&1
echo "ncu launch list exit $?" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -40
