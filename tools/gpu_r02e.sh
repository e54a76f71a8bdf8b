#!/bin/bash
# Round 2, fifth GPU call: one-kernel 1D-CNN as the default; full GPU suite; bench; ncu --set full of cnn1d_fused_kernel.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for t in round2 probes models cae_layers dropin cli dlq eer bench_contract; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $? $(tail -n 1 gpurun_out/test_$t.log)" | tee -a gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench reference exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python tools/prof_c1d_small.py > gpurun_out/prof_c1d_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cnn1d_fused -s 1 -c 1 -f -o gpurun_out/prof_c1d_fused python tools/prof_c1d_small.py > gpurun_out/ncu_c1d.log 2>&1
echo "ncu cnn1d_fused exit $?" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -40
