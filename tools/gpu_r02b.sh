#!/bin/bash
# Round 2, second GPU call: GPU suite after the conv1 weight split / error-diffusion rounding / fused 1D-CNN, rate of the fused
# 1D-CNN, corrected MMA micro-benchmark, the default bench line, and one ncu --set full capture of conv1_tc_kernel.
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
head -n 24 gpurun_out/umma_bench.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
timeout 120 python tools/prof_cnn2d_small.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_tc_kernel -s 1 -c 2 -f -o gpurun_out/prof_conv1 python tools/prof_cnn2d_small.py > gpurun_out/ncu_conv1.log 2>&1
echo "ncu conv1 exit $?" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -40
