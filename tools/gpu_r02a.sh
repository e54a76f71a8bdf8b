#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite (new group / trained-like / census / NaN tests included), smoke(), the
# micro-benchmarks (TMEM read-out, MMA nmma sweep, H2D), and the default bench line with all five configs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
: > gpurun_out/summary.txt
for t in round2 probes eer models cae_layers dropin cli dlq bench_contract; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q --tb=short > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $? $(tail -n 1 gpurun_out/test_$t.log)" | tee -a gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python tools/micro/tmem_ld_bench.py > gpurun_out/tmem_ld_bench.txt 2>&1
echo "tmem_ld_bench exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python tools/umma_bench.py > gpurun_out/umma_bench.txt 2>&1
echo "umma_bench exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python tools/micro/h2d_bw.py > gpurun_out/h2d_bw_n1.txt 2>&1
echo "h2d_bw exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
cut -c1-600 gpurun_out/bench.json
grep -h "FAILED\|Error\|error" gpurun_out/test_*.log | head -40
tail -n 30 gpurun_out/tmem_ld_bench.txt
