#!/bin/bash
# Round 2: fused conv1 + conv2 (conflict-free window stores, 112 registers): parity + rate + default bench
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_models.py -m gpu -q --tb=short -x > gpurun_out/test_fused.log 2>&1
echo "tests exit $? $(tail -n 1 gpurun_out/test_fused.log)" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_*.log | head -20
timeout 300 python tools/split_rate.py 2>&1 | head -1 | tee gpurun_out/split_rate.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'clocks',d['clocks'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms','kernel_ms_share','whole_path_tflops','whole_path_frac_of_burst_peak')})
print('e2e',d['e2e']['value'],d['e2e']['frac_of_h2d_ceiling'],'f16slab',d['e2e_f16_slab']['value'])
for k,v in d['workloads'].items(): print(k,v['value'],v.get('e2e',{}).get('value'),v['clocks']['sm_mhz'], v.get('eer_select',{}).get('ms_per_step'))
print('parity',{k:v for k,v in d['parity'].items() if k not in('note','labels')})
PY
