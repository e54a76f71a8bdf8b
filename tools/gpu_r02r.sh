#!/bin/bash
# Round 2: validation of the committed state: full GPU suite, smoke, default bench, reference arm (short)
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/test_all.log 2>&1
echo "test_all exit $? $(tail -n 1 gpurun_out/test_all.log)" | tee -a gpurun_out/summary.txt
grep -h "FAILED\|Error" gpurun_out/test_all.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $? $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'steps',d['steps'],'ms',d['ms_per_step'],'clocks',d['clocks'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','frac_of_burst_peak','kernel_ms_share','conv12_fused_tflops','whole_path_frac_of_burst_peak')})
print('e2e',d['e2e']['value'],d['e2e']['frac_of_h2d_ceiling'],d['e2e'].get('numa'))
for k,v in d['workloads'].items(): print(k,v['value'],v.get('e2e',{}).get('value'),v['clocks']['sm_mhz'],v['clocks']['samples'])
print('parity split',d['parity']['split_mode'])
PY
