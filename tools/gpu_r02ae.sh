#!/bin/bash
# Round 2, session 3, full validation: every GPU test, smoke(), the default bench line (N = 1) and the reference arm's short run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/test_all.log 2>&1
echo "gpu tests exit $? $(tail -n 1 gpurun_out/test_all.log)"
grep -h "FAILED\|Error" gpurun_out/test_all.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke exit $? $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline'].get('frac_of_burst_peak'), 'clocks', d['clocks'])
for k, v in d.get('workloads', {}).items():
    print(k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value'), v.get('clocks', {}).get('samples'), v.get('roofline', {}).get('frac'))
print(json.dumps(d['workloads']['eer'].get('inputs'), indent=0))
print(d.get('cpu_baseline'))
PY
