#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eer.py -m gpu -q --tb=short -x > gpurun_out/test_eer.log 2>&1
echo "eer tests exit $? $(tail -n 1 gpurun_out/test_eer.log)"
grep -h "FAILED\|Error" gpurun_out/test_eer.log | head
timeout 300 python bench.py --workload eer --leg-seconds 1.0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('sort ms', d['ms_per_step'], 'select ms', d['eer_select']['ms_per_step'], {k:(v['sort_ms'],v['select_ms']) for k,v in d['inputs'].items()})"
