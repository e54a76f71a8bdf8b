#!/bin/bash
# usage: tools/sweep.sh "<bench args A>" "<bench args B>" ...   -> one summary line per configuration
for a in "$@"; do
  python bench.py --no-cpu-baseline --steps 20 --warmup 3 $a 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ARGS', sys.argv[1], '| utt/s %.0f' % d['value'], '| clk', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w_max'], '| conv3 TF %.0f' % r['achieved'], 'conv2 TF %.0f' % (r['conv2_tflops'] or 0), '| share', {k: round(v, 3) for k, v in r['kernel_ms_share'].items()}, '| e2e %.0f' % d['e2e']['value'])
" "$a"
done
