#!/bin/bash
# Round 2, 8-GPU call: the host's pinned H2D ceiling with 1..8 ranks copying at once, NUMA placement variants.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -30 >> gpurun_out/topo.txt 2>&1
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/micro/h2d_bw.py > gpurun_out/h2d_bw_n$N.txt 2> gpurun_out/h2d_bw_n$N.err
echo "h2d exit $?"
cat gpurun_out/h2d_bw_n$N.txt
tail -n 5 gpurun_out/h2d_bw_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 6 --warmup 3 --no-side --pool 16640 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"
python - <<'PY'
import json,sys,glob
for f in glob.glob('gpurun_out/bench_n*.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['e2e']['value'], d['e2e'].get('h2d_ceiling_gbs'), d['e2e'].get('numa'), d.get('parity',{}).get('sharded_equals_single'))
    except Exception as e: print(f, 'ERR', e)
PY
