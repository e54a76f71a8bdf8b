#!/bin/bash
# Round 2, final 8-GPU call: scaling curve 1 / 2 / 4 / 8 of the headline leg (device-resident and e2e with the measured host ceiling per N)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --pool 16640 --no-side --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 8 --warmup 3 --pool 16640 --no-side > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N exit $?"
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f'gpurun_out/scale_n{n}.json').read().strip().splitlines()[-1])
        e = d['e2e']
        print(f"N={n}: device-resident {d['value']:.0f} utt/s, e2e {e['value']:.0f} utt/s, host H2D ceiling {e['h2d_ceiling_gbs']:.1f} GB/s, e2e / ceiling {e['frac_of_h2d_ceiling']:.3f}, sharded_equals_single {d.get('parity', {}).get('sharded_equals_single')}, clocks {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")
    except Exception as ex:
        print(n, 'ERR', ex)
PY
