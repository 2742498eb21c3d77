#!/bin/bash
# N GPUs ($1): the driver's bench command (everything on), then the reference arm under the same launcher
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02ag_bench$N.json 2> gpurun_out/r02ag_$N.err
cut -c1-300 gpurun_out/r02ag_bench$N.json; tail -3 gpurun_out/r02ag_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ag_bench$N.json').read().strip().splitlines()[-1])
g=d['gather']
print(d['n_gpus'], 'ms', d['ms_per_step'], 'Gb/s', d['value']/1e9, 'ingress', g['ingress_GBps'], 'values_only', g['values_only']['ms_per_step'], 'chain', g['chain_only_ms'], 'resident', d['per_gpu_resident']['ms'])
print('e2e', d['e2e']['value']/1e9 if d.get('e2e') else None, 'e2e reduced', d['e2e']['reduced']['value']/1e9 if d.get('e2e') and d['e2e'].get('reduced') else None)
print('reduced', d['reduced']['value']/1e9 if d.get('reduced') else None, d['reduced']['ms_per_step'] if d.get('reduced') else None)
for s in d['secondary']: print(s['config'], round(s['value']/1e9,1), round(s['ms_per_step'],2), s['launches_per_step'], round(s['roofline']['frac'],3))
PY
