#!/bin/bash
# 2 GPUs: final check of the round's library -- the whole GPU suite, smoke(), the driver's bench command on one GPU + reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02az_pytest.txt 2>&1; tail -3 gpurun_out/r02az_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02az_smoke.txt 2>&1; tail -2 gpurun_out/r02az_smoke.txt | cut -c1-400
CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02az_bench.json 2> gpurun_out/r02az_bench.err
tail -2 gpurun_out/r02az_bench.err
CUDA_VISIBLE_DEVICES=0 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02az_bench_ref.json 2>> gpurun_out/r02az_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02az_bench.json').read().strip().splitlines()[-1])
print('value',d['value']/1e9,'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value']/1e9, 'reduced e2e', d['e2e']['reduced']['value']/1e9,'launches',d['gpu_launches'], 'parity', d['parity']['all_ok'])
for s in d['secondary']: print(s['config'], round(s['value']/1e9,1), round(s['ms_per_step'],2), s['launches_per_step'], round(s['roofline']['frac'],3))
r=json.loads(open('gpurun_out/r02az_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value']/1e6, r['cpu_baseline']['cores'])
PY
