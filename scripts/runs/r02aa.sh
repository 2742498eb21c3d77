#!/bin/bash
# 1 GPU: six-frame fused kernel tests + the secondary configs of the bench on a reduced C3 batch
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "six_frames or protein" > gpurun_out/r02aa_pytest.txt 2>&1; tail -15 gpurun_out/r02aa_pytest.txt
python bench.py --reads 20000000 --steps 6 --warmup 3 --no-e2e --no-cpu --no-reduce --parity-reads 20000 > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa.err
tail -3 gpurun_out/r02aa.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02aa_bench.json').read().strip().splitlines()[-1])
for s in d['secondary']: print(s['config'], round(s['value']/1e9,1), s['ms_per_step'], s['launches_per_step'], round(s['roofline']['frac'],3))
print(d['parity'])
PY
