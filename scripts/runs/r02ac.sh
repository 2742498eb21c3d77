#!/bin/bash
# N GPUs ($1): sweep of the chain's poll policy (free polls before the back-off)
mkdir -p gpurun_out
N=${1:-2}
for FREE in 0 4 16; do
B200SK_CHAIN_POLL_FREE=$FREE python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary --no-e2e --no-cpu --no-reduce --parity-reads 2000 > gpurun_out/r02ac_bench${N}_$FREE.json 2> gpurun_out/r02ac_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ac_bench${N}_$FREE.json').read().strip().splitlines()[-1])
print("free", $FREE, d['n_gpus'], d['ms_per_step'], d['gather']['ingress_GBps'], d['gather']['values_only']['ms_per_step'], d['gather']['chain_only_ms'])
PY
done
