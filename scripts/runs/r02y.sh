#!/bin/bash
# N GPUs ($1): sweep of the chain's poll back-off
mkdir -p gpurun_out
N=${1:-4}
for NS in ${NS_LIST:-100 1000 2500}; do
B200SK_CHAIN_POLL_NS=$NS python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary --no-e2e --no-cpu --no-reduce --parity-reads 2000 > gpurun_out/r02y_bench${N}_$NS.json 2> gpurun_out/r02y_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02y_bench${N}_$NS.json').read().strip().splitlines()[-1])
print($NS, d['n_gpus'], d['ms_per_step'], d['gather']['ingress_GBps'], d['gather']['values_only']['ms_per_step'], d['gather']['chain_only_ms'])
PY
done
