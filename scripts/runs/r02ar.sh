#!/bin/bash
# N GPUs ($1): chunk size of the sharded chain (reads per chunk; chunk c belongs to rank c mod N)
mkdir -p gpurun_out
N=${1:-2}
for C in 4096 32768 262144; do
B200SK_BENCH_CHUNK=$C python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary --no-e2e --no-cpu --no-reduce --parity-reads 2000 > gpurun_out/r02ar_bench${N}_$C.json 2> gpurun_out/r02ar_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ar_bench${N}_$C.json').read().strip().splitlines()[-1])
print("chunk", $C, "N", d['n_gpus'], d['ms_per_step'], d['gather']['ingress_GBps'], d['gather']['values_only']['ms_per_step'], d['gather']['chain_only_ms'], d['gather']['gathered_checksum_ok'])
PY
done
