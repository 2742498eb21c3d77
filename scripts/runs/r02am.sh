#!/bin/bash
# N GPUs ($1): whole-tile staging buffer of the sharded chain on (default for >= 3 ranks) / off
mkdir -p gpurun_out
N=${1:-4}
for R in 3 99; do
B200SK_WHOLE_TILE_RANKS=$R python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary --no-e2e --no-cpu --no-reduce --parity-reads 2000 > gpurun_out/r02am_bench${N}_$R.json 2> gpurun_out/r02am_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02am_bench${N}_$R.json').read().strip().splitlines()[-1])
print("whole-tile from", $R, "ranks:", d['n_gpus'], d['ms_per_step'], d['gather']['ingress_GBps'], d['gather']['values_only']['ms_per_step'], d['gather']['chain_only_ms'], d['gather']['gathered_checksum_ok'])
PY
done
