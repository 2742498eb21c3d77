#!/bin/bash
# N GPUs ($1): the strong-scaling step only
mkdir -p gpurun_out
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-e2e --no-cpu --no-reduce --parity-reads 20000 > gpurun_out/r02x_bench$N.json 2> gpurun_out/r02x_$N.err
cut -c1-300 gpurun_out/r02x_bench$N.json; tail -3 gpurun_out/r02x_$N.err
