#!/bin/bash
# 2 GPUs: the whole GPU suite (multi-GPU tests included), then the strong-scaling step with line-aligned peer stores
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02ab_pytest.txt 2>&1; tail -3 gpurun_out/r02ab_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary --no-e2e --no-cpu --no-reduce --parity-reads 20000 > gpurun_out/r02ab_bench2.json 2> gpurun_out/r02ab.err
cut -c1-300 gpurun_out/r02ab_bench2.json; tail -3 gpurun_out/r02ab.err
