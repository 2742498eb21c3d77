#!/bin/bash
# 2 GPUs: multi-GPU tests (bulk-store flush of the SHARD kernels), then the strong-scaling step
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02w_pytest.txt 2>&1; tail -15 gpurun_out/r02w_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary --no-e2e --no-cpu --no-reduce --parity-reads 20000 > gpurun_out/r02w_bench2.json 2> gpurun_out/r02w.err
cut -c1-300 gpurun_out/r02w_bench2.json; tail -3 gpurun_out/r02w.err
