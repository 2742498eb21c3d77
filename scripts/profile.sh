#!/bin/bash
# Runs on the GPU box (under gpurun): bench + ncu launch list + one full capture of the sketching kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
READS=${2:-100000000}
python bench.py --steps 10 --warmup 3 --reads $READS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --reads 20000000 --no-e2e --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -3 gpurun_out/ncu_launches_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:k_sparse_warp -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --reads 10000000 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
