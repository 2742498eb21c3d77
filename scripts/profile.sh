#!/bin/bash
# Runs on the GPU box (under gpurun): bench + ncu launch list + one full capture of the sketching kernel at the
# bench's own launch size.  $1 = tag, $2 = reads per GPU (default 100M = C3)
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
READS=${2:-100000000}
python bench.py --steps 10 --warmup 3 --reads $READS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json | cut -c1-1800
tail -5 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-600 gpurun_out/bench_ref_$TAG.json
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --reads $READS --no-e2e --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log | cut -c1-300
# full capture of the top kernel
ncu --set full --clock-control none --import-source on -k regex:k_sparse_warp -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --reads $READS --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
ls -la gpurun_out | tail -8
