#!/bin/bash
# Runs on the GPU box (under gpurun, 1 GPU): the driver's bench command, the reference arm, the ncu launch list of the
# bench command, and full captures of the sketching kernel (at the bench's launch size) and of the f4 kernels.
# $1 = tag
set -x
mkdir -p gpurun_out
TAG=${1:-r02}
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-400 gpurun_out/bench_ref_$TAG.json
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --parity-reads 20000 > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log | cut -c1-300
# full capture of the top kernel at the bench's launch size (100 M reads): its DRAM traffic goes to profiles/traffic.json
ncu --set full --clock-control none --import-source on -k regex:k_sparse_warp -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --no-reduce --parity-reads 2000 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
# the f4 kernels (filter, radix passes, unique) on a 20 M-read stream
ncu --set full --clock-control none --import-source on -k regex:"k_radix|k_filter|k_unique|k_scan_hist" -c 12 -f -o gpurun_out/prof_${TAG}_reduce \
    python bench.py --reads 20000000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --parity-reads 2000 > gpurun_out/ncu_full_${TAG}_reduce.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}_reduce.log | cut -c1-300
# the fused six-frame protein kernel (C5) and the long-read syncmer kernel (C4) at the bench's sizes
ncu --set full --clock-control none --import-source on -k regex:"k_protein6_warp" -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_c5 \
    python bench.py --reads 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-reduce --parity-reads 2000 > gpurun_out/ncu_full_${TAG}_c5.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}_c5.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:"k_sparse_warp<\(int\)3" -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_c4 \
    python bench.py --reads 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-reduce --parity-reads 2000 > gpurun_out/ncu_full_${TAG}_c4.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}_c4.log | cut -c1-300
ls -la gpurun_out | tail -8
