#!/bin/bash
# 1 GPU: read length vs shared-memory bank conflicts of the lane-per-read walk (lanes start read_len bytes apart)
mkdir -p gpurun_out
OUT=gpurun_out/r02ba_readlen.txt
: > $OUT
for RL in 150 128 132 160 100 256 250; do
  READ_LEN=$RL python scripts/time_c3.py 20000000 11 >> $OUT 2>&1
done
cut -c1-220 $OUT
