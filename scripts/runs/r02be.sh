#!/bin/bash
# 1 GPU: skewed staging for every read length that is a multiple of 32 bytes: the whole GPU suite, then read length vs
# Gbases/s (160: 4 banks before, 192: 2 banks)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02be_pytest.txt 2>&1; tail -3 gpurun_out/r02be_pytest.txt
OUT=gpurun_out/r02be_readlen.txt
: > $OUT
for RL in 160 192 128 96; do
  READ_LEN=$RL python scripts/time_c3.py 20000000 11 >> $OUT 2>&1
done
cut -c1-200 $OUT
