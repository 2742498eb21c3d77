#!/bin/bash
# 1 GPU: skewed staging, bulk copy into the list area + shared-to-shared skewed rewrite (SKEW instantiations):
# r02az): the whole GPU suite, then read length vs Gbases/s
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02bd_pytest.txt 2>&1; tail -3 gpurun_out/r02bd_pytest.txt
OUT=gpurun_out/r02bd_readlen.txt
: > $OUT
for RL in 150 128 256 384; do
  READ_LEN=$RL python scripts/time_c3.py 20000000 11 >> $OUT 2>&1
done
cut -c1-200 $OUT
