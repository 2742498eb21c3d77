#!/bin/bash
# 1 GPU: SimHash with signed bit-sliced sums (no threshold compare): parity + timing
mkdir -p gpurun_out
OUT=gpurun_out/r02ak_simhash.txt
: > $OUT
python -m pytest tests -x -q -m gpu -k "simhash or SimHash" 2>&1 | tail -2 >> $OUT
python scripts/run_mode.py simhash 5 >> $OUT 2>&1
cat $OUT
