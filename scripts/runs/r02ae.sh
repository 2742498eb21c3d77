#!/bin/bash
# 1 GPU: warps per SM of the ordered sparse kernels -- is a multiple of 4 (balanced schedulers) better than the most that fit?
mkdir -p gpurun_out
OUT=gpurun_out/r02ae_warps.txt
: > $OUT
for NW in 0 12 8; do
  echo "== B200SK_MAX_WARPS=$NW" >> $OUT
  B200SK_MAX_WARPS=$NW python scripts/run_ont.py syncmer 200000 5 >> $OUT 2>&1
  B200SK_MAX_WARPS=$NW python scripts/run_ont.py minimizer 200000 5 >> $OUT 2>&1
  B200SK_MAX_WARPS=$NW python scripts/run_mode.py syncmer 5; B200SK_MAX_WARPS=$NW python scripts/run_mode.py protmin 5 >> $OUT 2>&1
done
cat $OUT
