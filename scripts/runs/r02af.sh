#!/bin/bash
# 1 GPU: syncmer kernels at 12 warps / 166 registers (no spills); warps-per-SM sweep of protmin and C3
mkdir -p gpurun_out
OUT=gpurun_out/r02af_warps.txt
: > $OUT
python scripts/run_ont.py syncmer 200000 5 >> $OUT 2>&1
python scripts/run_mode.py syncmer 5 >> $OUT 2>&1
for NW in 16 14 12 10; do
  echo "== B200SK_MAX_WARPS=$NW" >> $OUT
  B200SK_MAX_WARPS=$NW python scripts/run_mode.py protmin 5 >> $OUT 2>&1
  B200SK_MAX_WARPS=$NW python scripts/time_c3.py 20000000 11 >> $OUT 2>&1
done
python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "syncmer" 2>&1 | tail -2 >> $OUT
cat $OUT
