#!/bin/bash
# 1 GPU: A/B -- committed library vs speculative L2 warm-up of the next tile
mkdir -p gpurun_out
OUT=gpurun_out/r02aq_ab.txt
: > $OUT
for V in head product head product; do
  if [ $V = product ]; then unset B200SK_LIB_PATH; else export B200SK_LIB_PATH=$PWD/bio_b200/lib/ab/libb200sketch_$V.so; fi
  echo "== $V" >> $OUT
  python scripts/time_c3.py 40000000 11 >> $OUT 2>&1
  python scripts/run_mode.py syncmer 5 >> $OUT 2>&1
done
unset B200SK_LIB_PATH
python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "minimizer or syncmer or fixture or two_streams" 2>&1 | tail -2 >> $OUT
cut -c1-200 $OUT
