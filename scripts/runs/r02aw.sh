#!/bin/bash
# 1 GPU: A/B -- status store next to the ticket (base) vs with the offsets after the resolve (product)
mkdir -p gpurun_out
OUT=gpurun_out/r02aw_ab.txt
: > $OUT
for V in base product base product; do
  if [ $V = product ]; then unset B200SK_LIB_PATH; else export B200SK_LIB_PATH=$PWD/bio_b200/lib/ab/libb200sketch_$V.so; fi
  echo "== $V" >> $OUT
  python scripts/time_c3.py 40000000 11 >> $OUT 2>&1; python scripts/run_ont.py syncmer 200000 5 >> $OUT 2>&1
done
cut -c1-200 $OUT
