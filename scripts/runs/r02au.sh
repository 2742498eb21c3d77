#!/bin/bash
# 1 GPU: A/B -- suffix-step selects on the FMA pipe (IMAD pairs) vs SEL
mkdir -p gpurun_out
OUT=gpurun_out/r02au_ab.txt
: > $OUT
for V in base fmasel base fmasel; do
  export B200SK_LIB_PATH=$PWD/bio_b200/lib/ab/libb200sketch_$V.so
  echo "== $V" >> $OUT
  python scripts/time_c3.py 40000000 11 >> $OUT 2>&1
done
cut -c1-200 $OUT
