#!/bin/bash
# 1 GPU: A/B of the look-back variants (scripts/ab_build.sh libraries) on C3 geometry and on ONT-like syncmers
mkdir -p gpurun_out
OUT=gpurun_out/r02z_ab.txt
: > $OUT
for V in "$@"; do
  export B200SK_LIB_PATH=$PWD/bio_b200/lib/ab/libb200sketch_$V.so
  echo "== $V" >> $OUT
  python scripts/time_c3.py 40000000 11 >> $OUT 2>&1
  python scripts/run_ont.py syncmer 200000 5 >> $OUT 2>&1
  python scripts/run_ont.py minimizer 200000 5 >> $OUT 2>&1
done
cat $OUT
