#!/bin/bash
# 1 GPU: A/B of the SimHash counters (old unsigned counts + threshold compare vs signed sums)
mkdir -p gpurun_out
OUT=gpurun_out/r02al_simhash_ab.txt
: > $OUT
for i in 1 2; do
echo "== signed sums (product)" >> $OUT
python scripts/run_mode.py simhash 5 >> $OUT 2>&1
echo "== unsigned counts + compare (round 1)" >> $OUT
B200SK_LIB_PATH=$PWD/bio_b200/lib/ab/libb200sketch_simold.so python scripts/run_mode.py simhash 5 >> $OUT 2>&1
done
cat $OUT
