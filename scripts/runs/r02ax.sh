#!/bin/bash
# 1 GPU: positions per chunk of the long-read path (register-window kernels), syncmers on ONT-like reads
mkdir -p gpurun_out
OUT=gpurun_out/r02ax_chunk.txt
: > $OUT
for C in ${C_LIST:-132 152 120 100}; do
  echo "== B200SK_CHUNK_REG=$C" >> $OUT
  B200SK_CHUNK_REG=$C python scripts/run_ont.py syncmer 200000 5 >> $OUT 2>&1
  B200SK_CHUNK_REG=$C python scripts/run_ont.py minimizer 200000 5 >> $OUT 2>&1
done
cut -c1-200 $OUT
