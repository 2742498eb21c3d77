#!/bin/bash
# 1 GPU: A/B -- HEAD library vs next-tile ticket/geometry prefetch vs prefetch + 12 warps for wide minimizer windows;
# then the sparse-kernel parity tests with the product library
mkdir -p gpurun_out
OUT=gpurun_out/r02ao_ab.txt
: > $OUT
for V in head prefetch product; do
  if [ $V = product ]; then unset B200SK_LIB_PATH; else export B200SK_LIB_PATH=$PWD/bio_b200/lib/ab/libb200sketch_$V.so; fi
  echo "== $V" >> $OUT
  python scripts/time_c3.py 40000000 11 >> $OUT 2>&1
  python scripts/time_c3.py 20000000 24 >> $OUT 2>&1
  python scripts/time_c3.py 20000000 19 >> $OUT 2>&1
  python scripts/run_ont.py syncmer 200000 5 >> $OUT 2>&1
  python scripts/run_ont.py minimizer 200000 5 >> $OUT 2>&1
  python scripts/run_mode.py syncmer 5 >> $OUT 2>&1
  python scripts/run_mode.py protmin 5 >> $OUT 2>&1
done
unset B200SK_LIB_PATH
python -m pytest tests/test_parity_gpu.py tests/test_sketches_api_gpu.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -2 >> $OUT
cut -c1-200 $OUT
