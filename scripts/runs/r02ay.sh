#!/bin/bash
# 1 GPU: chunk size chosen per plan -- parity of the whole suite's long-read cases, timing on ONT-like reads
mkdir -p gpurun_out
OUT=gpurun_out/r02ay_chunk_plan.txt
: > $OUT
python -m pytest tests/test_parity_gpu.py tests/test_sketches_api_gpu.py tests/test_reduce_gpu.py -x -q -m gpu 2>&1 | tail -2 >> $OUT
python scripts/run_ont.py syncmer 200000 5 >> $OUT 2>&1
python scripts/run_ont.py minimizer 200000 5 >> $OUT 2>&1
READS=2000000 python scripts/run_mode.py protmin 5 >> $OUT 2>&1
cut -c1-200 $OUT
