#!/bin/bash
# 1 GPU: compute-sanitizer over the final library -- memcheck on the whole GPU suite, racecheck on the kernels that stage
# through shared memory and the async proxy (SHARD bulk-store flush, six-frame kernel, feeder)
mkdir -p gpurun_out
OUT=gpurun_out/r02ai_sanitizer.txt
: > $OUT
echo "== compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q" >> $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q > gpurun_out/r02ai_memcheck.log 2>&1
echo "rc=$?" >> $OUT
grep -E "ERROR SUMMARY|passed|failed|skipped" gpurun_out/r02ai_memcheck.log | tail -4 >> $OUT
echo "== compute-sanitizer --tool racecheck: single-rank SHARD chain, six-frame kernel, two streams, fastx" >> $OUT
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -x -q -m gpu tests/test_multi_gpu.py tests/test_parity_gpu.py tests/test_fastx.py \
   -k "single_rank or six_frames or two_streams or fxstream_equals or committed_fixture" > gpurun_out/r02ai_racecheck.log 2>&1
echo "rc=$?" >> $OUT
grep -E "RACECHECK SUMMARY|passed|failed|skipped" gpurun_out/r02ai_racecheck.log | tail -4 >> $OUT
cat $OUT
