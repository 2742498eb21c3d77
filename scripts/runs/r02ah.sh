#!/bin/bash
# 1 GPU: protein-minimizer parity + timing with the branch-free run-time-k wyhash; compute-sanitizer over this round's new kernels
mkdir -p gpurun_out
OUT=gpurun_out/r02ah.txt
: > $OUT
python -m pytest tests/test_parity_gpu.py tests/test_sketches_api_gpu.py -x -q -m gpu -k "protein" 2>&1 | tail -2 >> $OUT
python scripts/run_mode.py protmin 5 >> $OUT 2>&1
python scripts/run_mode.py protein 5 >> $OUT 2>&1
echo "== compute-sanitizer --tool memcheck: six-frame kernel, SHARD bulk-store flush (single rank), two streams, protein minimizers, 12-warp syncmers" >> $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q -m gpu \
   "tests/test_parity_gpu.py::test_protein_six_frames_one_call" "tests/test_parity_gpu.py::test_protein_six_frames_host_entry" \
   "tests/test_parity_gpu.py::test_two_streams_one_context" "tests/test_multi_gpu.py::test_sharded_chain_single_rank_bulk_stores" \
   tests/test_parity_gpu.py -k "six_frames or two_streams or single_rank or protein_minimizer or syncmer" > gpurun_out/r02ah_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02ah_memcheck.log | tail -5 >> $OUT
cat $OUT
