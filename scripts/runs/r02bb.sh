#!/bin/bash
# 1 GPU: skewed staging of uniform reads of n x 128 bytes (k_sparse_warp, stage_skewed) -- the whole GPU suite with the
# new parity tests, smoke(), then read length vs Gbases/s: new library, and the library before the change at 150 / 128 bp
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02bb_pytest.txt 2>&1; tail -3 gpurun_out/r02bb_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02bb_smoke.txt 2>&1; tail -1 gpurun_out/r02bb_smoke.txt | cut -c1-300
OUT=gpurun_out/r02bb_readlen.txt
: > $OUT
for RL in 150 128 256 384; do
  echo "new RL=$RL" >> $OUT
  READ_LEN=$RL python scripts/time_c3.py 20000000 11 >> $OUT 2>&1
done
for RL in 150 128; do
  echo "before RL=$RL" >> $OUT
  B200SK_LIB_PATH=bio_b200/lib/ab/libb200sketch_r02az.so READ_LEN=$RL python scripts/time_c3.py 20000000 11 >> $OUT 2>&1
done
cut -c1-200 $OUT
