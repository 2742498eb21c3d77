#!/bin/bash
# 1 GPU: full ncu capture of the long-read syncmer kernel after the 12-warp change
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_sparse_warp -s 3 -c 1 -f -o gpurun_out/prof_r02aj_c4 \
    python scripts/run_ont.py syncmer 200000 3 > gpurun_out/ncu_full_r02aj_c4.log 2>&1
tail -2 gpurun_out/ncu_full_r02aj_c4.log | cut -c1-200
