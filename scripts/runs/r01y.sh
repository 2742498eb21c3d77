mkdir -p gpurun_out
cd scripts && timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_fx_lines$|k_fq_scan" -s 2 -c 2 -f -o ../gpurun_out/prof_r01y_feeder python fx_kernels.py 2>&1 | tail -3
