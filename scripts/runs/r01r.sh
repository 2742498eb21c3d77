mkdir -p gpurun_out
cd scripts && timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ../gpurun_out/fx_launches_r01r.csv python fx_kernels.py 2>&1 | tail -3
