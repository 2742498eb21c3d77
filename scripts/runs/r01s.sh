mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | head -8
free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r01s_${N}gpu.json 2> gpurun_out/bench_r01s_${N}gpu.err
cut -c1-1500 gpurun_out/bench_r01s_${N}gpu.json; tail -5 gpurun_out/bench_r01s_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_r01s_${N}gpu.json 2>> gpurun_out/bench_r01s_${N}gpu.err
cut -c1-300 gpurun_out/bench_ref_r01s_${N}gpu.json
