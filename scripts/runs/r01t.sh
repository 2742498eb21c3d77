mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -i "numa\|socket\|^CPU(s)" | head -8
for v in "--no-bind" ""; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-gather $v 2> gpurun_out/bench_r01t.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$v', d['n_gpus'], round(d['value']/1e9,1), json.dumps(d['e2e']))
"
tail -2 gpurun_out/bench_r01t.err | cut -c1-300
done
