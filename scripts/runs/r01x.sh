mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r01x.json 2> gpurun_out/bench_r01x.err
cut -c1-400 gpurun_out/bench_r01x.json; python -c "
import json
d=json.loads(open('gpurun_out/bench_r01x.json').read().strip().splitlines()[-1])
print(d['steps'], d['warmup'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('host_binding'), d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
"
timeout 120 python bench.py --impl reference > gpurun_out/bench_ref_r01x.json 2>> gpurun_out/bench_r01x.err
cut -c1-250 gpurun_out/bench_ref_r01x.json
