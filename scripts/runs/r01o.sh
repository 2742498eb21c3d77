set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01o.json 2> gpurun_out/bench_r01o.err
cut -c1-900 gpurun_out/bench_r01o.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01o.json 2>> gpurun_out/bench_r01o.err
cut -c1-300 gpurun_out/bench_ref_r01o.json
timeout 400 python scripts/bench_modes.py > gpurun_out/r01o_modes.jsonl 2>> gpurun_out/bench_r01o.err
cut -c1-260 gpurun_out/r01o_modes.jsonl
timeout 200 bash scripts/profile_only.sh r01o_c2 k_nthash_warp python scripts/run_mode.py nthash 2
