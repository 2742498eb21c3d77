set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_fastx.py -m gpu -x -q 2>&1 | tail -15
timeout 500 python scripts/bench_fastx.py 8000000 > gpurun_out/r01p_feeder.jsonl 2> gpurun_out/r01p_feeder.err
cut -c1-400 gpurun_out/r01p_feeder.jsonl; tail -5 gpurun_out/r01p_feeder.err
