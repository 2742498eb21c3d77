mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_fastx.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python scripts/bench_fastx.py 8000000 > gpurun_out/r01z_feeder.jsonl 2> gpurun_out/r01z_feeder.err
cut -c1-200 gpurun_out/r01z_feeder.jsonl; tail -3 gpurun_out/r01z_feeder.err
