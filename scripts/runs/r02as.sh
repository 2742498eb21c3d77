#!/bin/bash
# 1 GPU: feeder tests (incl. the long blank-line run) + mirror tests
mkdir -p gpurun_out
python -m pytest tests/test_fastx.py tests/test_sketches_api_gpu.py -x -q -m gpu > gpurun_out/r02as_pytest.txt 2>&1; tail -5 gpurun_out/r02as_pytest.txt
