mkdir -p gpurun_out
( echo 'compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -x -q -m gpu'
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
echo 'compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_fastx.py -x -q -m gpu'
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_fastx.py -x -q -m gpu 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8 ) > gpurun_out/r01u_sanitizer.txt 2>&1
cat gpurun_out/r01u_sanitizer.txt
