#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?" >> gpurun_out/tc_check.log
cat gpurun_out/tc_check.log
if grep -q "dense 257" gpurun_out/tc_check.log; then
  timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"
  cat gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
fi
