#!/bin/bash
# Round-2 call 4 (1 GPU): marching-cubes emission with pipelined lookups; MMA issue-pattern probe.
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --timeout 200 -k "marching or reconstruction or mesh" > gpurun_out/r02c4_pytest_mc.log 2>&1; echo "pytest mc rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02c4_pytest_mc.log
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn | grep -E "mcubes|WordCount|wall|busy" > gpurun_out/r02c4_recon_trace_mc.txt; cat gpurun_out/r02c4_recon_trace_mc.txt
for cfg in "1 0" "1 16" "1 32" "1 64" "1 96" "1 128" "1 256" "1 4" "1 20" "2 8" "2 24" "2 40"; do
  timeout 60 tools/bin/tc_rate $cfg 8192 2>&1 | tail -1
done | tee gpurun_out/r02c4_tc_rate_issue.txt
echo "t=$((SECONDS-T0))s"
