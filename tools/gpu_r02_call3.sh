#!/bin/bash
# Round-2 call 3 (1 GPU): bit-volume marching cubes (parity + per-kernel times), long rate probe (does cycles/MMA fall when the
# power cap lowers the clock?), N = 128 MMAs.
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --timeout 200 > gpurun_out/r02c3_pytest_engine.log 2>&1; echo "pytest engine rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02c3_pytest_engine.log
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/r02c3_recon_trace_mc.txt; head -30 gpurun_out/r02c3_recon_trace_mc.txt
for cfg in "1 0 400000" "1 11 400000" "2 8 400000"; do
  ( nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 20 > gpurun_out/r02c3_clk.txt & P=$!; timeout 60 tools/bin/tc_rate $cfg 2>&1 | tail -1; kill $P ) 
  sort gpurun_out/r02c3_clk.txt | uniq -c | sort -rn | head -3
done | tee gpurun_out/r02c3_tc_rate_long.txt
echo "t=$((SECONDS-T0))s"
