#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 512" "1 528" "1 513" "1 521" "2 520" "2 521" "1 64"; do
  timeout 60 tools/bin/tc_rate $cfg 8192 2>&1 | tail -1
done | tee gpurun_out/r02c5_tc_rate_uniform.txt
