#!/bin/bash
# Round-2 call 17 (1 GPU): tensor-pipe rate probe with the shipped (warp-uniform) issue pattern INSIDE the layer-1 pipeline:
# weights streaming (1), worker warps rewriting the A chunk + hand-offs (2), both (3), A from tensor memory (4), 4x the worker
# store bytes (1024).
mkdir -p gpurun_out
for cfg in "1 512" "1 513" "1 514" "1 515" "1 1539" "1 516" "1 517" "1 531"; do
  timeout 60 tools/bin/tc_rate $cfg 8192 2>&1 | tail -1
done | tee gpurun_out/r02c17_tc_rate_pipeline.txt
