#!/bin/bash
# Round-2 call 18 (1 GPU): which issue-loop structure keeps the tensor pipe at its floor INSIDE the layer-1 pipeline
# (weights streaming = 1, worker hand-offs = 2)?  512 = uniform issue; +2048 every lane polls; +4096 one elected region per
# stage (waits + fence + MMAs + commits); +8192 two stages per region; +16384 no __syncwarp after it; 32768 = the minimal
# control loop of profiles/r02_call5_*.
mkdir -p gpurun_out
for f in 33280 33281 512 513 515 2561 2563 4609 4611 12801 12803 20995 29187; do
  timeout 60 tools/bin/tc_rate 1 $f 8192 2>&1 | tail -1
done | tee gpurun_out/r02c18_tc_rate_issue_loops.txt
