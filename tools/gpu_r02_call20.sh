#!/bin/bash
# Round-2 call 20 (1 GPU): with the issuer at the pipe's rate the weight ring is what the issuer waits for (17.5 k of 58.6 k
# cycles per tile): re-measure the two flavours that halve the weight bytes per SM against the default, same box.
mkdir -p gpurun_out
{
for rep in 1 2; do
  echo "== default";  timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
  echo "== MONOPORT_B200_TC_CG=2"; MONOPORT_B200_TC_CG=2 timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
  echo "== MONOPORT_B200_TC_WM=1"; MONOPORT_B200_TC_WM=1 timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
done
} 2>&1 | tee gpurun_out/r02c20_weight_halving_ab.txt
MONOPORT_B200_TC_CG=2 timeout 200 python -m pytest tests/test_query_gpu.py -x -q -m gpu --timeout 120 2>&1 | tail -2
