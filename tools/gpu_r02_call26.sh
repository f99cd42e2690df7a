#!/bin/bash
# Round-2 call 26 (1 GPU): compute-sanitizer memcheck with the weight-multicast cluster launch forced for every tensor-core
# geometry query; in-kernel attribution of coarse-to-fine sized launches with the final issuer (what a lone tile costs now).
mkdir -p gpurun_out
T0=$SECONDS
MONOPORT_B200_TC_WM=1 MONOPORT_B200_RUN_SANITIZER=1 MONOPORT_B200_SANITIZER_LOG=gpurun_out/r02c26_sanitizer_memcheck_wm.log timeout 900 python -m pytest tests/test_sanitizer_gpu.py -q -m gpu 2>&1 | tail -2; echo "sanitizer (WM=1) t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02c26_sanitizer_memcheck_wm.log
{
for r in 17 21 33; do
  echo "== grid $r^3 = $((r*r*r)) points = $(( (r*r*r+127)/128 )) tiles"
  MONOPORT_B200_TC_PROF=2 timeout 120 python tools/tc_prof.py $r 2>&1 | grep -E "tc prof\]" | head -21
done
} 2>&1 | tee gpurun_out/r02c26_small_launch_attribution.txt | grep -E "==|total|wfull|h0ready|ph_L1hid|w_drain0|h1ready"
