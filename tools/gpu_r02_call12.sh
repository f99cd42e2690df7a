#!/bin/bash
# Round-2 call 12 (1 GPU): what-if timings of the geometry program (results are WRONG with these switches: timing only) --
# 1 = weights not re-streamed after the first ring fill, 2 = all layer-0 gathers hit texel 0 (L1), 4 = all X taps hit texel 0
mkdir -p gpurun_out
for e in 0 1 2 4 3 7; do
  echo "== MONOPORT_B200_TC_EXP=$e"
  MONOPORT_B200_TC_EXP=$e MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep -E "tc prof\] (total|wfull|h0ready|h1ready|ph_L1hid|w_h0free|w_drain0)" | head -7
  MONOPORT_B200_TC_EXP=$e timeout 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms|Mpts" | tail -2
done 2>&1 | tee gpurun_out/r02c12_whatif.txt
