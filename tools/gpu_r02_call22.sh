#!/bin/bash
# Round-2 call 22 (1 GPU): coarse-to-fine frames with / without weight multicast in the level launches (their capacity, not their
# device-side count, is what the host sees), same box.
mkdir -p gpurun_out
{
for rep in 1 2; do
  for wm in auto 0; do
    echo "== MONOPORT_B200_TC_WM=$wm"
    if [ $wm = auto ]; then unset MONOPORT_B200_TC_WM; else export MONOPORT_B200_TC_WM=$wm; fi
    timeout 120 python tools/recon_trace.py 2>&1 | grep -E "wall per frame|GPU busy|query_tc3_kernel" | head -5 | cut -c1-120
  done
done
unset MONOPORT_B200_TC_WM
} 2>&1 | tee gpurun_out/r02c22_frames_wm_ab.txt
