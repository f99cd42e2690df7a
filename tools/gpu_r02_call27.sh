#!/bin/bash
# Round-2 call 27 (1 GPU): 64-channel chunks + four-stage weight ring for latency-bound launches (coarse-to-fine levels) --
# parity, then frames with the new default against MONOPORT_B200_TC_KB=2 (128-channel chunks everywhere) on the same box.
mkdir -p gpurun_out
T0=$SECONDS
timeout 400 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02c27_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02c27_pytest.log
{
for rep in 1 2; do
  for kb in auto 2; do
    echo "== MONOPORT_B200_TC_KB=$kb"
    if [ $kb = auto ]; then unset MONOPORT_B200_TC_KB; else export MONOPORT_B200_TC_KB=$kb; fi
    timeout 120 python tools/recon_trace.py 2>&1 | grep -E "wall per frame|GPU busy|query_tc3_kernel" | head -6 | cut -c1-120
  done
done
unset MONOPORT_B200_TC_KB
echo "== dense 257^3, default"; timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
echo "== dense 257^3, MONOPORT_B200_TC_KB=1 MONOPORT_B200_TC_WM=0"; MONOPORT_B200_TC_KB=1 MONOPORT_B200_TC_WM=0 timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
echo "== dense 257^3, MONOPORT_B200_TC_KB=2 MONOPORT_B200_TC_WM=0"; MONOPORT_B200_TC_KB=2 MONOPORT_B200_TC_WM=0 timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
} 2>&1 | tee gpurun_out/r02c27_chunk_depth_ab.txt
