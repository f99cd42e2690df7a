#!/bin/bash
# Round-2 call 14 (1 GPU): sliced first-hit kernel + list-driven marching-cubes emission (parity + timings), then the A/B of
# the weight-multicast flavour of the geometry program (MONOPORT_B200_TC_WM=1) against the default on the same box.
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --timeout 200 > gpurun_out/r02c14_pytest_engine.log 2>&1; echo "pytest engine rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02c14_pytest_engine.log
timeout 200 python -m pytest tests/test_query_gpu.py -x -q -m gpu -k 'host or full_size' --timeout 150 2>&1 | tail -3
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/r02c14_recon_trace_fv_mc.txt; grep -E "first_hit|mesh_emit|bits_kernel|classify|block_sums|emit_kernel|per frame" gpurun_out/r02c14_recon_trace_fv_mc.txt | head -14
{
for rep in 1 2; do
  for wm in 0 1; do
    echo "== MONOPORT_B200_TC_WM=$wm"
    MONOPORT_B200_TC_WM=$wm timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms|Mpts" | tail -2
  done
done
echo "== MONOPORT_B200_TC_WM=1 in-kernel attribution"
MONOPORT_B200_TC_WM=1 MONOPORT_B200_TC_PROF=1 timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "tc prof\]" | head -24
} 2>&1 | tee gpurun_out/r02c14_wm_ab.txt
MONOPORT_B200_TC_WM=1 timeout -k 5 300 python -m pytest tests/test_query_gpu.py -x -q -m gpu --timeout 120 > gpurun_out/r02c14_pytest_query_wm.log 2>&1; echo "pytest query (WM=1) rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02c14_pytest_query_wm.log
