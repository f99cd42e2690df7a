#!/bin/bash
# Round-2 call 16 (1 GPU): marching-cubes count pass with the case counts in shared memory (parity + timings); in-kernel
# attribution of the geometry program on coarse-to-fine sized launches (one tile per SM or less: what a lone tile costs).
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --timeout 200 > gpurun_out/r02c16_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02c16_pytest.log
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/r02c16_recon_trace_fv_mc.txt; grep -E "mcubes|per frame" gpurun_out/r02c16_recon_trace_fv_mc.txt | head -12 | cut -c1-120
{
for r in 17 21 33; do
  echo "== grid $r^3 = $((r*r*r)) points = $(( (r*r*r+127)/128 )) tiles"
  timeout 120 python tools/tc_prof.py $r 2>&1 | grep -E "ms per volume"
  MONOPORT_B200_TC_PROF=2 timeout 120 python tools/tc_prof.py $r 2>&1 | grep -E "tc prof\]" | head -21
done
} 2>&1 | tee gpurun_out/r02c16_small_launch_attribution.txt | grep -E "==|ms per|total|wfull|h0ready|ph_L1hid|w_drain0|xready" 
