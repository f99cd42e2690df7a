#!/bin/bash
# ~40 s micro-call: parity suite on the new marching-cubes / scan kernels, then one traced frame
mkdir -p gpurun_out
T0=$SECONDS
timeout 100 python -m pytest tests -x -q -m gpu --timeout 60 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s" >> gpurun_out/pytest_gpu2.log; tail -3 gpurun_out/pytest_gpu2.log
timeout 60 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/recon_trace_mc2.txt; echo "trace rc=$? t=$((SECONDS-T0))s"; head -30 gpurun_out/recon_trace_mc2.txt
