#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 2; do
  MONOPORT_B200_TC_EXP=$e timeout 120 python tools/tc_check.py 2>&1 | grep "dense 257" | sed "s/^/cg1 exp=$e /"
done | tee gpurun_out/exp_v3.txt
MONOPORT_B200_TC_CG=2 timeout 120 python tools/tc_check.py 2>&1 | grep "dense 257\|20000" | sed "s/^/cg2 /" | tee -a gpurun_out/exp_v3.txt
MONOPORT_B200_TC_CG=2 MONOPORT_B200_TC_EXP=1 timeout 120 python tools/tc_check.py 2>&1 | grep "dense 257" | sed "s/^/cg2 exp=1 /" | tee -a gpurun_out/exp_v3.txt
MONOPORT_B200_TC_CG=2 MONOPORT_B200_TC_TRACE=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc trace" > gpurun_out/trace_v3_cg2.txt
