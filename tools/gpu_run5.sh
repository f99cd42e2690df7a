#!/bin/bash
mkdir -p gpurun_out
MONOPORT_B200_TC_CG=1 timeout 180 python tools/tc_check.py > gpurun_out/tc_check_cg1.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check_cg1.log; tail -4 gpurun_out/tc_check_cg1.log
timeout 180 python tools/tc_check.py > gpurun_out/tc_check_cg2.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check_cg2.log; tail -4 gpurun_out/tc_check_cg2.log
MONOPORT_B200_TC_PROF=1 MONOPORT_B200_TC_CG=1 timeout 120 python tools/tc_prof.py 2>&1 | tail -21 > gpurun_out/prof_cg1.txt; cat gpurun_out/prof_cg1.txt
MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 2>&1 | tail -21 > gpurun_out/prof_cg2.txt; cat gpurun_out/prof_cg2.txt
