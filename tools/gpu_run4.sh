#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/tc_check.py > gpurun_out/tc_check_cg2.log 2>&1; echo "tc_check cg2 rc=$?" >> gpurun_out/tc_check_cg2.log
cat gpurun_out/tc_check_cg2.log
MONOPORT_B200_TC_CG=1 timeout 180 python tools/tc_check.py > gpurun_out/tc_check_cg1.log 2>&1; echo "tc_check cg1 rc=$?" >> gpurun_out/tc_check_cg1.log
cat gpurun_out/tc_check_cg1.log
if grep -q "dense 257" gpurun_out/tc_check_cg2.log; then
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"
  cat gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
fi
