#!/bin/bash
mkdir -p gpurun_out
MONOPORT_B200_TC_VER=3 timeout 180 python tools/tc_check.py > gpurun_out/tc_check_v3.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check_v3.log; tail -8 gpurun_out/tc_check_v3.log
if grep -q "dense 257" gpurun_out/tc_check_v3.log; then
MONOPORT_B200_TC_VER=3 MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 2>&1 | tail -25 > gpurun_out/prof_v3_cg1.txt; cat gpurun_out/prof_v3_cg1.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-recon --no-cpu-baseline > gpurun_out/bench_v3_cg1.json 2> gpurun_out/bench_v3_cg1.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v3_cg1.json'))
print('v3 cg1', round(d['value'],1),'Mpts/s', round(d['ms_per_step'],2),'ms', 'frac',round(d['roofline']['frac'],3), d['clocks'])
PY
fi
