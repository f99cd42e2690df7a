#!/bin/bash
mkdir -p gpurun_out
MONOPORT_B200_TC_VER=3 MONOPORT_B200_TC_CG=2 timeout 180 python tools/tc_check.py > gpurun_out/tc_check_v3cg2.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check_v3cg2.log; tail -4 gpurun_out/tc_check_v3cg2.log
if grep -q "dense 257" gpurun_out/tc_check_v3cg2.log; then
MONOPORT_B200_TC_VER=3 MONOPORT_B200_TC_CG=2 MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 2>&1 | tail -21 > gpurun_out/prof_v3_cg2.txt; cat gpurun_out/prof_v3_cg2.txt
fi
MONOPORT_B200_TC_VER=3 timeout 180 python tools/tc_check.py 2>&1 | tail -3
for cg in 1 2; do
MONOPORT_B200_TC_CG=$cg timeout 300 python bench.py --steps 30 --warmup 5 --no-recon --no-cpu-baseline > gpurun_out/bench_v3_cg$cg.json 2> gpurun_out/bench_v3_cg$cg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v3_cg$cg.json'))
print('v3 cg$cg', round(d['value'],1),'Mpts/s', round(d['ms_per_step'],2),'ms', 'frac',round(d['roofline']['frac'],3), d['clocks'])
PY
done
