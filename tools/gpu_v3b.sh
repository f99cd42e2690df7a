#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check.log; tail -8 gpurun_out/tc_check.log
MONOPORT_B200_TC_TRACE=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc trace" > gpurun_out/trace_v3.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-recon --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('bench', round(d['value'],1),'Mpts/s', round(d['ms_per_step'],2),'ms', 'frac',round(d['roofline']['frac'],3), d['clocks'])
PY
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
