#!/bin/bash
mkdir -p gpurun_out
for cg in 1 2; do
MONOPORT_B200_TC_CG=$cg timeout 300 python bench.py --steps 30 --warmup 5 --no-recon --no-cpu-baseline > gpurun_out/bench_cg$cg.json 2> gpurun_out/bench_cg$cg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cg$cg.json'))
print('cg$cg', round(d['value'],1),'Mpts/s', round(d['ms_per_step'],2),'ms', 'frac',round(d['roofline']['frac'],3), d['clocks'])
PY
done
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm --format=csv
