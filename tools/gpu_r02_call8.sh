#!/bin/bash
# Round-2 call 8 (1 GPU): A/B in bench conditions -- issue pattern (uniform / round-1) x CTA group (1 / 2)
mkdir -p gpurun_out
for v in "1 0" "1 8" "2 0" "2 8" "1 0" "1 8"; do
  set -- $v
  MONOPORT_B200_TC_CG=$1 MONOPORT_B200_TC_EXP=$2 timeout 200 python bench.py --no-cpu-baseline --no-recon --steps 20 > gpurun_out/r02c8_bench_cg$1_exp$2.json 2> gpurun_out/r02c8_bench.err
  python -c "import json; d=json.load(open('gpurun_out/r02c8_bench_cg$1_exp$2.json')); print('cg$1 exp$2', round(d['value'],1), round(d['ms_per_step'],2), d['clocks'])"
done | tee gpurun_out/r02c8_ab.txt
MONOPORT_B200_TC_CG=1 MONOPORT_B200_TC_EXP=8 MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc prof" | head -20 > gpurun_out/r02c8_tc_inkernel_cycles_cg1_slowissue.txt; cat gpurun_out/r02c8_tc_inkernel_cycles_cg1_slowissue.txt
