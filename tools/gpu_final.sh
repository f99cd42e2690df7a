#!/bin/bash
mkdir -p gpurun_out
MONOPORT_B200_TC_VER=3 MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 2>&1 | grep "tc prof" | tail -24 > gpurun_out/prof_v3_cg1.txt; grep -E "total|drain|h0ready|h1ready" gpurun_out/prof_v3_cg1.txt
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']); print(d['recon']); print(d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-250 gpurun_out/bench_reference.json
