#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/g0_check.py > gpurun_out/g0_check.log 2>&1; echo "rc=$?" >> gpurun_out/g0_check.log; tail -12 gpurun_out/g0_check.log
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_g0.json 2> gpurun_out/bench_g0.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_g0.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']); print(d['recon'])"
