#!/bin/bash
# one development iteration on the GPU box: parity tests, per-kernel timeline of a recon frame, short bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/recon_trace.py > gpurun_out/recon_trace.txt 2>&1; grep -v Warning gpurun_out/recon_trace.txt | head -28
if [ "$1" == "bench" ]; then
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_iter.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']); print(d['recon'])"
fi
