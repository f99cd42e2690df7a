#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/g0_check.py > gpurun_out/g0_check.log 2>&1; echo "rc=$?" >> gpurun_out/g0_check.log; tail -7 gpurun_out/g0_check.log
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/recon_trace.py > gpurun_out/recon_trace.txt 2>&1; grep -v Warning gpurun_out/recon_trace.txt | head -24
