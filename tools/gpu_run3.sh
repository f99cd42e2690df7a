#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
# launch list of one bench run (shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c query_tc gpurun_out/launches.csv
# full capture of the dominant kernel on a 129^3 grid (keeps the 40x replay short)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_tc -c 1 -o gpurun_out/prof_tc python bench.py --res 129 --steps 1 --warmup 1 --no-recon --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
